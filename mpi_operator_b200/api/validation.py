"""Validation of MPIJob objects, with Kubernetes-style field errors.

Reference: pkg/apis/kubeflow/validation/validation.go:29-160 (same checks, same
field paths, same error types and detail strings; the test table
validation_test.go:29-415 is ported in tests/test_api.py).  ``FieldError`` /
``ErrorList`` mirror k8s.io/apimachinery/pkg/util/validation/field.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Any, List, Optional

from . import constants as C
from .types import MPIJob, MPIJobSpec, ReplicaSpec, RunPolicy

ERR_REQUIRED = "FieldValueRequired"
ERR_INVALID = "FieldValueInvalid"
ERR_NOT_SUPPORTED = "FieldValueNotSupported"

VALID_CLEAN_POLICIES = sorted([C.CLEAN_POD_POLICY_NONE, C.CLEAN_POD_POLICY_RUNNING, C.CLEAN_POD_POLICY_ALL])
VALID_MPI_IMPLEMENTATIONS = sorted([C.MPI_IMPLEMENTATION_OPENMPI, C.MPI_IMPLEMENTATION_INTEL, C.MPI_IMPLEMENTATION_MPICH])
VALID_RESTART_POLICIES = sorted([C.RESTART_POLICY_NEVER, C.RESTART_POLICY_ON_FAILURE])
VALID_MANAGED_BY = sorted([C.MULTIKUEUE_CONTROLLER, C.KUBEFLOW_JOB_CONTROLLER])

_DNS1035_RE = re.compile(r"^[a-z]([-a-z0-9]*[a-z0-9])?$")
_DNS1035_MAX = 63
_DNS1035_MSG = ("a DNS-1035 label must consist of lower case alphanumeric characters or '-', start with an "
                "alphabetic character, and end with an alphanumeric character (e.g. 'my-name',  or 'abc-123', "
                "regex used for validation is '[a-z]([-a-z0-9]*[a-z0-9])?')")


@dataclass
class FieldError:
    type: str
    field: str
    bad_value: Any = None
    detail: str = ""
    origin: str = ""

    def error(self) -> str:
        if self.type == ERR_REQUIRED:
            s = f"{self.field}: Required value"
        elif self.type == ERR_NOT_SUPPORTED:
            s = f"{self.field}: Unsupported value: {_quote(self.bad_value)}"
        else:
            s = f"{self.field}: Invalid value: {_quote(self.bad_value)}"
        return f"{s}: {self.detail}" if self.detail else s

    __str__ = error


class ErrorList(list):
    def to_aggregate(self) -> str:
        """field.ErrorList.ToAggregate().Error() formatting."""
        msgs = []
        for e in self:
            m = e.error()
            if m not in msgs:
                msgs.append(m)
        if not msgs:
            return ""
        return msgs[0] if len(msgs) == 1 else "[" + ", ".join(msgs) + "]"


def _quote(v) -> str:
    if isinstance(v, str):
        return '"' + v + '"'
    return str(v)


def required(path: str, detail: str) -> FieldError:
    return FieldError(ERR_REQUIRED, path, "", detail)


def invalid(path: str, value, detail: str, origin: str = "") -> FieldError:
    return FieldError(ERR_INVALID, path, value, detail, origin)


def not_supported(path: str, value, valid: List[str]) -> FieldError:
    return FieldError(ERR_NOT_SUPPORTED, path, value, "supported values: " + ", ".join(f'"{v}"' for v in valid))


def is_dns1035_label(value: str) -> List[str]:
    errs = []
    if len(value) > _DNS1035_MAX:
        errs.append(f"must be no more than {_DNS1035_MAX} characters")
    if not _DNS1035_RE.match(value):
        errs.append(_DNS1035_MSG)
    return errs


def _nonneg(value: int, path: str) -> ErrorList:
    errs = ErrorList()
    if value < 0:
        errs.append(invalid(path, value, "must be greater than or equal to 0", origin="minimum"))
    return errs


def validate_mpijob(job: MPIJob) -> ErrorList:
    """validation.go:49-53."""
    errs = _validate_name(job)
    errs.extend(_validate_spec(job.spec, "spec"))
    return errs


def _validate_name(job: MPIJob) -> ErrorList:
    """validation.go:55-68: the *largest* worker hostname must be a DNS-1035 label."""
    errs = ErrorList()
    replicas = 1
    w = job.spec.replica(C.REPLICA_TYPE_WORKER)
    if w is not None and w.replicas is not None and w.replicas > 0:
        replicas = w.replicas
    host = f"{job.name}-worker-{replicas - 1}"
    problems = is_dns1035_label(host)
    if problems:
        errs.append(invalid("metadata.name", job.name,
                            f'will not able to create pod and service with invalid DNS label "{host}": ' + ", ".join(problems)))
    return errs


def _validate_spec(spec: MPIJobSpec, path: str) -> ErrorList:
    """validation.go:70-85."""
    errs = _validate_replica_specs(spec.mpi_replica_specs, f"{path}.mpiReplicaSpecs")
    if spec.slots_per_worker is None:
        errs.append(required(f"{path}.slotsPerWorker", "must have number of slots per worker"))
    else:
        errs.extend(_nonneg(spec.slots_per_worker, f"{path}.slotsPerWorker"))
    errs.extend(_validate_run_policy(spec.run_policy, f"{path}.runPolicy"))
    if spec.ssh_auth_mount_path == "":
        errs.append(required(f"{path}.sshAuthMountPath", "must have a mount path for SSH credentials"))
    if spec.mpi_implementation not in VALID_MPI_IMPLEMENTATIONS:
        errs.append(not_supported(f"{path}.mpiImplementation", spec.mpi_implementation, VALID_MPI_IMPLEMENTATIONS))
    return errs


def _validate_run_policy(policy: RunPolicy, path: str) -> ErrorList:
    """validation.go:87-110."""
    errs = ErrorList()
    if policy.clean_pod_policy is None:
        errs.append(required(f"{path}.cleanPodPolicy", "must have clean Pod policy"))
    elif policy.clean_pod_policy not in VALID_CLEAN_POLICIES:
        errs.append(not_supported(f"{path}.cleanPodPolicy", policy.clean_pod_policy, VALID_CLEAN_POLICIES))
    if policy.ttl_seconds_after_finished is not None:
        errs.extend(_nonneg(policy.ttl_seconds_after_finished, f"{path}.ttlSecondsAfterFinished"))
    if policy.active_deadline_seconds is not None:
        errs.extend(_nonneg(policy.active_deadline_seconds, f"{path}.activeDeadlineSeconds"))
    if policy.backoff_limit is not None:
        errs.extend(_nonneg(policy.backoff_limit, f"{path}.backoffLimit"))
    if policy.managed_by is not None and policy.managed_by not in VALID_MANAGED_BY:
        errs.append(not_supported(f"{path}.managedBy", policy.managed_by, VALID_MANAGED_BY))
    return errs


def _validate_replica_specs(specs, path: str) -> ErrorList:
    """validation.go:112-121."""
    errs = ErrorList()
    if specs is None:
        errs.append(required(path, "must have replica specs"))
        return errs
    errs.extend(_validate_launcher(specs.get(C.REPLICA_TYPE_LAUNCHER), f"{path}[{C.REPLICA_TYPE_LAUNCHER}]"))
    errs.extend(_validate_worker(specs.get(C.REPLICA_TYPE_WORKER), f"{path}[{C.REPLICA_TYPE_WORKER}]"))
    return errs


def _validate_launcher(spec: Optional[ReplicaSpec], path: str) -> ErrorList:
    """validation.go:123-134."""
    errs = ErrorList()
    if spec is None:
        errs.append(required(path, f"must have {C.REPLICA_TYPE_LAUNCHER} replica spec"))
        return errs
    errs.extend(_validate_replica(spec, path))
    if spec.replicas is not None and spec.replicas != 1:
        errs.append(invalid(f"{path}.replicas", spec.replicas, "must be 1"))
    return errs


def _validate_worker(spec: Optional[ReplicaSpec], path: str) -> ErrorList:
    """validation.go:136-146."""
    errs = ErrorList()
    if spec is None:
        return errs
    errs.extend(_validate_replica(spec, path))
    if spec.replicas is not None and spec.replicas <= 0:
        errs.append(invalid(f"{path}.replicas", spec.replicas, "must be greater than or equal to 1"))
    return errs


def _validate_replica(spec: ReplicaSpec, path: str) -> ErrorList:
    """validation.go:148-160."""
    errs = ErrorList()
    if spec.replicas is None:
        errs.append(required(f"{path}.replicas", "must define number of replicas"))
    if spec.restart_policy not in VALID_RESTART_POLICIES:
        errs.append(not_supported(f"{path}.restartPolicy", spec.restart_policy, VALID_RESTART_POLICIES))
    if len(spec.containers) == 0:
        errs.append(required(f"{path}.template.spec.containers", "must define at least one container"))
    return errs
