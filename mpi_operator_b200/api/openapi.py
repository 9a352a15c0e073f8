"""OpenAPI definitions, Swagger 2.0 document and the CRD manifest for MPIJob.

Counterparts of the reference's generated artefacts: ``GetOpenAPIDefinitions``
(pkg/apis/kubeflow/v2beta1/zz_generated.openapi.go:31), ``swagger.json``
(pkg/apis/kubeflow/v2beta1/swagger.json, the SDK generator's input,
hack/python-sdk/main.go:33-106) and the controller-gen CRD
(manifests/base/kubeflow.org_mpijobs.yaml: scope Namespaced, served+storage,
status sub-resource, schema defaults/enums, ``required: [mpiReplicaSpecs]``,
conditions as a list-map keyed by ``type``).  Generated from the typed model
(``api/types.py``) by reflection instead of by Go code generators; PodTemplateSpec
is kept open (``x-kubernetes-preserve-unknown-fields``) rather than inlined twice.
"""
from __future__ import annotations

import dataclasses
import typing
from typing import Any, Dict

from . import constants as C
from . import types as T

PREFIX = "v2beta1."

_DESCRIPTIONS = {
    "MPIJobSpec.slots_per_worker": "Specifies the number of slots per worker used in hostfile. Defaults to 1.",
    "MPIJobSpec.run_launcher_as_worker": "RunLauncherAsWorker indicates whether to run worker process in launcher. Defaults to false.",
    "MPIJobSpec.run_policy": "RunPolicy encapsulates various runtime policies of the job.",
    "MPIJobSpec.mpi_replica_specs": "MPIReplicaSpecs contains maps from `MPIReplicaType` to `ReplicaSpec` that specify the MPI replicas to run.",
    "MPIJobSpec.ssh_auth_mount_path": "SSHAuthMountPath is the directory where SSH keys are mounted. Defaults to \"/root/.ssh\".",
    "MPIJobSpec.launcher_creation_policy": "launcherCreationPolicy if WaitForWorkersReady, the launcher is created only after all workers are in Ready state. Defaults to AtStartup.",
    "MPIJobSpec.mpi_implementation": "MPIImplementation is the MPI implementation. Options are \"OpenMPI\" (default), \"Intel\" and \"MPICH\".",
    "RunPolicy.clean_pod_policy": "CleanPodPolicy defines the policy to kill pods after the job completes. Default to Running.",
    "RunPolicy.ttl_seconds_after_finished": "TTLSecondsAfterFinished is the TTL to clean up jobs. Default to infinite.",
    "RunPolicy.active_deadline_seconds": "Specifies the duration in seconds relative to the startTime that the job may be active before the system tries to terminate it.",
    "RunPolicy.backoff_limit": "Optional number of retries before marking this job failed.",
    "RunPolicy.scheduling_policy": "SchedulingPolicy defines the policy related to scheduling, e.g. gang-scheduling",
    "RunPolicy.suspend": "suspend specifies whether the MPIJob controller should create Pods or not. Defaults to false.",
    "RunPolicy.managed_by": "ManagedBy is used to indicate the controller or entity that manages a MPIJob. The field is immutable.",
    "SchedulingPolicy.min_available": "MinAvailable defines the minimal number of member to run the PodGroup. Defaults to NUM(workers)+1.",
    "SchedulingPolicy.queue": "Queue defines the queue name to allocate resource for PodGroup (volcano only).",
    "SchedulingPolicy.min_resources": "MinResources defines the minimal resources of members to run the PodGroup.",
    "SchedulingPolicy.priority_class": "PriorityClass defines the PodGroup's PriorityClass (volcano only).",
    "SchedulingPolicy.schedule_timeout_seconds": "SchedulerTimeoutSeconds defines the maximal time of members to wait before run the PodGroup (scheduler-plugins only).",
}
_ENUMS = {
    "MPIJobSpec.launcher_creation_policy": [C.LAUNCHER_CREATION_POLICY_AT_STARTUP, C.LAUNCHER_CREATION_POLICY_WAIT_FOR_WORKERS_READY],
    "MPIJobSpec.mpi_implementation": [C.MPI_IMPLEMENTATION_OPENMPI, C.MPI_IMPLEMENTATION_INTEL, C.MPI_IMPLEMENTATION_MPICH],
}
_DEFAULTS = {
    "MPIJobSpec.slots_per_worker": 1, "MPIJobSpec.ssh_auth_mount_path": "/root/.ssh",
    "MPIJobSpec.launcher_creation_policy": C.LAUNCHER_CREATION_POLICY_AT_STARTUP,
    "MPIJobSpec.mpi_implementation": C.MPI_IMPLEMENTATION_OPENMPI, "RunPolicy.suspend": False,
}
_REQUIRED = {"MPIJobSpec": ["mpiReplicaSpecs"], "JobCondition": ["type", "status"], "MPIJobList": ["items"]}
_INT64 = {"RunPolicy.active_deadline_seconds"}
_MODELS = [T.JobCondition, T.ReplicaStatus, T.JobStatus, T.SchedulingPolicy, T.RunPolicy, T.ReplicaSpec, T.MPIJobSpec,
           T.MPIJob, T.MPIJobList]


def _camel(name: str) -> str:
    parts = name.split("_")
    return parts[0] + "".join(p[:1].upper() + p[1:] for p in parts[1:])


def _schema_for(tp, owner: str, fname: str, refs: bool) -> Dict[str, Any]:
    origin = typing.get_origin(tp)
    args = typing.get_args(tp)
    if origin is typing.Union:
        inner = [a for a in args if a is not type(None)]
        return _schema_for(inner[0], owner, fname, refs)
    if dataclasses.is_dataclass(tp):
        return {"$ref": f"#/definitions/{PREFIX}{tp.__name__}"} if refs else schema_of(tp, refs=False)
    if origin in (list, typing.List):
        return {"type": "array", "items": _schema_for(args[0], owner, fname, refs)}
    if origin in (dict, typing.Dict):
        key = f"{owner}.{fname}"
        if key in ("ReplicaSpec.template", "MPIJob.metadata", "MPIJobList.metadata", "ReplicaStatus.label_selector"):
            return {"type": "object", "x-kubernetes-preserve-unknown-fields": True}
        return {"type": "object", "additionalProperties": _schema_for(args[1], owner, fname, refs)}
    if tp is int:
        return {"type": "integer", "format": "int64" if f"{owner}.{fname}" in _INT64 else "int32"}
    if tp is bool:
        return {"type": "boolean"}
    if tp is str:
        if fname.endswith("_time"):
            return {"type": "string", "format": "date-time"}
        return {"type": "string"}
    return {"type": "object", "x-kubernetes-preserve-unknown-fields": True}


def schema_of(cls, refs: bool = True) -> Dict[str, Any]:
    hints = typing.get_type_hints(cls)
    props: Dict[str, Any] = {}
    for f in dataclasses.fields(cls):
        key = f"{cls.__name__}.{f.name}"
        s = _schema_for(hints[f.name], cls.__name__, f.name, refs)
        if key in _DESCRIPTIONS:
            s = {**s, "description": _DESCRIPTIONS[key]} if "$ref" not in s else s
        if key in _ENUMS:
            s["enum"] = _ENUMS[key]
        if key in _DEFAULTS and not refs:
            s["default"] = _DEFAULTS[key]
        props[f.metadata.get("json", _camel(f.name))] = s
    out: Dict[str, Any] = {"type": "object", "properties": props}
    if cls.__name__ in _REQUIRED:
        out["required"] = _REQUIRED[cls.__name__]
    if cls is T.JobStatus and not refs:
        props["conditions"]["x-kubernetes-list-type"] = "map"
        props["conditions"]["x-kubernetes-list-map-keys"] = ["type"]
    return out


def get_openapi_definitions() -> Dict[str, Any]:
    """name -> schema, the shape of GetOpenAPIDefinitions."""
    return {PREFIX + c.__name__: schema_of(c, refs=True) for c in _MODELS}


def swagger() -> Dict[str, Any]:
    """Swagger 2.0 document (the SDK generator input)."""
    return {"swagger": "2.0", "info": {"title": "mpijob", "description": "Python SDK for MPI-Operator", "version": "v0.1"},
            "paths": {}, "definitions": get_openapi_definitions()}


def crd() -> Dict[str, Any]:
    """apiextensions.k8s.io/v1 CustomResourceDefinition for MPIJob."""
    root = schema_of(T.MPIJob, refs=False)
    root["properties"]["apiVersion"] = {"type": "string"}
    root["properties"]["kind"] = {"type": "string"}
    root["properties"]["metadata"] = {"type": "object"}
    return {
        "apiVersion": "apiextensions.k8s.io/v1", "kind": "CustomResourceDefinition",
        "metadata": {"name": f"{C.PLURAL}.{C.GROUP_NAME}", "annotations": {"generated-by": "mpi_operator_b200.api.openapi"}},
        "spec": {
            "group": C.GROUP_NAME,
            "names": {"kind": C.KIND, "listKind": "MPIJobList", "plural": C.PLURAL, "singular": C.SINGULAR},
            "scope": "Namespaced",
            "versions": [{"name": C.GROUP_VERSION, "served": True, "storage": True, "subresources": {"status": {}},
                          "schema": {"openAPIV3Schema": root}}],
        },
    }
