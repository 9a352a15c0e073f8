"""MPIJob ``kubeflow.org/v2beta1`` typed model.

Field-for-field counterpart of the reference's Go structs
(pkg/apis/kubeflow/v2beta1/types.go:27-382): same JSON names, same optionality
(Go ``*T`` pointers become ``Optional[T]``), same enums.  ``PodTemplateSpec``
and ``ObjectMeta`` stay JSON-shaped dicts (they are k8s core types the
reference imports rather than defines).  ``from_dict`` / ``to_dict`` round-trip
the YAML users already have; ``deepcopy`` replaces zz_generated.deepcopy.go.
"""
from __future__ import annotations

import copy
import dataclasses
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

from . import constants as C


def _camel(name: str) -> str:
    parts = name.split("_")
    return parts[0] + "".join(p[:1].upper() + p[1:] for p in parts[1:])


class APIObject:
    """dataclass mix-in: camelCase JSON <-> snake_case attributes."""

    def to_dict(self) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for f in dataclasses.fields(self):  # type: ignore[arg-type]
            v = getattr(self, f.name)
            if v is None:
                continue
            omit_empty = f.metadata.get("omitempty", True)
            if omit_empty and (v == "" or v == {} or v == []):
                continue
            out[f.metadata.get("json", _camel(f.name))] = _to_json(v)
        return out

    @classmethod
    def from_dict(cls, d: Optional[Dict[str, Any]]):
        d = d or {}
        kw = {}
        for f in dataclasses.fields(cls):  # type: ignore[arg-type]
            key = f.metadata.get("json", _camel(f.name))
            if key not in d or d[key] is None:
                continue
            conv = f.metadata.get("conv")
            kw[f.name] = conv(d[key]) if conv else copy.deepcopy(d[key])
        return cls(**kw)

    def deepcopy(self):
        return copy.deepcopy(self)


def _to_json(v):
    if isinstance(v, APIObject):
        return v.to_dict()
    if isinstance(v, dict):
        return {k: _to_json(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_to_json(x) for x in v]
    return v


def _int(v):
    return None if v is None else int(v)


@dataclass
class JobCondition(APIObject):
    """types.go:283-306."""
    type: str = ""
    status: str = ""
    reason: str = ""
    message: str = ""
    last_update_time: Optional[str] = None
    last_transition_time: Optional[str] = None


@dataclass
class ReplicaStatus(APIObject):
    """types.go:258-280. ``selector``/``label_selector`` are never set by the controller."""
    active: int = field(default=0, metadata={"omitempty": True})
    succeeded: int = field(default=0, metadata={"omitempty": True})
    failed: int = field(default=0, metadata={"omitempty": True})
    label_selector: Optional[Dict[str, Any]] = None
    selector: str = ""

    def to_dict(self):
        d = super().to_dict()
        for k in ("active", "succeeded", "failed"):
            if not d.get(k):
                d.pop(k, None)
        return d


@dataclass
class JobStatus(APIObject):
    """types.go:226-255."""
    conditions: List[JobCondition] = field(
        default_factory=list, metadata={"conv": lambda v: [JobCondition.from_dict(c) for c in v]})
    replica_statuses: Dict[str, ReplicaStatus] = field(
        default_factory=dict, metadata={"conv": lambda v: {k: ReplicaStatus.from_dict(x) for k, x in v.items()}})
    start_time: Optional[str] = None
    completion_time: Optional[str] = None
    last_reconcile_time: Optional[str] = None


@dataclass
class SchedulingPolicy(APIObject):
    """types.go:44-94 (gang-scheduling knobs)."""
    min_available: Optional[int] = field(default=None, metadata={"conv": _int})
    queue: str = ""
    min_resources: Optional[Dict[str, str]] = None
    priority_class: str = ""
    schedule_timeout_seconds: Optional[int] = field(default=None, metadata={"conv": _int})


@dataclass
class RunPolicy(APIObject):
    """types.go:107-153."""
    clean_pod_policy: Optional[str] = None
    ttl_seconds_after_finished: Optional[int] = field(default=None, metadata={"conv": _int})
    active_deadline_seconds: Optional[int] = field(default=None, metadata={"conv": _int})
    backoff_limit: Optional[int] = field(default=None, metadata={"conv": _int})
    scheduling_policy: Optional[SchedulingPolicy] = field(default=None, metadata={"conv": SchedulingPolicy.from_dict})
    suspend: Optional[bool] = None
    managed_by: Optional[str] = None


@dataclass
class ReplicaSpec(APIObject):
    """types.go:348-362. ``template`` is a core/v1 PodTemplateSpec as a dict."""
    replicas: Optional[int] = field(default=None, metadata={"conv": _int})
    template: Dict[str, Any] = field(default_factory=dict, metadata={"omitempty": False})
    restart_policy: str = ""

    # typed accessors over the opaque template ---------------------------------
    @property
    def pod_spec(self) -> Dict[str, Any]:
        return self.template.setdefault("spec", {})

    @property
    def containers(self) -> List[Dict[str, Any]]:
        return self.pod_spec.get("containers") or []

    def main_container(self) -> Dict[str, Any]:
        return self.containers[0]

    def gpu_limit(self) -> int:
        """``resources.limits["nvidia.com/gpu"]`` of container[0] (0 if unset)."""
        if not self.containers:
            return 0
        res = self.containers[0].get("resources") or {}
        v = (res.get("limits") or {}).get(C.GPU_RESOURCE, (res.get("requests") or {}).get(C.GPU_RESOURCE, 0))
        return int(v)


def _replica_specs(v):
    return {k: (ReplicaSpec.from_dict(x) if x is not None else None) for k, x in (v or {}).items()}


@dataclass
class MPIJobSpec(APIObject):
    """types.go:168-204."""
    slots_per_worker: Optional[int] = field(default=None, metadata={"conv": _int})
    run_launcher_as_worker: Optional[bool] = None
    run_policy: RunPolicy = field(default_factory=RunPolicy, metadata={"conv": RunPolicy.from_dict, "omitempty": False})
    mpi_replica_specs: Optional[Dict[str, Optional[ReplicaSpec]]] = field(
        default=None, metadata={"conv": _replica_specs, "omitempty": False})
    ssh_auth_mount_path: str = ""
    launcher_creation_policy: str = ""
    mpi_implementation: str = field(default="", metadata={"json": "mpiImplementation"})

    def replica(self, rtype: str) -> Optional[ReplicaSpec]:
        return (self.mpi_replica_specs or {}).get(rtype)


@dataclass
class MPIJob(APIObject):
    """types.go:27-32 (TypeMeta + ObjectMeta + Spec + Status)."""
    api_version: str = C.API_VERSION
    kind: str = C.KIND
    metadata: Dict[str, Any] = field(default_factory=dict, metadata={"omitempty": False})
    spec: MPIJobSpec = field(default_factory=MPIJobSpec, metadata={"conv": MPIJobSpec.from_dict, "omitempty": False})
    status: JobStatus = field(default_factory=JobStatus, metadata={"conv": JobStatus.from_dict, "omitempty": False})

    # ObjectMeta sugar ---------------------------------------------------------
    @property
    def name(self) -> str:
        return self.metadata.get("name", "")

    @property
    def namespace(self) -> str:
        return self.metadata.get("namespace", "")

    @property
    def uid(self) -> str:
        return self.metadata.get("uid", "")

    @property
    def labels(self) -> Dict[str, str]:
        return self.metadata.get("labels") or {}

    @property
    def annotations(self) -> Dict[str, str]:
        return self.metadata.get("annotations") or {}

    @property
    def deletion_timestamp(self) -> Optional[str]:
        return self.metadata.get("deletionTimestamp")

    @property
    def key(self) -> str:
        return f"{self.namespace}/{self.name}" if self.namespace else self.name

    def worker_replicas(self) -> int:
        """controller.go:1756-1762 workerReplicas."""
        w = self.spec.replica(C.REPLICA_TYPE_WORKER)
        return int(w.replicas) if w is not None and w.replicas is not None else 0


@dataclass
class MPIJobList(APIObject):
    """types.go:37-41."""
    api_version: str = C.API_VERSION
    kind: str = "MPIJobList"
    metadata: Dict[str, Any] = field(default_factory=dict, metadata={"omitempty": False})
    items: List[MPIJob] = field(default_factory=list,
                                metadata={"conv": lambda v: [MPIJob.from_dict(x) for x in v], "omitempty": False})
