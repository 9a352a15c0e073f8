"""Apply configurations: fluent builders for server-side apply.

Counterpart of pkg/client/applyconfiguration/kubeflow/v2beta1/*.go (e.g.
mpijobspec.go:49-113, runpolicy.go:68-126): ``MPIJob(name, ns).with_spec(
MPIJobSpec().with_slots_per_worker(2)...)``; ``build()`` yields the partial
object that ``MPIJobInterface.apply`` merges.  The ``with_*`` methods are
generated from the typed model's dataclass fields instead of by code-gen.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Dict

from ..api import constants as C
from ..api import types as T


def _camel(name: str) -> str:
    parts = name.split("_")
    return parts[0] + "".join(p[:1].upper() + p[1:] for p in parts[1:])


class _ApplyConfiguration:
    _model = None  # dataclass the builder mirrors

    def __init__(self):
        self._fields: Dict[str, Any] = {}

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if cls._model is None:
            return
        for f in dataclasses.fields(cls._model):
            json_name = f.metadata.get("json", _camel(f.name))
            if f"with_{f.name}" in cls.__dict__:
                continue  # hand-written accumulating setter wins

            def setter(self, value, _k=json_name):
                self._fields[_k] = value
                return self
            setter.__name__ = f"with_{f.name}"
            setter.__doc__ = f"Set ``{json_name}`` in the declarative configuration."
            setattr(cls, f"with_{f.name}", setter)

    def build(self) -> Dict[str, Any]:
        def conv(v):
            if isinstance(v, _ApplyConfiguration):
                return v.build()
            if isinstance(v, dict):
                return {k: conv(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return [conv(x) for x in v]
            return v
        return {k: conv(v) for k, v in self._fields.items()}


class JobConditionApplyConfiguration(_ApplyConfiguration):
    _model = T.JobCondition


class ReplicaStatusApplyConfiguration(_ApplyConfiguration):
    _model = T.ReplicaStatus


class JobStatusApplyConfiguration(_ApplyConfiguration):
    _model = T.JobStatus

    def with_conditions(self, *conds):
        self._fields.setdefault("conditions", []).extend(conds)
        return self


class SchedulingPolicyApplyConfiguration(_ApplyConfiguration):
    _model = T.SchedulingPolicy


class RunPolicyApplyConfiguration(_ApplyConfiguration):
    _model = T.RunPolicy


class ReplicaSpecApplyConfiguration(_ApplyConfiguration):
    _model = T.ReplicaSpec


class MPIJobSpecApplyConfiguration(_ApplyConfiguration):
    _model = T.MPIJobSpec

    def with_mpi_replica_specs(self, entries: Dict[str, Any]):
        self._fields.setdefault("mpiReplicaSpecs", {}).update(entries)
        return self


class MPIJobApplyConfiguration(_ApplyConfiguration):
    _model = T.MPIJob

    def __init__(self, name: str = "", namespace: str = ""):
        super().__init__()
        self._fields["apiVersion"] = C.API_VERSION
        self._fields["kind"] = C.KIND
        self._fields["metadata"] = {}
        if name:
            self.with_name(name)
        if namespace:
            self.with_namespace(namespace)

    def with_name(self, v):
        self._fields["metadata"]["name"] = v
        return self

    def with_namespace(self, v):
        self._fields["metadata"]["namespace"] = v
        return self

    def with_labels(self, entries: Dict[str, str]):
        self._fields["metadata"].setdefault("labels", {}).update(entries)
        return self

    def with_annotations(self, entries: Dict[str, str]):
        self._fields["metadata"].setdefault("annotations", {}).update(entries)
        return self

    def with_owner_references(self, *refs):
        self._fields["metadata"].setdefault("ownerReferences", []).extend(refs)
        return self

    def with_finalizers(self, *vals):
        self._fields["metadata"].setdefault("finalizers", []).extend(vals)
        return self


# constructor sugar, same names as the Go package-level functions
def MPIJob(name: str, namespace: str) -> MPIJobApplyConfiguration:  # noqa: N802
    return MPIJobApplyConfiguration(name, namespace)


def MPIJobSpec() -> MPIJobSpecApplyConfiguration:  # noqa: N802
    return MPIJobSpecApplyConfiguration()


def RunPolicy() -> RunPolicyApplyConfiguration:  # noqa: N802
    return RunPolicyApplyConfiguration()


def SchedulingPolicy() -> SchedulingPolicyApplyConfiguration:  # noqa: N802
    return SchedulingPolicyApplyConfiguration()


def ReplicaSpec() -> ReplicaSpecApplyConfiguration:  # noqa: N802
    return ReplicaSpecApplyConfiguration()


def JobStatus() -> JobStatusApplyConfiguration:  # noqa: N802
    return JobStatusApplyConfiguration()


def JobCondition() -> JobConditionApplyConfiguration:  # noqa: N802
    return JobConditionApplyConfiguration()


def ReplicaStatus() -> ReplicaStatusApplyConfiguration:  # noqa: N802
    return ReplicaStatusApplyConfiguration()
