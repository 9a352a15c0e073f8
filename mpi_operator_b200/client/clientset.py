"""Typed clients over the object store.

Counterparts of the reference's generated client machinery (SURVEY.md §2.1 B1,
B2): ``Clientset.kubeflow_v2beta1().mpijobs(ns)`` with Create / Update /
UpdateStatus / Delete / DeleteCollection / Get / List / Watch / Patch / Apply /
ApplyStatus (pkg/client/clientset/versioned/typed/kubeflow/v2beta1/
mpijob.go:37-53), a ``KubeClient`` for the core/batch/scheduling resources the
controller touches, and fakes that record every action for the unit tests
(pkg/client/clientset/versioned/fake/clientset_generated.go:40-140).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional

from ..api import meta as M
from ..api.types import MPIJob, MPIJobList
from . import errors
from .store import ObjectStore, _merge_patch


@dataclass
class Action:
    """One recorded API call (client-go testing.Action)."""
    verb: str
    resource: str
    namespace: str = ""
    name: str = ""
    subresource: str = ""
    obj: Any = None

    def matches(self, verb: str, resource: str) -> bool:
        return self.verb == verb and self.resource == resource


class ResourceClient:
    """Untyped (dict) client for one resource in one namespace."""

    def __init__(self, store: ObjectStore, resource: str, namespace: str = "", recorder: Optional[List[Action]] = None,
                 reactors: Optional[list] = None):
        self.store, self.resource, self.namespace = store, resource, namespace
        self._rec, self._reactors = recorder, reactors

    def _record(self, verb: str, name: str = "", obj=None, sub: str = ""):
        if self._rec is not None:
            self._rec.append(Action(verb, self.resource, self.namespace, name, sub, copy.deepcopy(obj)))
        for r in self._reactors or []:
            r(Action(verb, self.resource, self.namespace, name, sub, obj))  # may raise to inject a fault

    def _ns(self, obj: dict) -> dict:
        if self.namespace and not M.meta(obj).get("namespace"):
            obj = copy.deepcopy(obj)
            M.meta(obj)["namespace"] = self.namespace
        return obj

    def create(self, obj: dict) -> dict:
        obj = self._ns(obj)
        self._record("create", M.name_of(obj), obj)
        return self.store.create(self.resource, obj)

    def update(self, obj: dict) -> dict:
        obj = self._ns(obj)
        self._record("update", M.name_of(obj), obj)
        return self.store.update(self.resource, obj)

    def update_status(self, obj: dict) -> dict:
        obj = self._ns(obj)
        self._record("update", M.name_of(obj), obj, "status")
        return self.store.update_status(self.resource, obj)

    def delete(self, name: str) -> None:
        self._record("delete", name)
        self.store.delete(self.resource, self.namespace, name)

    def delete_collection(self, label_selector: Optional[Dict[str, str]] = None) -> int:
        self._record("delete-collection")
        return self.store.delete_collection(self.resource, self.namespace or None, label_selector)

    def get(self, name: str) -> dict:
        self._record("get", name)
        return self.store.get(self.resource, self.namespace, name)

    def list(self, label_selector: Optional[Dict[str, str]] = None) -> List[dict]:
        self._record("list")
        return self.store.list(self.resource, self.namespace or None, label_selector)

    def watch(self, handler, replay: bool = True) -> Callable[[], None]:
        self._record("watch")
        ns = self.namespace

        def filtered(etype, obj, old):
            if not ns or M.namespace_of(obj) == ns:
                handler(etype, obj, old)
        return self.store.watch(self.resource, filtered, replay)

    def patch(self, name: str, patch: dict, subresource: str = "") -> dict:
        self._record("patch", name, patch, subresource)
        return self.store.patch(self.resource, self.namespace, name, patch, status=(subresource == "status"))

    def apply(self, config: dict, field_manager: str = "", subresource: str = "") -> dict:
        """Server-side apply, approximated as create-or-merge of the applied fields."""
        name = M.name_of(config)
        self._record("patch", name, config, subresource)
        try:
            cur = self.store.get(self.resource, self.namespace, name)
        except errors.ApiError as e:
            if not errors.is_not_found(e):
                raise
            return self.store.create(self.resource, self._ns(config))
        merged = _merge_patch(cur, config)
        M.meta(merged)["resourceVersion"] = M.meta(cur)["resourceVersion"]
        return self.store.update_status(self.resource, merged) if subresource == "status" else self.store.update(self.resource, merged)


class MPIJobInterface:
    """Typed MPIJob client (mpijob.go:37-53)."""

    def __init__(self, rc: ResourceClient):
        self._rc = rc

    def create(self, job: MPIJob) -> MPIJob:
        return MPIJob.from_dict(self._rc.create(job.to_dict()))

    def update(self, job: MPIJob) -> MPIJob:
        return MPIJob.from_dict(self._rc.update(job.to_dict()))

    def update_status(self, job: MPIJob) -> MPIJob:
        return MPIJob.from_dict(self._rc.update_status(job.to_dict()))

    def delete(self, name: str) -> None:
        self._rc.delete(name)

    def delete_collection(self, label_selector=None) -> int:
        return self._rc.delete_collection(label_selector)

    def get(self, name: str) -> MPIJob:
        return MPIJob.from_dict(self._rc.get(name))

    def list(self, label_selector=None) -> MPIJobList:
        return MPIJobList(items=[MPIJob.from_dict(o) for o in self._rc.list(label_selector)])

    def watch(self, handler, replay: bool = True):
        return self._rc.watch(lambda t, o, old: handler(t, MPIJob.from_dict(o), MPIJob.from_dict(old) if old else None), replay)

    def patch(self, name: str, patch: dict, subresource: str = "") -> MPIJob:
        return MPIJob.from_dict(self._rc.patch(name, patch, subresource))

    def apply(self, config, field_manager: str = "mpi-operator") -> MPIJob:
        body = config.build() if hasattr(config, "build") else config
        return MPIJob.from_dict(self._rc.apply(body, field_manager))

    def apply_status(self, config, field_manager: str = "mpi-operator") -> MPIJob:
        body = config.build() if hasattr(config, "build") else config
        return MPIJob.from_dict(self._rc.apply(body, field_manager, "status"))


class KubeflowV2beta1Client:
    def __init__(self, store, recorder=None, reactors=None):
        self._a = (store, recorder, reactors)

    def mpijobs(self, namespace: str = "") -> MPIJobInterface:
        store, rec, rx = self._a
        return MPIJobInterface(ResourceClient(store, "mpijobs", namespace, rec, rx))


class Clientset:
    """versioned.Clientset (pkg/client/clientset/versioned/clientset.go:41-118)."""

    def __init__(self, store: ObjectStore, recorder=None, reactors=None):
        self.store = store
        self._v2beta1 = KubeflowV2beta1Client(store, recorder, reactors)

    def kubeflow_v2beta1(self) -> KubeflowV2beta1Client:
        return self._v2beta1

    def discovery_has_mpijob_crd(self) -> bool:
        """server.go:302-314 checkCRDExists analogue."""
        return "mpijobs" in self.store._objs


class KubeClient:
    """The slice of kubernetes.Interface the controller uses."""

    def __init__(self, store: ObjectStore, recorder=None, reactors=None):
        self.store, self._rec, self._rx = store, recorder, reactors

    def _rc(self, resource, ns=""):
        return ResourceClient(self.store, resource, ns, self._rec, self._rx)

    def pods(self, ns=""): return self._rc("pods", ns)  # noqa: E704
    def services(self, ns=""): return self._rc("services", ns)  # noqa: E704
    def config_maps(self, ns=""): return self._rc("configmaps", ns)  # noqa: E704
    def secrets(self, ns=""): return self._rc("secrets", ns)  # noqa: E704
    def events(self, ns=""): return self._rc("events", ns)  # noqa: E704
    def jobs(self, ns=""): return self._rc("jobs", ns)  # noqa: E704
    def leases(self, ns=""): return self._rc("leases", ns)  # noqa: E704
    def priority_classes(self): return self._rc("priorityclasses", "")  # noqa: E704
    def volcano_pod_groups(self, ns=""): return self._rc("volcano-podgroups", ns)  # noqa: E704
    def sched_pod_groups(self, ns=""): return self._rc("sched-podgroups", ns)  # noqa: E704


class FakeClientset(Clientset):
    """fake.NewSimpleClientset(objs...): private tracker + recorded actions + reactors."""

    def __init__(self, *objects: dict, store: Optional[ObjectStore] = None):
        self.actions: List[Action] = []
        self.reactors: list = []
        store = store or ObjectStore()
        for o in objects:
            store.create(_resource_of(o), o)
        super().__init__(store, self.actions, self.reactors)

    def kube(self) -> KubeClient:
        return KubeClient(self.store, self.actions, self.reactors)

    def prepend_reactor(self, fn) -> None:
        self.reactors.insert(0, fn)

    def clear_actions(self) -> None:
        del self.actions[:]


def _resource_of(obj: dict) -> str:
    from .store import RESOURCES
    for r, (api_version, kind, _) in RESOURCES.items():
        if obj.get("kind") == kind and obj.get("apiVersion", api_version) == api_version:
            return r
    raise ValueError(f"unknown kind {obj.get('kind')!r}")
