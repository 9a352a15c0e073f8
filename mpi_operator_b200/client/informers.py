"""Shared informers + listers over the object store.

Counterparts of pkg/client/informers/externalversions/factory.go:57-261 and
pkg/client/listers/kubeflow/v2beta1/mpijob.go:43-68: a watch-fed local cache
(``Indexer``) per resource, ResourceEventHandler fan-out (add/update/delete),
``has_synced``, namespace scoping (``WithNamespace``) and read-only listers.
Unit tests may pre-load ``indexer`` by hand instead of starting the informer,
exactly like the reference's fixture (mpi_job_controller_test.go:202-263).
"""
from __future__ import annotations

import copy
import threading
from typing import Callable, Dict, List, Optional

from ..api import meta as M
from ..api.types import MPIJob
from . import errors
from .store import ADDED, DELETED, MODIFIED, ObjectStore


class Indexer:
    """Thread-safe object cache keyed by namespace/name, with an inverted index over labels: a reconcile lists "the pods
    of this job" (one label pair) several times and must not scan, sort or copy every object in the cache for it."""

    def __init__(self):
        self._lock = threading.RLock()
        self._items: Dict[str, dict] = {}
        self._by_label: Dict[tuple, set] = {}

    def _unindex(self, key: str, obj: dict) -> None:
        for pair in (M.meta(obj).get("labels") or {}).items():
            keys = self._by_label.get(pair)
            if keys is not None:
                keys.discard(key)
                if not keys:
                    del self._by_label[pair]

    # Cached objects are replaced wholesale, never mutated in place, so copies are taken outside the lock.
    def add(self, obj: dict) -> None:
        mine = copy.deepcopy(obj)
        key = M.key_of(obj)
        with self._lock:
            old = self._items.get(key)
            if old is not None:
                self._unindex(key, old)
            self._items[key] = mine
            for pair in (M.meta(mine).get("labels") or {}).items():
                self._by_label.setdefault(pair, set()).add(key)

    update = add

    def delete(self, obj: dict) -> None:
        key = M.key_of(obj)
        with self._lock:
            old = self._items.pop(key, None)
            if old is not None:
                self._unindex(key, old)

    def get_by_key(self, key: str) -> Optional[dict]:
        with self._lock:
            o = self._items.get(key)
        return copy.deepcopy(o) if o is not None else None

    def list(self, namespace: str = "", selector: Optional[Dict[str, str]] = None) -> List[dict]:
        """Sorted by key. Equality selectors start from the smallest matching label bucket; namespace and the full selector
        are applied before anything is copied."""
        with self._lock:
            if selector:
                buckets = [self._by_label.get(pair) for pair in selector.items()]
                if any(b is None for b in buckets):
                    return []
                keys = sorted(min(buckets, key=len))
            else:
                keys = sorted(self._items)
            hits = []
            for key in keys:
                o = self._items[key]
                if namespace and M.namespace_of(o) != namespace:
                    continue
                if selector and not M.label_selector_matches(selector, M.meta(o).get("labels")):
                    continue
                hits.append(o)
        return [copy.deepcopy(o) for o in hits]


class SharedIndexInformer:
    def __init__(self, store: ObjectStore, resource: str, namespace: str = ""):
        self.store, self.resource, self.namespace = store, resource, namespace
        self.indexer = Indexer()
        self._handlers: List[tuple] = []
        self._synced = False
        self._cancel: Optional[Callable[[], None]] = None

    def add_event_handler(self, add=None, update=None, delete=None) -> None:
        self._handlers.append((add, update, delete))

    def has_synced(self) -> bool:
        return self._synced

    def _on_event(self, etype: str, obj: dict, old: Optional[dict]) -> None:
        if self.namespace and M.namespace_of(obj) != self.namespace:
            return
        if etype == DELETED:
            self.indexer.delete(obj)
        else:
            self.indexer.add(obj)
        for add, update, delete in self._handlers:
            if etype == ADDED and add:
                add(obj)
            elif etype == MODIFIED and update:
                update(old if old is not None else obj, obj)
            elif etype == DELETED and delete:
                delete(obj)

    def run(self) -> None:
        if self._cancel is None:
            self._cancel = self.store.watch(self.resource, self._on_event, replay=True)
            self._synced = True

    def stop(self) -> None:
        if self._cancel:
            self._cancel()
            self._cancel = None


class NamespaceLister:
    def __init__(self, indexer: Indexer, resource: str, namespace: str, convert=None):
        self._ix, self._res, self._ns, self._conv = indexer, resource, namespace, convert or (lambda o: o)

    def get(self, name: str):
        o = self._ix.get_by_key(f"{self._ns}/{name}" if self._ns else name)
        if o is None:
            raise errors.not_found(self._res, name)
        return self._conv(o)

    def list(self, selector: Optional[Dict[str, str]] = None) -> list:
        return [self._conv(o) for o in self._ix.list(self._ns, selector)]


class Lister:
    """``lister.pods(ns).get(name)`` style access for any resource."""

    def __init__(self, indexer: Indexer, resource: str, convert=None):
        self._ix, self._res, self._conv = indexer, resource, convert

    def namespaced(self, namespace: str) -> NamespaceLister:
        return NamespaceLister(self._ix, self._res, namespace, self._conv)

    def get(self, name: str):  # cluster-scoped resources (PriorityClass)
        return NamespaceLister(self._ix, self._res, "", self._conv).get(name)

    def list(self, selector=None):
        return NamespaceLister(self._ix, self._res, "", self._conv).list(selector)


class MPIJobLister(Lister):
    """listers/kubeflow/v2beta1/mpijob.go: typed results."""

    def __init__(self, indexer: Indexer):
        super().__init__(indexer, "mpijobs", MPIJob.from_dict)

    def mpijobs(self, namespace: str) -> NamespaceLister:
        return self.namespaced(namespace)


class SharedInformerFactory:
    """NewSharedInformerFactoryWithOptions(client, resync=0, WithNamespace(ns))."""

    def __init__(self, store: ObjectStore, namespace: str = ""):
        self.store, self.namespace = store, namespace
        self._informers: Dict[str, SharedIndexInformer] = {}

    def informer_for(self, resource: str) -> SharedIndexInformer:
        if resource not in self._informers:
            from .store import RESOURCES
            ns = self.namespace if RESOURCES[resource][2] else ""
            self._informers[resource] = SharedIndexInformer(self.store, resource, ns)
        return self._informers[resource]

    def lister_for(self, resource: str) -> Lister:
        ix = self.informer_for(resource).indexer
        return MPIJobLister(ix) if resource == "mpijobs" else Lister(ix, resource)

    # typed sugar mirroring factory.Kubeflow().V2beta1().MPIJobs()
    def mpijobs(self) -> SharedIndexInformer:
        return self.informer_for("mpijobs")

    def start(self) -> None:
        for inf in self._informers.values():
            inf.run()

    def wait_for_cache_sync(self) -> bool:
        return all(i.has_synced() for i in self._informers.values())

    def stop(self) -> None:
        for inf in self._informers.values():
            inf.stop()
