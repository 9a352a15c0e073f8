"""In-memory object store: the single-box stand-in for kube-apiserver + etcd.

The reference's controller is written against the Kubernetes API machinery
(resourceVersion, ownerReferences + garbage collection, label selectors, watch
streams, status sub-resource).  This store reproduces those
semantics locally so the reconcile algorithm (pkg/controller/
mpi_job_controller.go:567-735) ports 1:1 in behaviour and its "test plays
kubelet" strategy works unchanged.  Optional JSON persistence
gives the daemon restart-adoption (SURVEY.md §5.4 [NEW]).
"""
from __future__ import annotations

import copy
import json
import os
import threading
from typing import Any, Callable, Dict, List, Optional, Tuple

from ..api import meta as M
from . import errors

ADDED, MODIFIED, DELETED = "ADDED", "MODIFIED", "DELETED"

# resource -> (apiVersion, kind, namespaced)
RESOURCES: Dict[str, Tuple[str, str, bool]] = {
    "mpijobs": ("kubeflow.org/v2beta1", "MPIJob", True),
    "pods": ("v1", "Pod", True),
    "services": ("v1", "Service", True),
    "configmaps": ("v1", "ConfigMap", True),
    "secrets": ("v1", "Secret", True),
    "events": ("v1", "Event", True),
    "jobs": ("batch/v1", "Job", True),
    "leases": ("coordination.k8s.io/v1", "Lease", True),
    "priorityclasses": ("scheduling.k8s.io/v1", "PriorityClass", False),
    "volcano-podgroups": ("scheduling.volcano.sh/v1beta1", "PodGroup", True),
    "sched-podgroups": ("scheduling.x-k8s.io/v1alpha1", "PodGroup", True),
    "queues": ("scheduling.volcano.sh/v1beta1", "Queue", False),
    "customresourcedefinitions": ("apiextensions.k8s.io/v1", "CustomResourceDefinition", False),
}

# resources whose status is a sub-resource: update() keeps status, update_status() keeps the rest
STATUS_SUBRESOURCE = {"mpijobs", "jobs", "pods", "volcano-podgroups", "sched-podgroups"}

WatchHandler = Callable[[str, dict, Optional[dict]], None]  # (event_type, obj, old_obj)


class ObjectStore:
    def __init__(self, persist_path: Optional[str] = None):
        self._lock = threading.RLock()
        self._objs: Dict[str, Dict[str, dict]] = {r: {} for r in RESOURCES}
        self._rv = 0
        self._watchers: Dict[str, List[WatchHandler]] = {r: [] for r in RESOURCES}
        self._persist_path = persist_path
        self._pending: List[Tuple[str, str, dict, Optional[dict]]] = []
        self._dispatch_lock = threading.RLock()
        self._save_lock = threading.Lock()
        self._saved_rv = -1
        # Persistence = snapshot file + append-only journal of mutations (O(1) work per write instead of re-serialising
        # the whole store); the journal is folded into a fresh snapshot every kCompactEvery records and at load time.
        self._journal: List[str] = []
        self._journal_records = 0
        self._wal = None
        if persist_path and (os.path.exists(persist_path) or os.path.exists(persist_path + ".wal")):
            self._load()

    # ------------------------------------------------------------ helpers --
    @staticmethod
    def _key(ns: str, name: str) -> str:
        return f"{ns}/{name}" if ns else name

    def _check(self, resource: str) -> None:
        if resource not in RESOURCES:
            raise errors.ApiError("NotFound", f"the server could not find the requested resource ({resource})", 404)

    def _next_rv(self) -> str:
        self._rv += 1
        return str(self._rv)

    def _emit(self, resource: str, etype: str, obj: dict, old: Optional[dict]) -> None:
        # Handlers run outside the mutation that produced the event and never re-entrantly.
        self._pending.append((resource, etype, copy.deepcopy(obj), copy.deepcopy(old) if old else None))
        if self._persist_path:
            namespaced = RESOURCES[resource][2]
            key = self._key(M.namespace_of(obj) if namespaced else "", M.name_of(obj))
            rec = {"rv": self._rv, "r": resource, "k": key}
            if etype != DELETED:
                rec["o"] = obj
            self._journal.append(json.dumps(rec))

    def _flush(self) -> None:
        with self._dispatch_lock:  # ordered delivery; re-entrant for handlers that mutate the store
            while True:
                with self._lock:
                    if not self._pending:
                        break
                    resource, etype, obj, old = self._pending.pop(0)
                    handlers = list(self._watchers[resource])
                for h in handlers:
                    h(etype, obj, old)
        self._save()

    @staticmethod
    def _check_structure(resource: str, obj: dict) -> None:
        """The CRD's structural schema, applied on every write of the main resource like kube-apiserver does (types and enums
        only: api/schema.py); a mistyped field never reaches the controller."""
        from ..api.schema import core_structural_errors, structural_errors
        errs = structural_errors(obj) if resource == "mpijobs" else core_structural_errors(resource, obj)
        if errs:
            shown = errs[0] if len(errs) == 1 else "[" + ", ".join(errs[:8]) + (", ..." if len(errs) > 8 else "") + "]"
            md = obj.get("metadata")
            name = md.get("name") if isinstance(md, dict) else ""
            raise errors.invalid("MPIJob.kubeflow.org" if resource == "mpijobs" else resource, str(name or ""), shown)

    # ---------------------------------------------------------------- CRUD --
    def create(self, resource: str, obj: dict) -> dict:
        self._check(resource)
        self._check_structure(resource, obj)
        with self._lock:
            api_version, kind, namespaced = RESOURCES[resource]
            obj = copy.deepcopy(obj)
            md = M.meta(obj)
            if not md.get("name"):
                if md.get("generateName"):
                    md["name"] = md["generateName"] + M.new_uid()[:5]
                else:
                    raise errors.invalid(resource, "", "metadata.name: Required value")
            ns = md.get("namespace", "") if namespaced else ""
            if namespaced and not ns:
                ns = md["namespace"] = "default"
            problem = M.name_problem(md["name"]) or (M.name_problem(ns, "metadata.namespace") if namespaced else None)
            if problem:
                raise errors.invalid(resource, str(md["name"]), problem)
            key = self._key(ns, md["name"])
            if key in self._objs[resource]:
                raise errors.already_exists(resource, md["name"])
            obj.setdefault("apiVersion", api_version)
            obj.setdefault("kind", kind)
            md["uid"] = md.get("uid") or M.new_uid()
            md["resourceVersion"] = self._next_rv()
            md.setdefault("creationTimestamp", M.now_rfc3339())
            md.setdefault("generation", 1)
            self._objs[resource][key] = obj
            self._emit(resource, ADDED, obj, None)
            out = copy.deepcopy(obj)
        self._flush()
        return out

    def get(self, resource: str, namespace: str, name: str) -> dict:
        self._check(resource)
        with self._lock:
            o = self._objs[resource].get(self._key(namespace if RESOURCES[resource][2] else "", name))
            if o is None:
                raise errors.not_found(resource, name)
            return copy.deepcopy(o)

    def list(self, resource: str, namespace: Optional[str] = None, label_selector: Optional[Dict[str, str]] = None) -> List[dict]:
        self._check(resource)
        with self._lock:
            out = []
            for o in self._objs[resource].values():
                if namespace and RESOURCES[resource][2] and M.namespace_of(o) != namespace:
                    continue
                if label_selector and not M.label_selector_matches(label_selector, M.meta(o).get("labels")):
                    continue
                out.append(copy.deepcopy(o))
            out.sort(key=lambda o: (M.namespace_of(o), M.name_of(o)))
            return out

    def _update(self, resource: str, obj: dict, status_only: bool) -> dict:
        self._check(resource)
        if not status_only:
            self._check_structure(resource, obj)
        with self._lock:
            md = M.meta(obj)
            ns = md.get("namespace", "") if RESOURCES[resource][2] else ""
            key = self._key(ns, md.get("name", ""))
            cur = self._objs[resource].get(key)
            if cur is None:
                raise errors.not_found(resource, md.get("name", ""))
            rv = md.get("resourceVersion")
            if rv and rv != M.meta(cur)["resourceVersion"]:
                raise errors.conflict(resource, md.get("name", ""))
            old = cur
            new = copy.deepcopy(cur)
            if status_only:
                new["status"] = copy.deepcopy(obj.get("status", {}))
            else:
                keep_status = cur.get("status")
                new = copy.deepcopy(obj)
                nmd = M.meta(new)
                for k in ("uid", "creationTimestamp"):
                    nmd[k] = M.meta(cur).get(k)
                if resource in STATUS_SUBRESOURCE and keep_status is not None:
                    new["status"] = copy.deepcopy(keep_status)  # e.g. CRD subresources.status of MPIJob
                if new.get("spec") != cur.get("spec"):
                    nmd["generation"] = int(M.meta(cur).get("generation", 1)) + 1
                else:
                    nmd["generation"] = M.meta(cur).get("generation", 1)
            if new == cur:
                return copy.deepcopy(cur)  # no-op update: no new resourceVersion, no event
            M.meta(new)["resourceVersion"] = self._next_rv()
            self._objs[resource][key] = new
            self._emit(resource, MODIFIED, new, old)
            out = copy.deepcopy(new)
        self._flush()
        return out

    def update(self, resource: str, obj: dict) -> dict:
        return self._update(resource, obj, status_only=False)

    def update_status(self, resource: str, obj: dict) -> dict:
        return self._update(resource, obj, status_only=True)

    def patch(self, resource: str, namespace: str, name: str, patch: dict, status: bool = False) -> dict:
        """JSON merge patch (RFC 7386)."""
        with self._lock:
            cur = self.get(resource, namespace, name)
            merged = _merge_patch(cur, patch)
            M.meta(merged)["resourceVersion"] = M.meta(cur)["resourceVersion"]
        return self.update_status(resource, merged) if status else self.update(resource, merged)

    def delete(self, resource: str, namespace: str, name: str) -> dict:
        self._check(resource)
        with self._lock:
            key = self._key(namespace if RESOURCES[resource][2] else "", name)
            cur = self._objs[resource].pop(key, None)
            if cur is None:
                raise errors.not_found(resource, name)
            self._next_rv()
            self._emit(resource, DELETED, cur, None)
            self._collect_garbage(M.meta(cur).get("uid"))
            out = copy.deepcopy(cur)
        self._flush()
        return out

    def delete_collection(self, resource: str, namespace: Optional[str] = None, label_selector=None) -> int:
        n = 0
        for o in self.list(resource, namespace, label_selector):
            try:
                self.delete(resource, M.namespace_of(o), M.name_of(o))
                n += 1
            except errors.ApiError:
                pass
        return n

    def _collect_garbage(self, owner_uid: Optional[str]) -> None:
        """ownerReferences cascade (what kube-controller-manager's GC does for the reference)."""
        if not owner_uid:
            return
        for resource, objs in self._objs.items():
            for key in [k for k, o in objs.items()
                        if any(r.get("uid") == owner_uid for r in M.meta(o).get("ownerReferences", []) or [])]:
                dead = objs.pop(key)
                self._emit(resource, DELETED, dead, None)
                self._collect_garbage(M.meta(dead).get("uid"))

    # --------------------------------------------------------------- watch --
    def watch(self, resource: str, handler: WatchHandler, replay: bool = True) -> Callable[[], None]:
        self._check(resource)
        with self._lock:
            self._watchers[resource].append(handler)
            existing = [copy.deepcopy(o) for o in self._objs[resource].values()] if replay else []
        for o in existing:
            handler(ADDED, o, None)

        def cancel():
            with self._lock:
                if handler in self._watchers[resource]:
                    self._watchers[resource].remove(handler)
        return cancel

    # --------------------------------------------------------- persistence --
    kCompactEvery = 4000

    def _save(self) -> None:
        """Append the pending journal records (or compact). One writer at a time; a thread that finds the writer busy
        leaves its records to it — the writer re-checks after releasing, so nothing stays behind."""
        if not self._persist_path:
            return
        while True:
            if not self._save_lock.acquire(blocking=False):
                return
            try:
                with self._lock:
                    lines, self._journal = self._journal, []
                    compact = self._journal_records + len(lines) >= self.kCompactEvery
                    snap = None
                    if compact:
                        snap = {"rv": self._rv, "objects": {r: list(o.values()) for r, o in self._objs.items() if o}}
                try:
                    if compact:
                        self._write_snapshot(snap)
                    elif lines:
                        if self._wal is None:   # kept open: one write + flush per batch of records, no open/close
                            self._wal = open(self._persist_path + ".wal", "a")
                        self._wal.write("\n".join(lines) + "\n")
                        self._wal.flush()
                        self._journal_records += len(lines)
                except (OSError, ValueError):
                    self._wal = None  # state dir removed underneath us (shutdown): persistence is best effort
            finally:
                self._save_lock.release()
            if not self._journal:
                return

    def close(self) -> None:
        """Flush pending journal records and release the journal file."""
        self._save()
        with self._save_lock:
            if self._wal is not None:
                try:
                    self._wal.close()
                except OSError:
                    pass
                self._wal = None

    def _write_snapshot(self, snap: dict) -> None:
        tmp = f"{self._persist_path}.tmp{os.getpid()}"
        with open(tmp, "w") as f:
            json.dump(snap, f)
        os.replace(tmp, self._persist_path)
        if self._wal is not None:
            try:
                self._wal.close()
            except OSError:
                pass
            self._wal = None
        try:
            os.unlink(self._persist_path + ".wal")   # records up to snap["rv"] are in the snapshot now
        except OSError:
            pass
        self._journal_records = 0

    def _load(self) -> None:
        snap = {"rv": 0, "objects": {}}
        if os.path.exists(self._persist_path):
            with open(self._persist_path) as f:
                snap = json.load(f)
        self._rv = int(snap.get("rv", 0))
        for r, objs in snap.get("objects", {}).items():
            if r in self._objs:
                for o in objs:
                    ns = M.namespace_of(o) if RESOURCES[r][2] else ""
                    self._objs[r][self._key(ns, M.name_of(o))] = o
        base_rv, replayed = self._rv, 0
        try:
            with open(self._persist_path + ".wal") as f:
                for line in f:
                    try:
                        rec = json.loads(line)
                    except ValueError:
                        break  # torn last record of a crashed writer
                    if int(rec.get("rv", 0)) <= base_rv or rec.get("r") not in self._objs:
                        continue  # already contained in the snapshot
                    if "o" in rec:
                        self._objs[rec["r"]][rec["k"]] = rec["o"]
                    else:
                        self._objs[rec["r"]].pop(rec["k"], None)
                    self._rv = max(self._rv, int(rec["rv"]))
                    replayed += 1
        except OSError:
            pass
        if replayed:  # start the new process from a clean snapshot
            try:
                self._write_snapshot({"rv": self._rv, "objects": {r: list(o.values()) for r, o in self._objs.items() if o}})
            except OSError:
                pass


def _merge_patch(target: Any, patch: Any) -> Any:
    if not isinstance(patch, dict):
        return copy.deepcopy(patch)
    if not isinstance(target, dict):
        target = {}
    out = copy.deepcopy(target)
    for k, v in patch.items():
        if v is None:
            out.pop(k, None)
        else:
            out[k] = _merge_patch(out.get(k), v)
    return out
