"""Event recording (record.EventRecorder of client-go; the reference's user
facing audit trail).  Events are stored as ``v1.Event``
objects in the namespace of the involved object and mirrored to the log."""
from __future__ import annotations

import logging
from typing import List, Optional

from ..api import meta as M
from ..client.store import ObjectStore

log = logging.getLogger("mpi-job-controller")

EVENT_TYPE_NORMAL = "Normal"
EVENT_TYPE_WARNING = "Warning"


def _ref(obj) -> dict:
    d = obj.to_dict() if hasattr(obj, "to_dict") else obj
    md = d.get("metadata", {})
    return {"apiVersion": d.get("apiVersion", ""), "kind": d.get("kind", ""), "name": md.get("name", ""),
            "namespace": md.get("namespace", ""), "uid": md.get("uid", "")}


class EventRecorder:
    def __init__(self, store: Optional[ObjectStore], component: str = "mpi-job-controller", ttl_seconds: float = 3600.0):
        self.store, self.component = store, component
        self._seq = 0
        self._recent = {}
        self.ttl_seconds = ttl_seconds   # kube-apiserver --event-ttl default: 1 h
        self._since_prune = 0

    def prune(self, now: Optional[float] = None) -> int:
        """Delete events whose lastTimestamp is older than the TTL (the apiserver does this for the reference; without it
        a long-running daemon's store and journal grow with every job ever run). Returns the number removed."""
        if self.store is None or self.ttl_seconds <= 0:
            return 0
        import time
        cutoff = M.now_rfc3339((time.time() if now is None else now) - self.ttl_seconds)
        n = 0
        for ev in self.store.list("events"):
            if (ev.get("lastTimestamp") or ev.get("firstTimestamp") or "") < cutoff:
                try:
                    self.store.delete("events", M.namespace_of(ev), M.name_of(ev))
                    n += 1
                except Exception:  # noqa: BLE001 - already gone
                    pass
        return n

    def event(self, obj, etype: str, reason: str, message: str) -> None:
        ref = _ref(obj)
        log.info("Event(%s/%s): type: '%s' reason: '%s' %s", ref["namespace"], ref["name"], etype, reason, message)
        if self.store is None:
            return
        now = M.now_rfc3339()
        # event correlation (client-go EventCorrelator): identical events bump count/lastTimestamp
        agg_key = (ref["uid"], ref["name"], etype, reason, message)
        prev = self._recent.get(agg_key)
        if prev is not None:
            try:
                ev = self.store.get("events", prev[0], prev[1])
                ev["count"] = int(ev.get("count", 1)) + 1
                ev["lastTimestamp"] = now
                self.store.update("events", ev)
                return
            except Exception:  # noqa: BLE001 - fall through to a fresh event
                self._recent.pop(agg_key, None)
        self._seq += 1
        ev = {
            "apiVersion": "v1", "kind": "Event",
            "metadata": {"name": f"{ref['name']}.{self._seq:08x}{M.new_uid()[:4]}", "namespace": ref["namespace"] or "default"},
            "involvedObject": ref, "reason": reason, "message": message, "type": etype,
            "source": {"component": self.component}, "firstTimestamp": now, "lastTimestamp": now, "count": 1,
        }
        self._since_prune += 1
        if self._since_prune >= 256:
            self._since_prune = 0
            self.prune()
        try:
            self.store.create("events", ev)
            self._recent[agg_key] = (ev["metadata"]["namespace"], ev["metadata"]["name"])
            if len(self._recent) > 4096:
                self._recent.pop(next(iter(self._recent)))
        except Exception:  # events are best effort
            log.exception("could not record event")

    def eventf(self, obj, etype: str, reason: str, fmt: str, *args) -> None:
        self.event(obj, etype, reason, fmt % args if args else fmt)


class FakeRecorder(EventRecorder):
    """record.FakeRecorder: collects "<type> <reason> <message>" strings."""

    def __init__(self):
        super().__init__(None)
        self.events: List[str] = []

    def event(self, obj, etype: str, reason: str, message: str) -> None:
        self.events.append(f"{etype} {reason} {message}")
