"""All-or-nothing GPU slot allocation: the local analogue of gang scheduling.

Reference behaviour being replaced: Volcano / scheduler-plugins co-scheduling
driven by the PodGroup the controller creates (pkg/controller/podgroup.go;
SURVEY.md §5.8): a group starts only when ``minMember`` pods (and
``minResources``) fit, ``queue``/priority order the pending list and
``scheduleTimeoutSeconds`` bounds the wait.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from .topology import Topology


@dataclass
class SlotRequest:
    key: str                 # pod key "ns/name"
    gpus: int = 0            # nvidia.com/gpu of the pod
    group: str = ""          # pod group key ("" = schedule individually)
    priority: int = 0
    queue: str = ""
    created: float = field(default_factory=time.time)


class GangAllocator:
    def __init__(self, topology: Topology):
        self.topology = topology
        self._lock = threading.RLock()
        self._free: List[int] = [g.index for g in topology.gpus]
        self._held: Dict[str, List[int]] = {}

    @property
    def free_gpus(self) -> int:
        with self._lock:
            return len(self._free)

    def held(self, key: str) -> Optional[List[int]]:
        with self._lock:
            return list(self._held[key]) if key in self._held else None

    def _take(self, n: int) -> Optional[List[int]]:
        if n > len(self._free):
            return None
        got, self._free = self._free[:n], self._free[n:]
        return got

    def allocate(self, req: SlotRequest) -> Optional[List[int]]:
        """Single pod, no gang."""
        with self._lock:
            if req.key in self._held:
                return list(self._held[req.key])
            got = self._take(req.gpus)
            if got is None:
                return None
            self._held[req.key] = got
            return list(got)

    def allocate_gang(self, reqs: List[SlotRequest], min_member: int, min_gpus: int = 0) -> Optional[Dict[str, List[int]]]:
        """Grant every request of the group or none.

        ``min_member``: the group only starts once at least that many member pods
        exist.  ``min_gpus``: PodGroup minResources["nvidia.com/gpu"] (0 = sum of
        the requests).
        """
        with self._lock:
            pending = [r for r in reqs if r.key not in self._held]
            if not pending:
                return {r.key: list(self._held[r.key]) for r in reqs}
            already = len(reqs) - len(pending)
            if already == 0 and len(reqs) < min_member:
                return None
            need = sum(r.gpus for r in pending)
            if already == 0 and min_gpus > need:
                # the group reserves at least min_gpus worth of capacity before it may start
                if min_gpus > len(self._free):
                    return None
            if need > len(self._free):
                return None
            out = {}
            for r in sorted(pending, key=lambda r: r.key):
                self._held[r.key] = self._take(r.gpus)
            for r in reqs:
                out[r.key] = list(self._held[r.key])
            return out

    def adopt(self, key: str, gpus: List[int]) -> None:
        """Re-register a reservation that predates this allocator (daemon restart)."""
        with self._lock:
            self._held[key] = list(gpus)
            self._free = [g for g in self._free if g not in gpus]

    def release(self, key: str) -> None:
        with self._lock:
            got = self._held.pop(key, None)
            if got:
                self._free = sorted(set(self._free) | set(got))
