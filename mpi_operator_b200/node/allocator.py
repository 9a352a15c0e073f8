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
        self._cordoned: Dict[int, str] = {}   # GPU index -> reason (`kubectl cordon` for a GPU; set by hand or by node/health.py)

    @property
    def free_gpus(self) -> int:
        with self._lock:
            return len(self._free)

    # ---- cordon: a GPU that must not receive new ranks (operator decision or failed health probe). Reservations that already
    #      hold it keep it until they end; it does not return to the free pool while cordoned.
    def cordon(self, gpu: int, reason: str = "cordoned") -> bool:
        with self._lock:
            if gpu not in {g.index for g in self.topology.gpus}:
                raise ValueError(f"no GPU {gpu} on this box")
            changed = self._cordoned.get(gpu) != reason
            self._cordoned[gpu] = reason
            self._free = [g for g in self._free if g != gpu]
            return changed

    def uncordon(self, gpu: int) -> bool:
        with self._lock:
            if gpu not in self._cordoned:
                return False
            del self._cordoned[gpu]
            if not any(gpu in held for held in self._held.values()):
                self._free = sorted(set(self._free) | {gpu})
            return True

    @property
    def cordoned(self) -> Dict[int, str]:
        with self._lock:
            return dict(self._cordoned)

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
                self._free = sorted(set(self._free) | {g for g in got if g not in self._cordoned})
