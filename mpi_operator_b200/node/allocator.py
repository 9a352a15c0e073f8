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

    def _select(self, n: int) -> Optional[List[int]]:
        """Which free GPUs a reservation of ``n`` gets. Every pair is an NVSwitch peer, so the fabric does not care; the host side
        does: GPUs 0-3 / 4-7 of an HGX board hang off different sockets. A reservation that fits on one socket is placed on the
        socket with the FEWEST free GPUs that still holds it (best fit: a later 4-GPU job still finds a whole socket, and the
        ranks' pinned staging buffers stay local with ``mpirun --bind-to numa``); one that does not fit takes the fullest sockets
        first. Without NUMA information this is "the first n free GPUs"."""
        if n > len(self._free):
            return None
        node_of = {g.index: g.numa_node for g in self.topology.gpus}
        by_node: Dict[object, List[int]] = {}
        for g in self._free:
            by_node.setdefault(node_of.get(g), []).append(g)
        order = lambda k: (k is None, k if k is not None else 0)   # noqa: E731  (None sorts last)
        fits = [k for k, v in by_node.items() if len(v) >= n]
        if fits:
            node = min(fits, key=lambda k: (len(by_node[k]), order(k)))
            return by_node[node][:n]
        chosen: List[int] = []
        for k in sorted(by_node, key=lambda k: (-len(by_node[k]), order(k))):
            chosen.extend(by_node[k][:n - len(chosen)])
            if len(chosen) == n:
                break
        return sorted(chosen)

    def _take(self, n: int, pool: Optional[List[int]] = None) -> Optional[List[int]]:
        """Removes ``n`` GPUs from the free list: the first ``n`` of ``pool`` (a gang's pre-selected set) or ``_select(n)``."""
        got = pool[:n] if pool is not None else self._select(n)
        if got is None or len(got) < n:
            return None
        if pool is not None:
            del pool[:n]
        taken = set(got)
        self._free = [g for g in self._free if g not in taken]
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
            pool = self._select(need)          # one placement decision for the whole gang, dealt out in pod order
            for r in sorted(pending, key=lambda r: r.key):
                self._held[r.key] = self._take(r.gpus, pool)
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
