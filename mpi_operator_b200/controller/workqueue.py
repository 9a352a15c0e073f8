"""Rate-limited work queue (client-go util/workqueue semantics).

The reference builds ``NewTypedRateLimitingQueue(MaxOf(ItemExponentialFailure
(5ms, 1000s), BucketRateLimiter(qps=ControllerRateLimit, burst=
ControllerBurst)))`` (pkg/controller/mpi_job_controller.go:123-124,348-354).
Same guarantees here: an item is processed by at most one worker at a time,
re-adds while processing are deferred until ``done``, duplicates collapse.
"""
from __future__ import annotations

import heapq
import threading
import time
from typing import Dict, Hashable, List, Optional, Set, Tuple


class ItemExponentialFailureRateLimiter:
    def __init__(self, base_delay: float = 0.005, max_delay: float = 1000.0):
        self.base, self.max = base_delay, max_delay
        self._failures: Dict[Hashable, int] = {}
        self._lock = threading.Lock()

    def when(self, item) -> float:
        with self._lock:
            n = self._failures.get(item, 0)
            self._failures[item] = n + 1
        d = self.base * (2 ** n) if n < 64 else self.max
        return min(d, self.max)

    def forget(self, item) -> None:
        with self._lock:
            self._failures.pop(item, None)

    def num_requeues(self, item) -> int:
        with self._lock:
            return self._failures.get(item, 0)


class BucketRateLimiter:
    """Token bucket: ``qps`` refill, ``burst`` capacity; reservations queue up."""

    def __init__(self, qps: float = 10.0, burst: int = 100, now=time.monotonic):
        self.qps, self.burst, self._now = float(qps), float(burst), now
        self._tokens = float(burst)
        self._last = now()
        self._lock = threading.Lock()

    def when(self, item) -> float:
        with self._lock:
            t = self._now()
            self._tokens = min(self.burst, self._tokens + (t - self._last) * self.qps)
            self._last = t
            self._tokens -= 1.0
            return 0.0 if self._tokens >= 0 else -self._tokens / self.qps

    def forget(self, item) -> None:
        pass

    def num_requeues(self, item) -> int:
        return 0


class MaxOfRateLimiter:
    def __init__(self, *limiters):
        self.limiters = limiters

    def when(self, item) -> float:
        return max(l.when(item) for l in self.limiters)

    def forget(self, item) -> None:
        for l in self.limiters:
            l.forget(item)

    def num_requeues(self, item) -> int:
        return max(l.num_requeues(item) for l in self.limiters)


def default_controller_rate_limiter(qps: float = 10.0, burst: int = 100) -> MaxOfRateLimiter:
    return MaxOfRateLimiter(ItemExponentialFailureRateLimiter(0.005, 1000.0), BucketRateLimiter(qps, burst))


class RateLimitingQueue:
    def __init__(self, rate_limiter=None, name: str = "MPIJob"):
        self.name = name
        self.rl = rate_limiter or default_controller_rate_limiter()
        self._cv = threading.Condition()
        self._queue: List[Hashable] = []
        self._dirty: Set[Hashable] = set()
        self._processing: Set[Hashable] = set()
        self._delayed: List[Tuple[float, int, Hashable]] = []
        self._seq = 0
        self._shutdown = False
        self._timer = threading.Thread(target=self._delay_loop, name=f"{name}-delay", daemon=True)
        self._timer.start()

    # -- plain queue ---------------------------------------------------------
    def add(self, item) -> None:
        with self._cv:
            if self._shutdown or item in self._dirty:
                return
            self._dirty.add(item)
            if item in self._processing:
                return
            self._queue.append(item)
            self._cv.notify()

    def get(self, timeout: Optional[float] = None):
        """Returns (item, shutdown)."""
        with self._cv:
            end = None if timeout is None else time.monotonic() + timeout
            while not self._queue and not self._shutdown:
                rem = None if end is None else end - time.monotonic()
                if rem is not None and rem <= 0:
                    return None, False
                self._cv.wait(rem)
            if not self._queue:
                return None, True
            item = self._queue.pop(0)
            self._processing.add(item)
            self._dirty.discard(item)
            return item, False

    def done(self, item) -> None:
        with self._cv:
            self._processing.discard(item)
            if item in self._dirty:
                self._queue.append(item)
                self._cv.notify()

    def __len__(self) -> int:
        with self._cv:
            return len(self._queue)

    def shut_down(self) -> None:
        with self._cv:
            self._shutdown = True
            self._cv.notify_all()

    # -- delaying / rate limiting -------------------------------------------
    def add_after(self, item, delay: float) -> None:
        if delay <= 0:
            self.add(item)
            return
        with self._cv:
            if self._shutdown:
                return
            self._seq += 1
            heapq.heappush(self._delayed, (time.monotonic() + delay, self._seq, item))
            self._cv.notify_all()

    def add_rate_limited(self, item) -> None:
        self.add_after(item, self.rl.when(item))

    def forget(self, item) -> None:
        self.rl.forget(item)

    def num_requeues(self, item) -> int:
        return self.rl.num_requeues(item)

    def _delay_loop(self) -> None:
        while True:
            with self._cv:
                if self._shutdown:
                    return
                now = time.monotonic()
                ready = []
                while self._delayed and self._delayed[0][0] <= now:
                    ready.append(heapq.heappop(self._delayed)[2])
                wait = (self._delayed[0][0] - now) if self._delayed else 0.5
            for it in ready:
                self.add(it)
            with self._cv:
                if self._shutdown:
                    return
                self._cv.wait(min(wait, 0.5) if not ready else 0)
