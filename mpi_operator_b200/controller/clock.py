"""Injectable clock (k8s.io/utils/clock; used at controller.go:263-264,568-571,1099)."""
from __future__ import annotations

import threading
import time


class RealClock:
    def now(self) -> float:
        return time.time()

    def since(self, t: float) -> float:
        return time.time() - t

    def sleep(self, d: float) -> None:
        time.sleep(d)


class FakeClock(RealClock):
    """clocktesting.FakeClock: time only moves when the test says so."""

    def __init__(self, start: float = 1_600_000_000.0):
        self._t = start
        self._cv = threading.Condition()

    def now(self) -> float:
        return self._t

    def since(self, t: float) -> float:
        return self._t - t

    def step(self, d: float) -> None:
        with self._cv:
            self._t += d
            self._cv.notify_all()

    def set_time(self, t: float) -> None:
        with self._cv:
            self._t = t
            self._cv.notify_all()

    def sleep(self, d: float) -> None:
        self.step(d)
