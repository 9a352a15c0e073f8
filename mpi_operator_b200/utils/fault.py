"""Fault injection for recovery tests.

The reference has no fault-injection subsystem: its tests "play kubelet" and write failed pod statuses
(pkg/controller/mpi_job_controller_test.go:667-679, test/integration/mpi_job_controller_test.go:610-655) or submit a
malformed command (test/e2e/mpi_job_test.go:92-100). With real rank processes on one box the equivalent is to make
a chosen rank die, exit or hang at a chosen point and watch the launcher / backoffLimit / elastic machinery react.

    B200MPI_FAULT="kill_rank:3@step:50"          rank 3 SIGKILLs itself when it reaches training step 50
    B200MPI_FAULT="exit_rank:1@step:10:code=7"   rank 1 calls os._exit(7)
    B200MPI_FAULT="hang_rank:2@step:5"           rank 2 stops making progress (collective watchdog / activeDeadline tests)
    B200MPI_FAULT="kill_rank:0@time:2.5"         wall-clock trigger, seconds after the injector was created; the native
                                                 mpirun honours the same spec from outside the rank (csrc/spawner/mpirun.cc)
    ...;once                                     fire on the first attempt only: a marker file next to the job's slot map
                                                 (or $B200MPI_FAULT_DIR) survives the launcher's OnFailure restart

Hook points: ``DataParallelTrainer.step``, ``hvd.DistributedOptimizer.step`` and ``hvd.elastic.State.commit`` call
``injector().on_step()``; scripts may call it directly."""
from __future__ import annotations

import hashlib
import os
import signal
import time
from dataclasses import dataclass
from typing import Optional

ENV = "B200MPI_FAULT"
ACTIONS = ("kill", "exit", "hang")


@dataclass(frozen=True)
class FaultSpec:
    action: str            # kill | exit | hang
    rank: int
    trigger: str           # step | time
    at: float
    code: int = 1
    once: bool = False
    raw: str = ""

    @staticmethod
    def parse(text: str) -> "FaultSpec":
        raw = text.strip()
        parts = [p.strip() for p in raw.split(";") if p.strip()]
        if not parts:
            raise ValueError("empty fault spec")
        once = False
        for flag in parts[1:]:
            if flag != "once":
                raise ValueError(f"unknown fault flag {flag!r} in {raw!r}")
            once = True
        try:
            what, when = parts[0].split("@", 1)
            act_s, rank_s = what.split(":", 1)
            if not act_s.endswith("_rank") or act_s[:-5] not in ACTIONS:
                raise ValueError(f"action must be one of {[a + '_rank' for a in ACTIONS]}")
            fields = when.split(":")
            trigger, at = fields[0], float(fields[1])
            if trigger not in ("step", "time") or at < 0:
                raise ValueError("trigger must be step:<n> or time:<seconds>")
            code = 1
            for extra in fields[2:]:
                k, v = extra.split("=", 1)
                if k != "code":
                    raise ValueError(f"unknown option {k!r}")
                code = int(v)
            return FaultSpec(act_s[:-5], int(rank_s), trigger, at, code, once, raw)
        except (ValueError, IndexError) as e:
            raise ValueError(f"bad {ENV} spec {raw!r}: {e}") from None


def _marker_path(spec: FaultSpec, env) -> str:
    base = env.get("B200MPI_FAULT_DIR")
    if not base:
        slots = env.get("B200MPI_SLOTS_FILE")
        base = os.path.dirname(slots) if slots else env.get("B200MPI_POD_DIR") or os.path.join(env.get("TMPDIR", "/tmp"), "b200mpi-fault")
    job = env.get("B200MPI_MPIJOB_NAME", "job")
    return os.path.join(base, f"{job}.fault-{hashlib.sha1(spec.raw.encode()).hexdigest()[:10]}.fired")


class FaultInjector:
    """Per-process injector. ``on_step()`` is cheap when no fault is configured or it targets another rank."""

    def __init__(self, spec: Optional[FaultSpec], rank: int, env=None):
        self.spec = spec if spec is not None and spec.rank == rank else None
        self.rank = rank
        self.steps = 0
        self._t0 = time.monotonic()
        self._env = dict(os.environ if env is None else env)
        self.fired = False

    @property
    def armed(self) -> bool:
        if self.spec is None or self.fired:
            return False
        return not (self.spec.once and os.path.exists(_marker_path(self.spec, self._env)))

    def on_step(self, step: Optional[int] = None) -> None:
        """Call once per training step (``step`` defaults to an internal counter starting at 1)."""
        self.steps = self.steps + 1 if step is None else step
        s = self.spec
        if s is None:
            return
        due = self.steps >= s.at if s.trigger == "step" else time.monotonic() - self._t0 >= s.at
        if due and self.armed:
            self.fire()

    def fire(self) -> None:
        s = self.spec
        self.fired = True
        if s.once:
            path = _marker_path(s, self._env)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as f:
                f.write(f"{s.raw} rank={self.rank} step={self.steps}\n")
        print(f"[b200mpi fault] rank {self.rank}: injecting {s.action} at step {self.steps} ({s.raw})", flush=True)
        if s.action == "kill":
            os.kill(os.getpid(), signal.SIGKILL)
        elif s.action == "exit":
            os._exit(s.code)
        else:  # hang: stop participating; the peers' collective watchdog or the job's activeDeadline must notice
            while True:
                time.sleep(3600)


_injector: Optional[FaultInjector] = None


def injector(rank: Optional[int] = None) -> FaultInjector:
    """Process-wide injector built from $B200MPI_FAULT (a disarmed one when unset)."""
    global _injector
    if _injector is None:
        text = os.environ.get(ENV, "")
        if rank is None:
            from ..launch.env import rank_info_from_env
            rank = rank_info_from_env().rank
        _injector = FaultInjector(FaultSpec.parse(text) if text.strip() else None, rank)
    return _injector


def reset() -> None:
    global _injector
    _injector = None
