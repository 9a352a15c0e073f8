"""hvd.elastic-style state for rescale (reference: proposals/elastic-horovod.md:13-31;
SURVEY.md §3.4, §5.4).  The daemon re-spawns ranks when Worker.replicas changes;
``TorchState.commit()/restore()/sync()`` keep model + optimizer state across a
re-formed communicator: rank 0 checkpoints on ``commit`` (rank-0-only convention,
tensorflow_mnist.py:159) and the state is re-broadcast (K3) on ``sync``."""
from __future__ import annotations

import functools
import os
from typing import Callable, Optional

import torch

from .exceptions import HorovodInternalError, HostsUpdatedInterrupt  # noqa: F401


class WorkersAvailableException(RuntimeError):
    pass


def _discover_hosts() -> Optional[str]:
    root = os.environ.get("B200MPI_POD_ROOTFS", "")
    for path in (os.path.join(root, "etc/mpi/discover_hosts.sh"), "/etc/mpi/discover_hosts.sh"):
        if os.path.exists(path):
            with open(path) as f:
                return f.read()
    return None


class State:
    def __init__(self, **kwargs):
        self._saved = {}
        self._hosts = _discover_hosts()
        for k, v in kwargs.items():
            setattr(self, k, v)
        self._keys = list(kwargs)

    def commit(self):
        from ..utils import fault
        fault.injector().on_step()
        self.save()
        self.check_host_updates()

    def check_host_updates(self):
        cur = _discover_hosts()
        if cur is not None and self._hosts is not None and cur != self._hosts:
            self._hosts = cur
            raise HostsUpdatedInterrupt("discover_hosts.sh changed")

    def save(self):
        self._saved = {k: getattr(self, k) for k in self._keys}

    def restore(self):
        for k, v in self._saved.items():
            setattr(self, k, v)

    def sync(self):
        from . import broadcast_object
        for k in self._keys:
            setattr(self, k, broadcast_object(getattr(self, k), 0))


class TorchState(State):
    def __init__(self, model=None, optimizer=None, checkpoint_path: Optional[str] = None, **kwargs):
        super().__init__(**kwargs)
        self.model, self.optimizer = model, optimizer
        self.checkpoint_path = checkpoint_path or os.environ.get("B200MPI_ELASTIC_CHECKPOINT")
        self._model_sd = self._opt_sd = None

    def save(self):
        super().save()
        from . import rank
        if self.model is not None:
            self._model_sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        if self.optimizer is not None:
            self._opt_sd = self.optimizer.state_dict()
        if self.checkpoint_path and rank() == 0:  # rank-0-only checkpoint
            tmp = self.checkpoint_path + ".tmp"
            torch.save({"model": self._model_sd, "optimizer": self._opt_sd, "extra": self._saved}, tmp)
            os.replace(tmp, self.checkpoint_path)

    def restore(self):
        super().restore()
        if self._model_sd is None and self.checkpoint_path and os.path.exists(self.checkpoint_path):
            ck = torch.load(self.checkpoint_path, map_location="cuda" if torch.cuda.is_available() else "cpu", weights_only=False)
            self._model_sd, self._opt_sd, self._saved = ck["model"], ck["optimizer"], ck.get("extra", {})
            super().restore()
        if self.model is not None and self._model_sd is not None:
            self.model.load_state_dict(self._model_sd)
        if self.optimizer is not None and self._opt_sd is not None:
            self.optimizer.load_state_dict(self._opt_sd)

    def sync(self):
        from . import broadcast_optimizer_state, broadcast_parameters
        super().sync()
        if self.model is not None:
            broadcast_parameters(self.model.state_dict(), root_rank=0)
        if self.optimizer is not None:
            broadcast_optimizer_state(self.optimizer, root_rank=0)


def run(func: Callable) -> Callable:
    """Decorator: restore + sync state, run; on HostsUpdatedInterrupt exit with the
    'rescale' code so the launcher re-spawns the new world (daemon-side elasticity).
    A HorovodInternalError (a peer died, the engine was shut down or stalled out) rolls the state back to the last
    commit and ends this incarnation the same way: the re-spawned world resumes from the committed checkpoint."""
    @functools.wraps(func)
    def wrapper(state, *args, **kwargs):
        state.restore()
        state.sync()
        try:
            return func(state, *args, **kwargs)
        except HostsUpdatedInterrupt:
            state.save()
            raise SystemExit(int(os.environ.get("B200MPI_RESCALE_EXIT_CODE", "75")))
        except HorovodInternalError as e:
            import sys
            print(f"[elastic] collective failed ({e}); rolling back to the last commit and leaving for a re-spawn", file=sys.stderr, flush=True)
            state.restore()
            raise SystemExit(int(os.environ.get("B200MPI_RESCALE_EXIT_CODE", "75")))
    return wrapper
