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


def _read_world():
    """(generation, world size) the launcher last published for this job, or None. With B200MPI_ELASTIC_DIR set the native
    mpirun runs in elastic mode: it watches discover_hosts.sh itself, spawns the additional ranks of a larger world, waits
    until they are ready and only then writes `<dir>/world`; surviving ranks re-form the communicator IN PLACE."""
    d = os.environ.get("B200MPI_ELASTIC_DIR")
    if not d:
        return None
    try:
        with open(os.path.join(d, "world")) as f:
            g, w = f.read().split()[:2]
        return int(g), int(w)
    except (OSError, ValueError):
        return None


class State:
    """Horovod's ``ObjectState``: named picklable attributes that are saved on ``commit()``, rolled back by ``restore()`` and
    broadcast from rank 0 by ``sync()``; callbacks registered with ``register_reset_callbacks`` run after every (re)start of the
    world (``on_reset``), e.g. to rescale the learning rate to the new ``hvd.size()``."""

    def __init__(self, **kwargs):
        self._saved = {}
        self._reset_callbacks = []
        self._hosts = _discover_hosts()
        for k, v in kwargs.items():
            setattr(self, k, v)
        self._keys = list(kwargs)

    def register_reset_callbacks(self, callbacks) -> None:
        self._reset_callbacks.extend(callbacks)

    def on_reset(self) -> None:
        for cb in self._reset_callbacks:
            cb()

    def commit(self):
        from ..utils import fault
        fault.injector().on_step()
        self.save()
        self.check_host_updates()

    def check_host_updates(self):
        if os.environ.get("B200MPI_ELASTIC_DIR"):      # in-place mode: the launcher decides when the new world is ready
            # every rank must leave the old world at the SAME commit: agree on the newest generation anyone has seen
            from . import Max, _dev, allreduce
            gw = _read_world()
            seen = torch.tensor([float(gw[0] if gw else 0)], device=_dev())
            newest = int(allreduce(seen, op=Max, name="elastic.generation").item())
            if newest > int(os.environ.get("B200MPI_GENERATION", "0") or 0):
                raise HostsUpdatedInterrupt(f"world generation {newest}")
            return
        cur = _discover_hosts()
        if cur is not None and self._hosts is not None and cur != self._hosts:
            self._hosts = cur
            raise HostsUpdatedInterrupt("discover_hosts.sh changed")

    @staticmethod
    def _is_sampler(v) -> bool:
        return hasattr(v, "state_dict") and hasattr(v, "load_state_dict") and hasattr(v, "processed_indices")

    def save(self):
        """Snapshot by VALUE (Horovod deep-copies too): a later ``restore()`` must roll back to the commit even when the
        live objects were mutated in place since. Samplers are saved through their ``state_dict()`` (never the dataset)."""
        import copy
        self._saved = {}
        for k in self._keys:
            v = getattr(self, k)
            self._saved[k] = ("__sampler__", copy.deepcopy(v.state_dict())) if self._is_sampler(v) else copy.deepcopy(v)

    def restore(self):
        import copy
        for k, v in self._saved.items():
            if isinstance(v, tuple) and len(v) == 2 and v[0] == "__sampler__":
                cur = getattr(self, k, None)
                if cur is not None and self._is_sampler(cur):
                    cur.load_state_dict(copy.deepcopy(v[1]))   # the SAME sampler object the DataLoader holds
                continue
            setattr(self, k, copy.deepcopy(v))

    def sync(self):
        """Rank 0's values everywhere. Samplers are merged instead: every rank's processed indices are gathered and
        united (each rank consumed a different shard), loaded into the existing sampler object and re-partitioned over
        the current world by its ``reset()`` — rank, shard and dataset stay local."""
        from . import allgather_object, broadcast_object
        for k in self._keys:
            v = getattr(self, k)
            if self._is_sampler(v):
                parts = allgather_object({"epoch": v.epoch, "processed": sorted(v.processed_indices)})
                epoch = max(p["epoch"] for p in parts)
                done = set()
                for p in parts:
                    if p["epoch"] == epoch:
                        done.update(p["processed"])
                v.load_state_dict({"epoch": epoch, "processed_indices": sorted(done)})
                continue
            setattr(self, k, broadcast_object(v, 0))


ObjectState = State   # Horovod's name for the generic state


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
            import copy
            self._opt_sd = copy.deepcopy(self.optimizer.state_dict())   # state_dict() aliases the live momentum tensors
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
            import copy
            self.optimizer.load_state_dict(copy.deepcopy(self._opt_sd))   # load_state_dict adopts the tensors it is given

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
        state.on_reset()      # every incarnation is a reset of the world: new size, new rank
        while True:
            try:
                return func(state, *args, **kwargs)
            except HostsUpdatedInterrupt:
                state.save()
                gw = _read_world()
                if gw is None:    # restart mode: leave with the rescale code, the launcher re-runs mpirun on the new hostfile
                    raise SystemExit(int(os.environ.get("B200MPI_RESCALE_EXIT_CODE", "75")))
                # in-place mode: survivors keep their process, CUDA context and model; ranks beyond the new world retire
                import time as _time
                from . import _reinit, rank, shutdown
                generation, world = gw
                t0 = _time.time()
                if rank() >= world:
                    shutdown()
                    raise SystemExit(0)
                _reinit(world, generation)
                state.sync()          # rank 0 (always a survivor) brings the new ranks up to the committed state
                state.on_reset()
                if rank() == 0:
                    import sys
                    print(f"[elastic] re-formed in place: generation {generation}, world size {world}, {(_time.time() - t0) * 1e3:.0f} ms",
                          file=sys.stderr, flush=True)
            except HorovodInternalError as e:
                import sys
                print(f"[elastic] collective failed ({e}); rolling back to the last commit and leaving for a re-spawn", file=sys.stderr, flush=True)
                state.restore()
                raise SystemExit(int(os.environ.get("B200MPI_RESCALE_EXIT_CODE", "75")))
    return wrapper


class ElasticSampler(torch.utils.data.Sampler):
    """Horovod's ``hvd.elastic.ElasticSampler``: shards the dataset over the CURRENT world and remembers which indices of
    the epoch were already consumed, so that after a rescale the remaining samples are re-partitioned over the new world
    instead of being repeated or dropped. Register it in the ``TorchState`` (``TorchState(model, opt, sampler=sampler)``)
    and call ``record_batch`` after every step."""

    def __init__(self, dataset, shuffle: bool = True, seed: int = 0):
        self.dataset, self.shuffle, self.seed = dataset, shuffle, seed
        self.epoch = 0
        self.processed_indices = set()
        self.reset()

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch
        self.processed_indices = set()
        self.reset()

    def record_batch(self, batch_idx: int, batch_size: int) -> None:
        self.processed_indices.update(self.get_indices(batch_idx, batch_size))

    def get_indices(self, batch_idx: int, batch_size: int):
        start = batch_idx * batch_size
        return self.indices[start:min(start + batch_size, len(self.indices))]

    def state_dict(self) -> dict:
        return {"epoch": self.epoch, "processed_indices": sorted(self.processed_indices)}

    def load_state_dict(self, sd: dict) -> None:
        self.epoch = sd["epoch"]
        self.processed_indices = set(sd["processed_indices"])
        self.reset()

    def reset(self) -> None:
        from . import is_initialized, rank, size
        self.num_replicas, self.rank = (size(), rank()) if is_initialized() else (1, 0)
        remaining = [i for i in range(len(self.dataset)) if i not in self.processed_indices]
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            remaining = [remaining[i] for i in torch.randperm(len(remaining), generator=g).tolist()]
        self.remaining_indices = remaining
        self.num_samples = (len(remaining) + self.num_replicas - 1) // self.num_replicas
        total = self.num_samples * self.num_replicas
        padded = remaining + remaining[:total - len(remaining)] if remaining else []
        self.indices = padded[self.rank:total:self.num_replicas]

    def __iter__(self):
        return iter(self.indices)

    def __len__(self):
        return self.num_samples
