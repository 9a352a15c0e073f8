"""hvd.DistributedOptimizer: gradient averaging wrapped around any torch optimizer
(reference call site: examples/v2beta1/horovod/tensorflow_mnist.py:133)."""
from __future__ import annotations

import os
from typing import List, Optional

import torch


def _align(n, a):
    return (n + a - 1) // a * a


class _DistributedOptimizer:
    """Gradients are re-homed into one symmetric window (per dtype fp32), bucketed in
    reverse parameter order; each bucket is averaged in place by one b200mpi kernel
    launched from the autograd hook of its last gradient, on a high-priority stream;
    ``step()`` joins that stream and runs the wrapped optimizer."""

    def __init__(self, optimizer, named_parameters=None, compression=None, backward_passes_per_step: int = 1, op="average",
                 gradient_predivide_factor: float = 1.0, bucket_bytes: Optional[int] = None):
        from . import _comm, _op_name
        self._opt = optimizer
        self._comm = _comm()
        self._op = _op_name(op, None)
        self._passes = backward_passes_per_step
        self._pass = 0
        params = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
        if any(p.dtype != torch.float32 for p in params):
            raise ValueError("DistributedOptimizer expects fp32 parameters (use autocast for low-precision compute)")
        self._params = list(reversed(params))
        cap = (bucket_bytes or int(os.environ.get("B200MPI_BUCKET_BYTES", 32 << 20))) // 4
        self._buckets: List[dict] = []
        cur = {"start": 0, "numel": 0, "params": []}
        self._slot = {}
        for p in self._params:
            n = _align(p.numel(), 4)
            if cur["params"] and cur["numel"] + n > cap:
                cur["numel"] = _align(cur["numel"], 8)
                self._buckets.append(cur)
                cur = {"start": cur["start"] + cur["numel"], "numel": 0, "params": []}
            self._slot[p] = (len(self._buckets), cur["start"] + cur["numel"])
            cur["params"].append(p)
            cur["numel"] += n
        cur["numel"] = _align(cur["numel"], 8)
        self._buckets.append(cur)
        total = cur["start"] + cur["numel"]
        self._win = self._comm.alloc_window(total * 4)
        self._flat = self._win.tensor(torch.float32, numel=total)
        self._flat.zero_()
        for p in params:
            _, start = self._slot[p]
            p.grad = self._flat[start:start + p.numel()].as_strided(p.size(), p.stride())
        self._gpu = self._comm.device != "cpu"   # CPU jobs (libmpi shim backend): same buckets, synchronous collectives
        self._stream = torch.cuda.Stream(priority=-1) if self._gpu else None
        for b in self._buckets:
            b["pending"] = len(b["params"])
            for p in b["params"]:
                p.register_post_accumulate_grad_hook(self._hook(b))

    def _hook(self, b):
        def fn(_p):
            if self._pass + 1 < self._passes:
                return
            b["pending"] -= 1
            if b["pending"] == 0:
                self._fire(b)
        return fn

    def _fire(self, b):
        import contextlib
        scale = 1.0 / self._passes if self._passes > 1 else None
        if self._gpu:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._stream.wait_event(ev)
        with (torch.cuda.stream(self._stream) if self._gpu else contextlib.nullcontext()):
            if self._op == "adasum":
                from .adasum import adasum_allreduce_
                adasum_allreduce_(self._comm, self._flat[b["start"]:b["start"] + b["numel"]], stream=self._stream)
            else:
                self._comm.allreduce_window(self._win, b["start"] * 4, b["numel"], torch.float32, op=self._op, scale=scale,
                                            stream=self._stream)
        b["pending"] = -1

    def synchronize(self):
        for b in self._buckets:
            if b["pending"] != -1:
                self._fire(b)
        if self._gpu:
            torch.cuda.current_stream().wait_stream(self._stream)

    def step(self, closure=None):
        from ..utils import fault
        fault.injector().on_step()
        self._pass += 1
        if self._pass < self._passes:
            return None
        self._pass = 0
        self.synchronize()
        out = self._opt.step(closure)
        for b in self._buckets:
            b["pending"] = len(b["params"])
        return out

    def zero_grad(self, set_to_none: bool = False):
        self._flat.zero_()  # gradients stay views of the symmetric window

    def __getattr__(self, name):
        return getattr(self._opt, name)

    @property
    def param_groups(self):
        return self._opt.param_groups

    def state_dict(self):
        return self._opt.state_dict()

    def load_state_dict(self, sd):
        return self._opt.load_state_dict(sd)


def DistributedOptimizer(optimizer, named_parameters=None, compression=None, backward_passes_per_step=1, op="average",  # noqa: N802
                         gradient_predivide_factor=1.0, **kw):
    return _DistributedOptimizer(optimizer, named_parameters, compression, backward_passes_per_step, op,
                                 gradient_predivide_factor, kw.get("bucket_bytes"))
