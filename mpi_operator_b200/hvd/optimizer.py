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
        from . import _op_name
        self._opt = optimizer
        self._op = _op_name(op, None)
        self._passes = max(1, int(backward_passes_per_step))
        from . import Compression
        if compression not in (None, Compression.none):
            raise ValueError("the window/bucket DistributedOptimizer does not compress gradients; pass engine=True (or set "
                             "B200MPI_HVD_OPTIMIZER=engine) for compression")
        self._predivide = float(gradient_predivide_factor)
        if self._predivide != 1.0 and self._op != "avg":
            raise ValueError("gradient_predivide_factor requires op=Average")
        params = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
        if any(p.dtype != torch.float32 for p in params):
            raise ValueError("DistributedOptimizer expects fp32 parameters (use autocast for low-precision compute)")
        self._all_params = params
        self._params = list(reversed(params))
        self._bucket_bytes = bucket_bytes
        self._hooked = False
        self._attach()
        from . import _state
        import weakref
        _state.setdefault("optimizers", []).append(weakref.ref(self))

    def _detach(self):
        """Before the communicator goes away (elastic rescale in place): gradients stop viewing its window."""
        for p in self._all_params:
            p.grad = None
        self._flat = self._win = None
        self._handles = []

    def _attach(self):
        """(Re)build the gradient window and the buckets in the CURRENT communicator."""
        from . import _comm, _state
        self._comm = _comm()
        params, bucket_bytes = self._all_params, self._bucket_bytes
        on_host = getattr(self._comm, "device", None) == "cpu"   # smaller buckets on the host: they are what overlaps with backward
        cap = (bucket_bytes or int(os.environ.get("B200MPI_BUCKET_BYTES", (4 << 20) if on_host else (32 << 20)))) // 4
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
        # Host tensors: each bucket goes to the background engine as a named async allreduce, so the reduction of bucket k runs
        # while autograd produces bucket k+1 (measured on the MNIST convnet, 2 ranks: 72-77 ms/step vs 77-83 with the synchronous
        # libmpi call in the hook; single rank 72.8). B200MPI_HVD_BUCKET_ASYNC=0 keeps the synchronous path.
        self._engine = _state.get("engine") if (not self._gpu and os.environ.get("B200MPI_HVD_BUCKET_ASYNC", "1") != "0") else None
        self._handles = []
        self._counts = {p: 0 for p in params}   # backward passes seen per parameter since the last step()
        bucket_of = {}
        for b in self._buckets:
            b["pending"] = len(b["params"])
            for p in b["params"]:
                bucket_of[p] = b
        self._bucket_of = bucket_of
        if not self._hooked:       # hooks are registered once; they look the bucket up at call time (buckets are rebuilt on re-attach)
            for p in params:
                p.register_post_accumulate_grad_hook(self._hook_for)
            self._hooked = True

    def _hook_for(self, p):
        self._hook(self._bucket_of[p])(p)

    def _rehome(self, p):
        """``p.grad`` must stay a view of the symmetric window (that is what the bucket kernel reduces). Code that drops
        it (``model.zero_grad()`` defaults to set_to_none=True, ``p.grad = None``) makes autograd allocate a fresh tensor:
        copy it into the window slot and point ``p.grad`` back at the view instead of silently reducing stale zeros."""
        _, start = self._slot[p]
        view = self._flat[start:start + p.numel()].as_strided(p.size(), p.stride())
        if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
            if self._counts[p] > 1:
                view.add_(p.grad)     # earlier passes of this accumulation round are already in the window
            else:
                view.copy_(p.grad)
            p.grad = view

    def _hook(self, b):
        def fn(p):
            # Horovod usage with backward_passes_per_step = N: N x backward() then ONE step(). Count backward passes per
            # parameter (autograd accumulates into the window view) and reduce the bucket when every parameter of it has
            # seen its N-th pass.
            self._counts[p] += 1
            self._rehome(p)
            if self._counts[p] == self._passes:
                b["pending"] -= 1
                if b["pending"] == 0:
                    self._fire(b)
        return fn

    def _fire(self, b):
        import contextlib
        scale = 1.0 / self._passes if self._passes > 1 else None   # mean over the local passes, fused into the reduction
        if self._engine is not None and self._op != "adasum":
            from . import allreduce_async_
            self._handles.append(allreduce_async_(self._flat[b["start"]:b["start"] + b["numel"]], name=f"DistributedOptimizer.bucket.{b['start']}",
                                                  op=_OPS[self._op], postscale_factor=scale or 1.0))
            b["pending"] = -1
            return
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
        for p in self._params:      # gradients replaced behind our back since the last hook
            if p.grad is None:
                _, start = self._slot[p]
                p.grad = self._flat[start:start + p.numel()].as_strided(p.size(), p.stride())
            else:
                self._rehome(p)
        for b in self._buckets:
            if b["pending"] != -1:  # parameters without a gradient this round, or fewer backward passes than configured
                self._fire(b)
        for h in self._handles:
            h.wait()
        self._handles = []
        if self._gpu:
            torch.cuda.current_stream().wait_stream(self._stream)

    def step(self, closure=None):
        from ..utils import fault
        fault.injector().on_step()
        self.synchronize()          # always: step() is the one call per update (Horovod semantics)
        out = self._opt.step(closure)
        for b in self._buckets:
            b["pending"] = len(b["params"])
        for p in self._counts:
            self._counts[p] = 0
        return out

    def zero_grad(self, set_to_none: bool = False):
        """Zeroes the window; gradients stay views of it whatever ``set_to_none`` says (they must remain peer-visible)."""
        self._flat.zero_()
        for p in self._params:
            if p.grad is None or p.grad.data_ptr() != self._flat[self._slot[p][1]:].data_ptr():
                _, start = self._slot[p]
                p.grad = self._flat[start:start + p.numel()].as_strided(p.size(), p.stride())

    def __getattr__(self, name):
        return getattr(self._opt, name)

    @property
    def param_groups(self):
        return self._opt.param_groups

    def state_dict(self):
        return self._opt.state_dict()

    def load_state_dict(self, sd):
        return self._opt.load_state_dict(sd)


class _EngineDistributedOptimizer:
    """Horovod's own scheme, on the native engine: every parameter's gradient is submitted as a NAMED async allreduce from
    its autograd hook, the background thread negotiates / fuses / reduces while backward is still running, ``step()``
    waits for the handles. Unlike the bucket optimizer it tolerates ranks producing gradients in different orders and
    parameters that receive no gradient on some ranks (they contribute zeros), and it supports ``compression``."""

    def __init__(self, optimizer, named_parameters=None, compression=None, backward_passes_per_step: int = 1, op="average",
                 gradient_predivide_factor: float = 1.0):
        from . import Compression, _op_name
        self._opt = optimizer
        self._op = _op_name(op, None)
        self._passes = max(1, int(backward_passes_per_step))
        self._compression = compression or Compression.none
        self._predivide = float(gradient_predivide_factor)
        params = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
        named = list(named_parameters) if named_parameters is not None else []
        by_id = {id(p): n for n, p in named}
        if named and len(by_id) != len(named):
            raise ValueError("named_parameters contains the same parameter twice")
        self._names = {p: by_id.get(id(p), f"noname.{i}") for i, p in enumerate(params)}
        if len(set(self._names.values())) != len(params):
            raise ValueError("parameter names must be unique")
        self._params = params
        self._handles = {}
        self._counts = {p: 0 for p in params}
        for p in params:
            p.register_post_accumulate_grad_hook(self._hook)

    def _submit(self, p):
        from . import allreduce_async_
        grad = p.grad
        if self._passes > 1:
            grad.div_(self._passes)
        t, ctx = self._compression.compress(grad)
        pre = post = 1.0
        if self._op == "avg" and self._predivide != 1.0:
            pre, post = 1.0 / self._predivide, self._predivide
        h = allreduce_async_(t, name="DistributedOptimizer.grad." + self._names[p], op=_OPS[self._op], prescale_factor=pre,
                             postscale_factor=post)
        self._handles[p] = (h, t, ctx)

    def _hook(self, p):
        self._counts[p] += 1
        if self._counts[p] == self._passes:
            self._submit(p)

    def synchronize(self):
        import torch
        for p in self._params:   # no gradient on this rank this step: the other ranks still expect the tensor
            if p not in self._handles and self._counts[p] == 0:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                self._submit(p)
        for p, (h, t, ctx) in list(self._handles.items()):
            h.wait()
            out = self._compression.decompress(t, ctx)
            if out.data_ptr() != p.grad.data_ptr():
                p.grad.copy_(out)
        self._handles.clear()
        for p in self._params:
            self._counts[p] = 0

    def step(self, closure=None):
        from ..utils import fault
        fault.injector().on_step()
        if any(0 < c < self._passes for c in self._counts.values()):
            return None          # still accumulating local backward passes
        self.synchronize()
        return self._opt.step(closure)

    def zero_grad(self, set_to_none: bool = False):
        if self._handles:
            raise AssertionError("optimizer.zero_grad() was called after loss.backward() but before optimizer.step() or "
                                 "optimizer.synchronize(): the submitted gradients would be lost")
        return self._opt.zero_grad(set_to_none=set_to_none)

    def __getattr__(self, name):
        return getattr(self._opt, name)

    @property
    def param_groups(self):
        return self._opt.param_groups

    def state_dict(self):
        return self._opt.state_dict()

    def load_state_dict(self, sd):
        return self._opt.load_state_dict(sd)


class _DistributedAdasumOptimizer:
    """``DistributedOptimizer(..., op=hvd.Adasum)``: Horovod applies Adasum to the MODEL DELTAS, not to the gradients -
    every rank takes its own optimizer step from the common starting point, the per-tensor deltas are combined with Adasum
    (orthogonal updates add up, parallel ones average), and the combined delta is applied to the starting point, so any
    wrapped optimizer (Adam's per-coordinate scaling included) keeps its meaning and the learning rate is NOT multiplied by
    the world size (reference example: examples/v2beta1/horovod/tensorflow_mnist.py:126-133). Host tensors go through
    the native engine (named async Adasum per parameter, csrc/hvd_core/hvd_core.cc: host_adasum); device tensors through
    one allgather kernel + a local fp32 tree (hvd/adasum.py)."""

    def __init__(self, optimizer, named_parameters=None, compression=None, backward_passes_per_step: int = 1,
                 gradient_predivide_factor: float = 1.0):
        from . import Compression
        if compression not in (None, Compression.none):
            raise ValueError("op=hvd.Adasum does not combine with gradient compression")
        if float(gradient_predivide_factor) != 1.0:
            raise ValueError("gradient_predivide_factor requires op=Average")
        self._opt = optimizer
        self._passes = max(1, int(backward_passes_per_step))
        params = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
        named = list(named_parameters) if named_parameters is not None else []
        by_id = {id(p): n for n, p in named}
        self._names = {p: by_id.get(id(p), f"noname.{i}") for i, p in enumerate(params)}
        if len(set(self._names.values())) != len(params):
            raise ValueError("parameter names must be unique")
        self._params = params
        self._counts = {p: 0 for p in params}
        self._step_no = 0
        for p in params:
            p.register_post_accumulate_grad_hook(self._hook)

    def _hook(self, p):
        self._counts[p] += 1

    def synchronize(self):   # API compatibility: the collective work happens inside step()
        return None

    def step(self, closure=None):
        from . import allreduce_async_, Adasum
        from ..utils import fault
        fault.injector().on_step()
        if any(0 < c < self._passes for c in self._counts.values()):
            return None          # still accumulating local backward passes
        if self._passes > 1:
            for p in self._params:
                if p.grad is not None:
                    p.grad.div_(self._passes)
        start = [p.detach().clone() for p in self._params]
        loss = self._opt.step(closure)                 # local step from the common starting point
        handles = []
        for p, s0 in zip(self._params, start):
            delta = p.detach() - s0                    # zero where this rank had no gradient: neutral for Adasum
            handles.append((allreduce_async_(delta, name=f"DistributedOptimizer.adasum.{self._names[p]}", op=Adasum), delta))
        with torch.no_grad():
            for (h, delta), p, s0 in zip(handles, self._params, start):
                h.wait()
                p.copy_(s0 + delta)
        for p in self._params:
            self._counts[p] = 0
        self._step_no += 1
        return loss

    def zero_grad(self, set_to_none: bool = False):
        return self._opt.zero_grad(set_to_none=set_to_none)

    def __getattr__(self, name):
        return getattr(self._opt, name)

    @property
    def param_groups(self):
        return self._opt.param_groups

    def state_dict(self):
        return self._opt.state_dict()

    def load_state_dict(self, sd):
        return self._opt.load_state_dict(sd)


_OPS = {"avg": "average", "sum": "sum", "min": "min", "max": "max", "adasum": "adasum"}


def DistributedOptimizer(optimizer, named_parameters=None, compression=None, backward_passes_per_step=1, op="average",  # noqa: N802
                         gradient_predivide_factor=1.0, **kw):
    """``B200MPI_HVD_OPTIMIZER=engine`` (or ``engine=True``) selects Horovod's per-parameter scheme on the native background
    engine; the default is the window/bucket optimizer (gradients live in NVLink-visible memory, no fusion copies)."""
    import os
    from . import Compression, _op_name, _state, global_process_set
    if kw.get("process_set") not in (None, global_process_set):
        raise NotImplementedError("DistributedOptimizer works on the global process set; reduce over a subset with "
                                  "hvd.allreduce(..., process_set=ps) in your own step")
    use_engine = kw.get("engine")
    if use_engine is None:
        use_engine = os.environ.get("B200MPI_HVD_OPTIMIZER", "") == "engine"
    needs_engine = compression not in (None, Compression.none)
    if _op_name(op, None) == "adasum":
        return _DistributedAdasumOptimizer(optimizer, named_parameters, compression, backward_passes_per_step,
                                           gradient_predivide_factor)
    if (use_engine or needs_engine) and _state.get("engine") is not None and _op_name(op, None) != "adasum":
        return _EngineDistributedOptimizer(optimizer, named_parameters, compression, backward_passes_per_step, op,
                                           gradient_predivide_factor)
    return _DistributedOptimizer(optimizer, named_parameters, compression, backward_passes_per_step, op,
                                 gradient_predivide_factor, kw.get("bucket_bytes"))
