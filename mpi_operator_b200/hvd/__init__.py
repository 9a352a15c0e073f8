"""Horovod-compatible front-end over the b200mpi runtime.

The reference's workloads call Horovod (examples/v2beta1/horovod/
tensorflow_mnist.py:90 ``hvd.init()``, :123-130 LR x ``hvd.size()``, :133
``hvd.DistributedOptimizer(opt, op=hvd.Average)``, :143 broadcast from rank 0,
:155 GPU pinning by ``hvd.local_rank()``, :159 rank-0-only checkpoints; and
``--variable_update=horovod`` in tensorflow-benchmarks.yaml:42).  Horovod's C++
core (negotiation thread, fusion buffer, NCCL calls) is replaced by:
gradients living in a symmetric window (no fusion-buffer copies), bucket
allreduce kernels with the average fused in, launched from autograd hooks on a
high-priority stream.  Usage is the ``horovod.torch`` one:

    import mpi_operator_b200.hvd as hvd      # or: import horovod.torch as hvd
    hvd.init(); torch.cuda.set_device(hvd.local_rank())
    opt = hvd.DistributedOptimizer(opt, named_parameters=model.named_parameters())
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

from ..launch.env import rank_info_from_env

Average, Sum, Adasum, Min, Max = "average", "sum", "adasum", "min", "max"

_state = {"comm": None, "info": None}


class HorovodNotInitialized(RuntimeError):
    pass


def _comm():
    if _state["comm"] is None:
        raise HorovodNotInitialized("hvd.init() has not been called")
    return _state["comm"]


def init(comm=None) -> None:
    """Join the job's rendezvous; device = LOCAL_RANK (set it first with torch.cuda.set_device if you prefer)."""
    if _state["comm"] is not None:
        return
    import torch
    from ..runtime.comm import Communicator
    info = rank_info_from_env()
    _state["info"] = info
    if comm is not None:
        _state["comm"] = comm
        return
    if not torch.cuda.is_available() or os.environ.get("B200MPI_HVD_DEVICE", "") == "cpu":
        # CPU job (the reference's Horovod MNIST example runs on CPU workers): collectives over the libmpi shim
        from .host_backend import HostCommunicator
        _state["comm"] = HostCommunicator()
        return
    dev = info.local_rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    _state["comm"] = Communicator.create(info.rank, info.world_size, dev, info.job_id)


def shutdown() -> None:
    c = _state["comm"]
    if c is not None:
        c.destroy()
    _state["comm"] = None


def is_initialized() -> bool:
    return _state["comm"] is not None


def rank() -> int: return _comm().rank  # noqa: E704
def size() -> int: return _comm().world  # noqa: E704
def local_rank() -> int: return _state["info"].local_rank  # noqa: E704
def local_size() -> int: return _state["info"].local_size  # noqa: E704
def cross_rank() -> int: return 0  # noqa: E704  (single box)
def cross_size() -> int: return 1  # noqa: E704
def mpi_threads_supported() -> bool: return False  # noqa: E704
def mpi_built() -> bool: return True  # noqa: E704
def mpi_enabled() -> bool: return True  # noqa: E704
def gloo_built() -> bool: return False  # noqa: E704
def gloo_enabled() -> bool: return False  # noqa: E704
def nccl_built() -> int: return 1  # noqa: E704  (tensorflow_mnist.py:127 checks this before Adasum)
def cuda_built() -> bool:
    import torch
    return torch.cuda.is_available()
def rocm_built() -> bool: return False  # noqa: E704
def ddl_built() -> bool: return False  # noqa: E704
def ccl_built() -> bool: return False  # noqa: E704


def _op_name(op, average):
    if average is not None:
        return "avg" if average else "sum"
    if op in (None, Average):
        return "avg"
    if op == Sum:
        return "sum"
    if op == Min:
        return "min"
    if op == Max:
        return "max"
    if op == Adasum:
        return "adasum"
    raise ValueError(f"unknown reduction op {op!r}")


def allreduce(tensor, average=None, name=None, op=None, prescale_factor=1.0, postscale_factor=1.0):
    out = tensor.clone()
    allreduce_(out, average, name, op, prescale_factor, postscale_factor)
    return out


def allreduce_(tensor, average=None, name=None, op=None, prescale_factor=1.0, postscale_factor=1.0):
    import torch
    if _op_name(op, average) == "adasum":
        from .adasum import adasum_allreduce_
        return adasum_allreduce_(_comm(), tensor)
    t = tensor if tensor.is_contiguous() else tensor.contiguous()
    if t.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        f = t.float()
        _comm().allreduce(f, f, op=_op_name(op, average), scale=prescale_factor * postscale_factor)
        tensor.copy_(f.to(tensor.dtype))
        return tensor
    _comm().allreduce(t, t, op=_op_name(op, average), scale=prescale_factor * postscale_factor)
    if t is not tensor:
        tensor.copy_(t)
    return tensor


def grouped_allreduce(tensors, average=None, name=None, op=None):
    return [allreduce(t, average, name, op) for t in tensors]


def allgather(tensor, name=None):
    import torch
    t = tensor.contiguous()
    out = torch.empty((size() * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    _comm().allgather(t, out)
    return out


def broadcast(tensor, root_rank, name=None):
    out = tensor.clone()
    return broadcast_(out, root_rank, name)


def broadcast_(tensor, root_rank, name=None):
    t = tensor if tensor.is_contiguous() else tensor.contiguous()
    _comm().broadcast(t, root=root_rank)
    if t is not tensor:
        tensor.copy_(t)
    return tensor


def alltoall(tensor, splits=None, name=None):
    """Even alltoall in one kernel; with ``splits`` (rows of dim 0 sent to each rank, horovod semantics) the rows are
    padded to the largest split so the same even kernel moves them, and ``(output, received_splits)`` is returned."""
    import torch
    t = tensor.contiguous()
    if splits is None:
        out = torch.empty_like(t)
        _comm().alltoall(t, out)
        return out
    n = size()
    sp = [int(v) for v in (splits.tolist() if hasattr(splits, "tolist") else splits)]
    if len(sp) != n or sum(sp) != t.shape[0] or min(sp) < 0:
        raise ValueError("alltoall: splits must have one non-negative entry per rank and sum to tensor.shape[0]")
    mine = torch.tensor(sp, dtype=torch.int32, device=t.device)
    all_splits = allgather(mine.view(1, n)).view(n, n)            # all_splits[src][dst]
    recv = [int(all_splits[src][rank()]) for src in range(n)]
    width = max(1, int(all_splits.max()))
    row = t.shape[1:]
    send = torch.zeros((n, width) + tuple(row), dtype=t.dtype, device=t.device)
    off = 0
    for dst in range(n):
        send[dst, :sp[dst]] = t[off:off + sp[dst]]
        off += sp[dst]
    got = torch.empty_like(send)
    _comm().alltoall(send, got)
    out = torch.cat([got[src, :recv[src]] for src in range(n)], dim=0) if sum(recv) else t.new_empty((0,) + tuple(row))
    return out, torch.tensor(recv, dtype=torch.int32)


def reducescatter(tensor, op=None, name=None):
    import torch
    t = tensor.contiguous()
    out = torch.empty((t.shape[0] // size(),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    _comm().reduce_scatter(t, out, op=_op_name(op, None))
    return out


def _on_gpu() -> bool:
    return _comm().device != "cpu"


def _dev():
    return "cuda" if _on_gpu() else "cpu"


def barrier():
    import torch
    _comm().barrier()
    if _on_gpu():
        torch.cuda.synchronize()


def join(device=-1) -> int:
    barrier()
    return size() - 1


def synchronize(handle=None):
    import torch
    if _on_gpu():
        torch.cuda.current_stream().synchronize()
    return handle


def poll(handle=None) -> bool:
    return True


# async flavours complete in stream order: the "handle" is the tensor itself
allreduce_async = allreduce
allreduce_async_ = allreduce_
allgather_async = allgather
broadcast_async = broadcast
broadcast_async_ = broadcast_


def broadcast_parameters(params, root_rank: int = 0) -> None:
    """K3 (tensorflow_mnist.py:143): state_dict / named_parameters / list of (name, tensor)."""
    import torch
    if isinstance(params, dict):
        items = sorted(params.items())
    else:
        items = list(params)
    for _, p in items:
        if isinstance(p, torch.Tensor) and (p.is_cuda or not _on_gpu()):
            t = p.data if p.is_contiguous() else p.data.contiguous()
            if t.numel():
                _comm().broadcast(t, root=root_rank)
                if t.data_ptr() != p.data_ptr():
                    p.data.copy_(t)


def broadcast_object(obj, root_rank: int = 0, name=None):
    import pickle
    import torch
    payload = pickle.dumps(obj) if rank() == root_rank else b""
    n = torch.tensor([len(payload)], dtype=torch.float32, device=_dev())
    _comm().broadcast(n, root=root_rank)
    ln = int(n.item())
    buf = torch.zeros(ln + (-ln) % 2, dtype=torch.uint8, device=_dev())
    if rank() == root_rank:
        buf[:ln] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(_dev())
    _comm().broadcast(buf, root=root_rank)
    return pickle.loads(bytes(buf[:ln].cpu().numpy()))


def allgather_object(obj, name=None) -> list:
    """One picklable object per rank -> list indexed by rank."""
    import pickle
    import torch
    payload = pickle.dumps(obj)
    lens = allgather(torch.tensor([len(payload)], dtype=torch.int64, device=_dev()))
    width = int(lens.max()) + (-int(lens.max())) % 2
    buf = torch.zeros(1, width, dtype=torch.uint8, device=_dev())
    buf[0, :len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(_dev())
    rows = allgather(buf).cpu()
    return [pickle.loads(bytes(rows[r, :int(lens[r])].numpy())) for r in range(size())]


def broadcast_optimizer_state(optimizer, root_rank: int = 0) -> None:
    sd = optimizer.state_dict() if rank() == root_rank else None
    sd = broadcast_object(sd, root_rank)
    if rank() != root_rank:
        optimizer.load_state_dict(sd)


class Compression:
    class none:  # noqa: N801
        @staticmethod
        def compress(t): return t, None  # noqa: E704

        @staticmethod
        def decompress(t, ctx): return t  # noqa: E704

    class fp16:  # noqa: N801
        @staticmethod
        def compress(t): return t.half(), t.dtype  # noqa: E704

        @staticmethod
        def decompress(t, ctx): return t.to(ctx)  # noqa: E704


from .optimizer import DistributedOptimizer  # noqa: E402,F401
from . import elastic  # noqa: E402,F401
