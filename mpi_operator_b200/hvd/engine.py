"""ctypes binding of the native hvdcore engine (csrc/hvd_core): Horovod's background thread — name-based negotiation,
tensor fusion, response cache, timeline, stall inspector, join — rebuilt for one box (SURVEY.md §2.2 "Horovod core",
§3.3 negotiate -> fuse -> allreduce -> unfuse).  The ``hvd`` module routes its ``*_async`` API through it:

    h = hvd.allreduce_async_(t, name="grad.7")     # returns immediately, any order across ranks
    ...
    hvd.synchronize(h)                             # or hvd.poll(h)

Host tensors always go through the engine.  Device tensors do when ``B200MPI_HVD_ENGINE=1`` (the GPU executor — the
b200mpi kernels on the engine's own stream and communicator — was written after the round's GPU budget was spent and has
not run on hardware yet); otherwise they keep the direct stream-ordered path.
"""
from __future__ import annotations

import ctypes as C
import json
import threading
from pathlib import Path
from typing import List, Optional

import torch

from .exceptions import HorovodInternalError

LIB_PATH = Path(__file__).resolve().parent.parent / "lib" / "libb200mpi_hvd.so"

ALLREDUCE, ALLGATHER, BROADCAST, ALLTOALL, BARRIER, JOIN, EXCHANGE = range(7)
_DTYPES = {torch.uint8: 0, torch.int8: 1, torch.int16: 2, torch.int32: 3, torch.int64: 4, torch.float16: 5, torch.bfloat16: 6,
           torch.float32: 7, torch.float64: 8, torch.bool: 9}
_REDOPS = {"sum": 0, "avg": 0, "min": 1, "max": 2, "prod": 3, "adasum": 4}   # adasum: host float tensors (hvd_core.cc host_adasum)
ERR_SHUTDOWN, ERR_MISMATCH, ERR_TRANSPORT, ERR_DUPLICATE, ERR_STALL = -3, -4, -5, -6, -8


class _GpuExec(C.Structure):
    _fields_ = [("comm", C.c_void_p), ("device", C.c_int), ("allreduce", C.c_void_p), ("broadcast_bytes", C.c_void_p),
                ("barrier", C.c_void_p), ("last_error", C.c_void_p)]


_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise RuntimeError(f"{LIB_PATH} is missing; run `make` (or __graft_entry__.build())")
            L = C.CDLL(str(LIB_PATH))
            vp, i, i64, d = C.c_void_p, C.c_int, C.c_int64, C.c_double
            L.hvdcore_init.argtypes = [C.c_char_p, i, i, C.POINTER(_GpuExec)]
            L.hvdcore_enqueue.argtypes = [i, C.c_char_p, vp, vp, i64, i, i, i, d, d, i, vp, C.POINTER(i64), i]
            L.hvdcore_poll.argtypes = [i]
            L.hvdcore_wait.argtypes = [i]
            L.hvdcore_last_error.restype = C.c_char_p
            L.hvdcore_start_timeline.argtypes = [C.c_char_p]
            L.hvdcore_stats_json.argtypes = [C.c_char_p, C.c_size_t]
            L.hvdcore_set_param.argtypes = [C.c_char_p, d]
            _lib = L
    return _lib


def _err() -> str:
    return (lib().hvdcore_last_error() or b"?").decode(errors="replace")


class Handle:
    """Completion handle of one submitted collective. The buffers (and the CUDA event) stay referenced until the engine is
    done with them — also when the handle itself is dropped without ``wait()``: the references then move to the engine's
    orphan list, because the background thread may still be about to read or write that memory."""

    __slots__ = ("id", "keep", "result", "_status", "post", "_engine")

    def __init__(self, hid: int, keep, result=None, post=None, engine=None):
        self.id, self.keep, self.result, self._status, self.post, self._engine = hid, keep, result, None, post, engine

    def done(self) -> bool:
        return self._status is not None or lib().hvdcore_poll(self.id) == 1

    def wait(self):
        if self._status is None:
            st = lib().hvdcore_wait(self.id)       # ctypes releases the GIL while the engine works
            msg = _err() if st < 0 else ""
            self._status = st
            self.keep = None
            if st < 0:
                raise HorovodInternalError(msg, st)
            if self.post is not None:
                self.result = self.post(self.result)
                self.post = None
        elif self._status < 0:
            raise HorovodInternalError("the collective already failed", self._status)
        return self.result

    def __del__(self):
        if self._status is None and self._engine is not None:
            try:
                self._engine._orphans.append((self.id, self.keep))
            except Exception:   # interpreter shutdown
                pass


class Engine:
    """Process-wide engine (one per ``hvd.init()``)."""

    def __init__(self, job_id: str, rank: int, world: int, gpu_comm=None):
        self.rank, self.world = rank, world
        self._gpu_comm = gpu_comm
        ex = None
        if gpu_comm is not None:
            from ..runtime import _lib as rt
            R = rt.lib()
            addr = lambda f: C.cast(f, C.c_void_p).value  # noqa: E731
            ex = _GpuExec(gpu_comm._h, int(gpu_comm.device), addr(R.b200mpi_allreduce), addr(R.b200mpi_broadcast_bytes),
                          addr(R.b200mpi_barrier), addr(R.b200mpi_last_error))
        rc = lib().hvdcore_init(job_id.encode(), rank, world, C.byref(ex) if ex is not None else None)
        if rc:
            raise RuntimeError(f"hvdcore_init failed ({rc}): {_err()}")
        self.has_gpu = gpu_comm is not None
        self._alive = True
        self._orphans = []     # (handle id, buffers) of handles dropped before completion

    # ------------------------------------------------------------------------------------------------ submit --
    def _submit(self, op: int, name: Optional[str], tin, tout, count: int, dtype, redop: str = "sum", root: int = 0,
                prescale: float = 1.0, postscale: float = 1.0, extra: Optional[List[int]] = None, result=None, post=None) -> Handle:
        if self._orphans:
            self._reap()
        device, ev = -1, None
        ref = tout if tout is not None else tin
        if ref is not None and ref.is_cuda:
            if not self.has_gpu:
                raise HorovodInternalError("a CUDA tensor was submitted to an engine started without a GPU executor")
            device = ref.device.index
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(ref.device))
        arr = None
        if extra:
            arr = (C.c_int64 * len(extra))(*[int(v) for v in extra])
        h = lib().hvdcore_enqueue(op, name.encode() if name else None, tin.data_ptr() if tin is not None and tin.numel() else None,
                                  tout.data_ptr() if tout is not None and tout.numel() else None, int(count), _DTYPES[dtype],
                                  _REDOPS[redop], int(root), float(prescale), float(postscale), device,
                                  C.c_void_p(ev.cuda_event) if ev is not None else None, arr, len(extra) if extra else 0)
        if h <= 0:
            raise HorovodInternalError(_err(), h)
        return Handle(h, (tin, tout, ev, arr), result, post, self)

    def _reap(self, block: bool = False) -> None:
        """Release orphaned handles the engine has finished with (all of them when ``block``)."""
        L = lib()
        rest = []
        for hid, keep in self._orphans:
            if block or L.hvdcore_poll(hid) == 1:
                L.hvdcore_wait(hid)
            else:
                rest.append((hid, keep))
        self._orphans = rest

    def allreduce_async(self, tensor, out, name=None, op="sum", prescale=1.0, postscale=1.0) -> Handle:
        if op == "avg":
            postscale = postscale / self.world
        if op not in _REDOPS:
            raise ValueError(f"unknown reduction {op!r}")
        if tensor.dtype not in _DTYPES:
            raise ValueError(f"unsupported dtype {tensor.dtype}")
        if tensor.numel() == 0:   # nothing to move, but the name still has to match up across ranks
            return self._submit(BARRIER, name, None, None, 0, torch.uint8, result=out)
        return self._submit(ALLREDUCE, name, tensor, out, tensor.numel(), tensor.dtype, op, 0, prescale, postscale, result=out)

    def broadcast_async(self, tensor, root: int, name=None) -> Handle:
        if tensor.numel() == 0:
            return self._submit(BARRIER, name, None, None, 0, torch.uint8, result=tensor)
        return self._submit(BROADCAST, name, None, tensor, tensor.numel(), tensor.dtype, root=root, result=tensor)

    def exchange(self, blob: bytes, name=None) -> List[bytes]:
        """<= 1024 bytes per rank, carried by the negotiation itself (sizes for allgather / alltoall)."""
        n = len(blob)
        src = torch.frombuffer(bytearray(blob), dtype=torch.uint8) if n else torch.zeros(0, dtype=torch.uint8)
        dst = torch.zeros(max(1, n * self.world), dtype=torch.uint8)
        h = self._submit(EXCHANGE, name, src, dst, n, torch.uint8)
        h.wait()
        raw = bytes(dst.numpy())
        return [raw[r * n:(r + 1) * n] for r in range(self.world)]

    def allgather_async(self, tensor, name=None) -> Handle:
        """Concatenation along dim 0; first dimensions may differ between ranks (host tensors)."""
        import struct
        t = tensor.contiguous()
        tail = tuple(t.shape[1:]) if t.dim() else ()
        rows = t.shape[0] if t.dim() else 1
        sizes = [struct.unpack("q", b)[0] for b in self.exchange(struct.pack("q", rows), (name or "allgather") + ".sizes" if name else None)]
        per_row = 1
        for s in tail:
            per_row *= s
        esz = t.element_size()
        out = torch.empty((sum(sizes),) + tail, dtype=t.dtype)
        counts = [s * per_row * esz for s in sizes]
        if sum(counts) == 0:
            return self._submit(BARRIER, name, None, None, 0, torch.uint8, result=out)
        return self._submit(ALLGATHER, name, t, out, t.numel(), t.dtype, extra=counts, result=out)

    def alltoall_async(self, tensor, splits=None, name=None) -> Handle:
        import struct
        t = tensor.contiguous()
        W = self.world
        rows = t.shape[0]
        if splits is None:
            if rows % W:
                raise ValueError("alltoall: the first dimension must divide by the world size (or pass splits)")
            splits = [rows // W] * W
        splits = [int(s) for s in splits]
        if len(splits) != W or sum(splits) != rows:
            raise ValueError("alltoall: splits must have one entry per rank and add up to the first dimension")
        got = self.exchange(struct.pack(f"{W}q", *splits), (name + ".splits") if name else None)
        recv = [struct.unpack(f"{W}q", b)[self.rank] for b in got]
        per_row = t.numel() // rows if rows else 1
        if rows == 0:
            per_row = 1
            for s in t.shape[1:]:
                per_row *= s
        esz = t.element_size()
        out = torch.empty((sum(recv),) + tuple(t.shape[1:]), dtype=t.dtype)
        extra = [s * per_row * esz for s in splits] + [r * per_row * esz for r in recv]
        recv_t = torch.tensor(recv, dtype=torch.int32)
        return self._submit(ALLTOALL, name, t, out, t.numel(), t.dtype, extra=extra, result=(out, recv_t))

    def barrier(self, name=None) -> None:
        self._submit(BARRIER, name, None, None, 0, torch.uint8).wait()

    def join(self) -> int:
        h = self._submit(JOIN, None, None, None, 0, torch.uint8)
        st = lib().hvdcore_wait(h.id)
        h._status = st
        if st < 0:
            raise HorovodInternalError(_err(), st)
        return st

    # -------------------------------------------------------------------------------------------------- misc --
    def stats(self) -> dict:
        buf = C.create_string_buffer(1024)
        lib().hvdcore_stats_json(buf, 1024)
        return json.loads(buf.value.decode())

    def set_param(self, key: str, value: float) -> None:
        if lib().hvdcore_set_param(key.encode(), float(value)):
            raise ValueError(_err())

    def start_timeline(self, path: str) -> None:
        lib().hvdcore_start_timeline(path.encode())

    def stop_timeline(self) -> None:
        lib().hvdcore_stop_timeline()

    @property
    def alive(self) -> bool:
        return self._alive and lib().hvdcore_initialized() == 1

    def shutdown(self) -> None:
        if self._alive:
            self._alive = False
            lib().hvdcore_shutdown()     # every outstanding handle has failed or finished once this returns
            self._orphans = []
            if self._gpu_comm is not None:
                self._gpu_comm.destroy()
                self._gpu_comm = None
