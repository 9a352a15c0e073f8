"""CPU backend of the Horovod-compatible front-end.

The reference's Horovod MNIST example is a CPU job (examples/v2beta1/horovod/tensorflow-mnist.yaml: workers with
``cpu: 2``, no GPU; Horovod falls back to MPI collectives on host memory). On a host without CUDA ``hvd.init()`` builds
this communicator instead of the NVLink one: same method surface as ``runtime.comm.Communicator`` for everything the
``hvd`` module calls, transport = the in-tree ``libmpi`` shim (``csrc/mpi_shim``: shared-memory mailboxes of the job's
rendezvous segment, ranks and sizes from the environment our ``mpirun`` sets). It is the control path made usable for
small models — the data plane of this framework is the GPU runtime."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import torch

LIB_PATH = Path(__file__).resolve().parent.parent / "lib" / "libmpi.so"

# csrc/mpi_shim/mpi.h
_MPI_TYPES = {torch.uint8: 3, torch.int8: 2, torch.int16: 5, torch.int32: 7, torch.int64: 11, torch.float32: 13,
              torch.float64: 14, torch.bool: 19}
_MPI_OPS = {"sum": 1, "avg": 1, "max": 2, "min": 3, "prod": 4}
_BYTE = 4
_COMM_WORLD = 0
_IN_PLACE = C.c_void_p(1)


class HostWindow:
    """Stand-in for a symmetric window: on the host it is just this rank's buffer."""

    def __init__(self, wid: int, nbytes: int):
        self.id, self.nbytes = wid, nbytes
        self.buf = torch.zeros(nbytes, dtype=torch.uint8)
        self.has_multicast = False

    def tensor(self, dtype=None, rank: int = -1, offset: int = 0, numel: Optional[int] = None):
        dtype = dtype or torch.uint8
        esz = torch.empty((), dtype=dtype).element_size()
        if numel is None:
            numel = (self.nbytes - offset) // esz
        return self.buf[offset:offset + numel * esz].view(dtype)

    def free(self) -> None:
        self.buf = None


class HostCommunicator:
    device = "cpu"
    is_local = False
    has_multicast = False

    def __init__(self, _parent: "Optional[HostCommunicator]" = None, _handle: int = _COMM_WORLD):
        if _parent is not None:      # a sub-communicator (sub()): shares the library and MPI_Init of its parent
            self._L, self._comm_h, self._owns_mpi = _parent._L, _handle, False
            r, n = C.c_int(0), C.c_int(1)
            self._L.MPI_Comm_rank(self._comm_h, C.byref(r))
            self._L.MPI_Comm_size(self._comm_h, C.byref(n))
            self.rank, self.world = r.value, n.value
            self.launch_count, self._windows, self._alive = 0, 0, True
            return
        self._comm_h = _COMM_WORLD
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing; run `make` (or __graft_entry__.build())")
        L = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
        vp, i = C.c_void_p, C.c_int
        L.MPI_Allreduce.argtypes = [vp, vp, i, i, i, i]
        L.MPI_Bcast.argtypes = [vp, i, i, i, i]
        L.MPI_Allgather.argtypes = [vp, i, i, vp, i, i, i]
        L.MPI_Alltoall.argtypes = [vp, i, i, vp, i, i, i]
        L.MPI_Barrier.argtypes = [i]
        L.MPI_Send.argtypes = [vp, i, i, i, i, i]
        L.MPI_Recv.argtypes = [vp, i, i, i, i, i, vp]
        self._L = L
        flag = i(0)
        L.MPI_Initialized(C.byref(flag))
        self._owns_mpi = not flag.value
        if self._owns_mpi:
            self._check(L.MPI_Init(None, None), "MPI_Init")
        r, n = i(0), i(1)
        L.MPI_Comm_rank(_COMM_WORLD, C.byref(r))
        L.MPI_Comm_size(_COMM_WORLD, C.byref(n))
        self.rank, self.world = r.value, n.value
        self.launch_count = 0
        self._windows = 0
        self._alive = True

    @staticmethod
    def _check(rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed with MPI error {rc}")

    @staticmethod
    def _host(t: torch.Tensor, what: str) -> torch.Tensor:
        if t.is_cuda:
            raise ValueError(f"{what}: the host backend takes CPU tensors")
        if not t.is_contiguous():
            raise ValueError(f"{what} needs contiguous tensors")
        return t

    # ---------------------------------------------------------- collectives --
    def allreduce(self, tensor, out=None, op: str = "sum", scale: Optional[float] = None, algo=None, stream=None):
        out = tensor if out is None else out
        t, o = self._host(tensor, "allreduce"), self._host(out, "allreduce")
        if op not in _MPI_OPS:
            raise ValueError(f"unknown reduction {op!r}")
        self.launch_count += 1
        # 16-bit floats (and anything MPI has no type for) are reduced in fp32, like the GPU kernels accumulate
        work = t if t.dtype in _MPI_TYPES and t.dtype != torch.bool else t.float()
        res = o if (work is t and o.dtype == t.dtype) else torch.empty_like(work)
        src = _IN_PLACE if res.data_ptr() == work.data_ptr() else C.c_void_p(work.data_ptr())
        if work.numel():
            self._check(self._L.MPI_Allreduce(src, res.data_ptr(), work.numel(), _MPI_TYPES[work.dtype], _MPI_OPS[op], self._comm_h),
                        "MPI_Allreduce")
        factor = (1.0 / self.world if op == "avg" else 1.0) * (1.0 if scale is None else float(scale))
        if factor != 1.0:
            if res.dtype.is_floating_point:
                res.mul_(factor)
            else:
                res.copy_((res.double() * factor).to(res.dtype))
        if res is not o:
            o.copy_(res.to(o.dtype))
        return out

    def allreduce_window(self, win: HostWindow, offset: int, count: int, dtype, op: str = "sum", scale: Optional[float] = None,
                         algo=None, stream=None) -> None:
        self.allreduce(win.tensor(dtype, offset=offset, numel=count), op=op, scale=scale)

    def broadcast(self, tensor, root: int = 0, stream=None):
        t = self._host(tensor, "broadcast")
        self.launch_count += 1
        if t.numel():
            self._check(self._L.MPI_Bcast(t.data_ptr(), t.numel() * t.element_size(), _BYTE, root, self._comm_h), "MPI_Bcast")
        return tensor

    def allgather(self, tensor, out, stream=None):
        t, o = self._host(tensor, "allgather"), self._host(out, "allgather")
        nbytes = t.numel() * t.element_size()
        if o.numel() * o.element_size() != nbytes * self.world:
            raise ValueError("allgather: output must hold world x input")
        self.launch_count += 1
        if nbytes:
            self._check(self._L.MPI_Allgather(t.data_ptr(), nbytes, _BYTE, o.data_ptr(), nbytes, _BYTE, self._comm_h), "MPI_Allgather")
        return out

    def alltoall(self, tensor, out, stream=None):
        t, o = self._host(tensor, "alltoall"), self._host(out, "alltoall")
        nbytes = t.numel() * t.element_size()
        if nbytes % self.world or o.numel() * o.element_size() != nbytes:
            raise ValueError("alltoall: size must divide by the world size and match the output")
        self.launch_count += 1
        if nbytes:
            per = nbytes // self.world
            self._check(self._L.MPI_Alltoall(t.data_ptr(), per, _BYTE, o.data_ptr(), per, _BYTE, self._comm_h), "MPI_Alltoall")
        return out

    def reduce_scatter(self, tensor, out, op: str = "sum", scale: Optional[float] = None, stream=None):
        t, o = self._host(tensor, "reduce_scatter"), self._host(out, "reduce_scatter")
        if t.numel() != o.numel() * self.world:
            raise ValueError("reduce_scatter: input must hold world x output")
        full = self.allreduce(t.clone(), op=op, scale=scale)
        o.copy_(full.view(self.world, -1)[self.rank].view_as(o))
        return out

    def barrier(self, stream=None) -> None:
        self.launch_count += 1
        self._check(self._L.MPI_Barrier(self._comm_h), "MPI_Barrier")

    # ------------------------------------------------------ point-to-point --
    _P2P_PIECE = 1 << 30   # MPI counts are C ints: larger messages travel as 1 GiB pieces (same order on both sides)

    def send(self, tensor, peer: int, tag: int = 0, stream=None) -> None:
        """Blocking standard-mode send of the tensor's bytes (MPI_Send, csrc/mpi_shim/mpi_p2p.cc: eager, buffered at the receiver)."""
        t = self._host(tensor, "send")
        nbytes, base = t.numel() * t.element_size(), t.data_ptr()
        self.launch_count += 1
        for off in range(0, max(nbytes, 1), self._P2P_PIECE):
            n = min(self._P2P_PIECE, nbytes - off)
            self._check(self._L.MPI_Send(base + off if n else None, n, _BYTE, int(peer), int(tag), self._comm_h), "MPI_Send")

    def recv(self, tensor, peer: int, tag: int = 0, stream=None) -> None:
        t = self._host(tensor, "recv")
        nbytes, base = t.numel() * t.element_size(), t.data_ptr()
        self.launch_count += 1
        for off in range(0, max(nbytes, 1), self._P2P_PIECE):
            n = min(self._P2P_PIECE, nbytes - off)
            self._check(self._L.MPI_Recv(base + off if n else None, n, _BYTE, int(peer), int(tag), self._comm_h, None), "MPI_Recv")

    host_barrier = barrier

    def sub(self, ranks, tag: int = 0) -> "Optional[HostCommunicator]":
        """Sub-communicator of the given WORLD ranks (in that order); collective over those ranks only
        (MPI_Comm_create_group, csrc/mpi_shim/mpi_comm.cc). None on a rank that is not a member."""
        L, i = self._L, C.c_int
        wg, sg, nc = i(0), i(0), i(-1)
        arr = (i * len(ranks))(*[int(r) for r in ranks])
        self._check(L.MPI_Comm_group(_COMM_WORLD, C.byref(wg)), "MPI_Comm_group")
        self._check(L.MPI_Group_incl(wg, len(ranks), arr, C.byref(sg)), "MPI_Group_incl")
        self._check(L.MPI_Comm_create_group(_COMM_WORLD, sg, int(tag) & 0xffff, C.byref(nc)), "MPI_Comm_create_group")
        L.MPI_Group_free(C.byref(wg))
        L.MPI_Group_free(C.byref(sg))
        return None if nc.value < 0 else HostCommunicator(_parent=self, _handle=nc.value)

    # ---------------------------------------------------------------- misc --
    def alloc_window(self, nbytes: int) -> HostWindow:
        self._windows += 1
        return HostWindow(self._windows - 1, nbytes)

    def slice_elems(self, count: int, dtype) -> int:
        return count

    def set_hyper(self, tensor) -> None:
        self._hyper = tensor

    def check_error(self) -> None:
        pass

    def stats(self, native_only: bool = False) -> dict:
        return {"rank": self.rank, "world": self.world, "launches": self.launch_count, "ops": []}

    def dump_stats(self, directory=None):
        return None

    def destroy(self) -> None:
        if self._alive and self._owns_mpi:
            self._L.MPI_Finalize()
        self._alive = False

    def reinit(self) -> None:
        """Elastic rescale in place: leave the current world and join the one the environment now describes
        (B200MPI_RANK / B200MPI_WORLD_SIZE / B200MPI_JOB_ID); every rank of the old world calls this or ``destroy``."""
        self._check(self._L.b200mpi_mpi_reinit(), "b200mpi_mpi_reinit")
        r, n = C.c_int(0), C.c_int(1)
        self._L.MPI_Comm_rank(_COMM_WORLD, C.byref(r))
        self._L.MPI_Comm_size(_COMM_WORLD, C.byref(n))
        self.rank, self.world = r.value, n.value
        self._alive = True
