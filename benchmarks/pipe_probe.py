"""Diagnose the pipelined user-pointer kernel: timeline of one launch (rank 0, a few lanes) and a small configuration
scan (chunk size / lanes / depth) at one message size.  python tests/mp_launch.py -n 4 benchmarks/pipe_probe.py"""
import argparse
import ctypes as C
import os
import sys

import torch

os.environ.setdefault("B200MPI_PIPE_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpi_operator_b200.launch.env import rank_info_from_env  # noqa: E402
from mpi_operator_b200.runtime import _lib  # noqa: E402
from mpi_operator_b200.runtime.comm import Communicator  # noqa: E402


def timeline(comm):
    L = _lib.lib()
    L.b200mpi_pipe_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_size_t]
    L.b200mpi_pipe_timeline.restype = C.c_size_t
    n = 3 * 48 * 32 * 3
    buf = (C.c_uint64 * n)()
    got = L.b200mpi_pipe_timeline(comm._h, buf, n)
    if not got:
        return None
    import numpy as np
    return np.frombuffer(buf, dtype=np.uint64).reshape(3, 48, 32, 3).astype(np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--staging-mb", type=int, default=256)
    a = ap.parse_args()
    info = rank_info_from_env()
    dev = info.local_rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    comm = Communicator.create(info.rank, info.world_size, dev, info.job_id, staging_bytes=a.staging_mb << 20)
    comm.set_reg(0)
    n = (a.mb << 20) // 4
    t = torch.ones(n, device="cuda")

    def run(label, **pipe):
        comm.set_pipe(min_bytes=0, **pipe)
        for _ in range(2):
            comm.allreduce(t, t, op="avg")
        torch.cuda.synchronize()
        comm.host_barrier()
        timeline(comm)   # clear
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        comm.allreduce(t, t, op="avg")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tl = timeline(comm)
        comm.host_barrier()
        if comm.rank == 0:
            bus = (a.mb << 20) / (ms * 1e-3) / 1e9 * 2 * (comm.world - 1) / comm.world
            print(f"[{label}] {a.mb} MiB: {ms * 1e3:.0f} us, busbw {bus:.0f} GB/s", flush=True)
        return tl

    tl = run("default 1MiB x16 lanes x3", lanes_nvls=16, lanes_p2p=16, depth=3, chunk_bytes=1 << 20)
    if comm.rank == 0 and tl is not None:
        t0 = tl[tl > 0].min()
        for lane in (0, 7):
            print(f"--- rank 0 lane {lane}: per chunk [wait_begin work_begin end] us per role (in | reduce | out)")
            for j in range(8):
                row = []
                for role in range(3):
                    w, b, e = [(x - t0) / 1e3 if x else -1 for x in tl[role, lane, j]]
                    row.append(f"{w:7.1f} {b:7.1f} {e:7.1f}")
                print(f"chunk {j}: " + " | ".join(row), flush=True)
    for label, kw in [("1MiB x24 x3", dict(lanes_nvls=24, lanes_p2p=24, depth=3, chunk_bytes=1 << 20)),
                      ("1MiB x32 x3", dict(lanes_nvls=32, lanes_p2p=32, depth=3, chunk_bytes=1 << 20)),
                      ("1MiB x48 x3", dict(lanes_nvls=48, lanes_p2p=48, depth=3, chunk_bytes=1 << 20)),
                      ("2MiB x32 x2", dict(lanes_nvls=32, lanes_p2p=32, depth=2, chunk_bytes=2 << 20)),
                      ("512KiB x32 x4", dict(lanes_nvls=32, lanes_p2p=32, depth=4, chunk_bytes=512 << 10)),
                      ("512KiB x48 x4", dict(lanes_nvls=48, lanes_p2p=48, depth=4, chunk_bytes=512 << 10))]:
        run(label, **kw)
    comm.destroy()


if __name__ == "__main__":
    main()
