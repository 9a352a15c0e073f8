"""Rank script (tests/mp_launch.py, B200MPI_P2P=1): point-to-point over the mailbox window (csrc/kernels/p2p.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mpi_operator_b200.launch.env import rank_info_from_env
from mpi_operator_b200.runtime.comm import Communicator

info = rank_info_from_env()
r, n = info.rank, info.world_size
torch.cuda.set_device(info.local_rank)
comm = Communicator.create(r, n, info.local_rank, info.job_id)
assert comm.has_p2p, "B200MPI_P2P=1 was not honoured"
nxt, prv = (r + 1) % n, (r - 1) % n

# ring shift, grouped: every rank sends to the next and receives from the previous in ONE batch
for nbytes in (4, 1000, 1 << 20, (1 << 20) + 13, 5 << 20, (9 << 20) + 7):
    src = torch.full((nbytes,), r + 1, dtype=torch.uint8, device="cuda")
    src[-1] = 200 + r
    dst = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    comm.p2p_batch([("send", src, nxt), ("recv", dst, prv)])
    torch.cuda.synchronize()
    comm.check_error()
    assert int(dst[0]) == prv + 1 and int(dst[-1]) == 200 + prv and (nbytes < 3 or int(dst[nbytes // 2]) == prv + 1), (nbytes, dst[:4])

# eager: an ungrouped send of <= 2 MiB completes before the receive is posted (even ranks send first, odd ranks receive first)
if n % 2 == 0:
    a = torch.arange(300000, dtype=torch.float32, device="cuda") + r
    b = torch.empty_like(a)
    peer = r ^ 1
    comm.send(a, peer)
    comm.recv(b, peer)
    torch.cuda.synchronize()
    assert torch.equal(b, torch.arange(300000, dtype=torch.float32, device="cuda") + peer)

# full exchange (all-to-all by point-to-point), two messages per peer in one batch -> per-stream chunk offsets
outs = {p: [torch.zeros(70000, dtype=torch.int32, device="cuda"), torch.zeros(3, dtype=torch.int32, device="cuda")] for p in range(n) if p != r}
ops = []
for p in range(n):
    if p == r:
        continue
    ops += [("send", torch.full((70000,), 1000 * r + p, dtype=torch.int32, device="cuda"), p),
            ("send", torch.tensor([r, p, 7], dtype=torch.int32, device="cuda"), p),
            ("recv", outs[p][0], p), ("recv", outs[p][1], p)]
keep = [t for _, t, _ in ops]
for it in range(3):   # repeated: the device-side chunk counters advance between batches
    comm.p2p_batch(ops)
torch.cuda.synchronize()
comm.check_error()
for p, (big, small) in outs.items():
    assert int(big[0]) == 1000 * p + r and int(big[-1]) == 1000 * p + r and small.tolist() == [p, r, 7], (p, big[:3], small)

# CUDA graph capture: replays use the device-side counters
g = torch.cuda.CUDAGraph()
src = torch.zeros(4096, device="cuda")
dst = torch.zeros(4096, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    comm.p2p_batch([("send", src, nxt), ("recv", dst, prv)], stream=s)   # warm-up outside the graph
    s.synchronize()
    with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
        comm.p2p_batch([("send", src, nxt), ("recv", dst, prv)], stream=s)
for it in range(4):
    src.fill_(10 * it + r)
    torch.cuda.synchronize()
    comm.host_barrier()
    g.replay()
    torch.cuda.synchronize()
    assert float(dst[0]) == 10 * it + prv, (it, float(dst[0]))
comm.check_error()
print(f"rank {r}/{n} p2p ok", flush=True)
comm.destroy()
