"""Small single-GPU driver for compute-sanitizer: every collective kernel family in emulated mode
(4 virtual ranks = gridDim.y, one launch each), sizes kept small because the tools slow kernels down 10-100x.
cross-CTA flag signalling is what racecheck cannot see (it checks shared memory); memcheck / synccheck / initcheck cover
the global-memory indexing of the slices, slots and staging rings."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpi_operator_b200.runtime.comm import Communicator  # noqa: E402

W = 4
comm = Communicator.local(W, device=0)
comm.set_tuning(timeout_ms=120000)     # the tools make the spin waits very slow: never let the watchdog fire
n = 65536 + 24
xs = [torch.randn(n, device="cuda") for _ in range(W)]
outs = [torch.empty_like(x) for x in xs]
want = torch.stack(xs).sum(0)
win = comm.alloc_window(n * 4 + 64)
for r in range(W):
    win.tensor(torch.float32, rank=r, numel=n).copy_(xs[r])
comm.allreduce(xs, outs, algo="oneshot")
assert torch.allclose(outs[0], want, atol=1e-4)
comm.allreduce(xs, outs, algo="twoshot")
assert torch.allclose(outs[1], want, atol=1e-4)
comm.allreduce_window(win, 0, (n // 4) * 4, torch.float32, op="sum", algo="twoshot")
comm.set_pipe(min_bytes=0, chunk_bytes=32 << 10, depth=2)
comm.allreduce(xs, outs, algo="twoshot")                       # k_pipe<ALLREDUCE> incl. TMA ring
assert torch.allclose(outs[2], want, atol=1e-4)
ag = [torch.empty(W * n, device="cuda") for _ in range(W)]
comm.allgather(xs, ag)                                         # k_pipe<ALLGATHER>
assert torch.equal(ag[0], torch.cat(xs))
full = [torch.randn(W * 4096, device="cuda") for _ in range(W)]
rs = [torch.empty(4096, device="cuda") for _ in range(W)]
comm.reduce_scatter(full, rs)                                  # k_pipe<REDUCE_SCATTER>
bs = [x.clone() for x in xs]
comm.broadcast(bs, root=1)                                     # k_pipe<BROADCAST>
assert torch.equal(bs[3], xs[1])
comm.set_pipe(min_bytes=1 << 40)
comm.allgather(xs, ag)
comm.reduce_scatter(full, rs)
comm.broadcast(bs, root=2)
comm.alltoall(full, [torch.empty_like(f) for f in full])
comm.allgather_window(win, 0, ((n * 4) // W) // 16 * 16)
comm.reduce_scatter_window(win, 0, ((n // W) // 4) * 4, torch.float32)
comm.broadcast_window(win, 0, (n // 4) * 16, root=0)
pwin = comm.alloc_window(n * 4 + 64)
moms = [torch.zeros(comm.slice_elems((n // 8) * 8, torch.float32), device="cuda") for _ in range(W)]
comm.allreduce_sgd_window(win, 0, pwin, 0, moms, (n // 8) * 8, torch.float32, lr=0.1, momentum_coef=0.9, first_step=True, algo="twoshot")
comm.barrier()
torch.cuda.synchronize()
comm.check_error()
# fused BN kernels
from mpi_operator_b200.ops.fused_bn import bn_act  # noqa: E402
bn = torch.nn.BatchNorm2d(64).cuda()
x = torch.randn(8, 64, 14, 14, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
y = bn_act(bn, x)
y.float().sum().backward()
torch.cuda.synchronize()
print("sanitize driver done, launches:", comm.launch_count)
