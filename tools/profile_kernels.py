"""Single-GPU driver for ncu: runs the multi-rank kernels in emulated mode (8 virtual
ranks = gridDim.y, one launch) so `ncu --set full` can replay them without a second
GPU.  ncu -k regex:k_allreduce picks the launches; numbers under ncu are for the
metrics, never for timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpi_operator_b200.runtime.comm import Communicator  # noqa: E402

W = 8
comm = Communicator.local(W, device=0)
comm.set_tuning(timeout_ms=20000)
n = (64 << 20) // 4  # 64 MiB fp32 per rank
win = comm.alloc_window(n * 4)
pwin = comm.alloc_window(n * 4)
moms = [torch.zeros(comm.slice_elems(n, torch.float32), device="cuda") for _ in range(W)]
for r in range(W):
    win.tensor(torch.float32, rank=r, numel=n).normal_()
    pwin.tensor(torch.float32, rank=r, numel=n).normal_()
small = [torch.randn(16384, device="cuda") for _ in range(W)]
for it in range(3):
    comm.allreduce_window(win, 0, n, torch.float32, op="avg", algo="twoshot")
    comm.allreduce(small, small, op="avg", algo="oneshot")
    comm.allreduce_sgd_window(win, 0, pwin, 0, moms, n, torch.float32, lr=0.1, momentum_coef=0.9, algo="twoshot")
torch.cuda.synchronize()
comm.check_error()
print("profile driver done, launches:", comm.launch_count)
