#!/usr/bin/env python
"""CPU Horovod step time: window/bucket DistributedOptimizer vs the engine-backed one (per-parameter named async allreduce
negotiated, fused and reduced by the native background thread while backward is still running).

    mpirun -n 2 python benchmarks/hvd_cpu_bench.py --optimizer engine --steps 100

Model = the reference example's MNIST convnet (examples/v2beta1/horovod/tensorflow_mnist.py:38-73), batch 100, Adam.
Rank 0 prints one JSON line (ms/step is the max over ranks)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
os.environ.setdefault("B200MPI_HVD_DEVICE", "cpu")
import torch
import torch.nn.functional as F

import horovod.torch as hvd
from mpi_operator_b200.models import MnistConvNet


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--optimizer", choices=["bucket", "engine"], default="bucket")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch-size", type=int, default=100)
    a = ap.parse_args()
    hvd.init()
    torch.manual_seed(1)
    model = MnistConvNet()
    opt = hvd.DistributedOptimizer(torch.optim.Adam(model.parameters(), lr=1e-3), named_parameters=model.named_parameters(),
                                   engine=a.optimizer == "engine")
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    x = torch.randn(a.batch_size, 784)
    y = torch.randint(0, 10, (a.batch_size,))

    def step():
        opt.zero_grad()
        F.cross_entropy(model(x), y).backward()
        opt.step()

    for _ in range(a.warmup):
        step()
    hvd.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    worst = float(hvd.allreduce(torch.tensor([dt]), op=hvd.Max))
    if hvd.rank() == 0:
        nparam = sum(p.numel() for p in model.parameters())
        print(json.dumps({"optimizer": type(opt).__name__, "ranks": hvd.size(), "steps": a.steps, "ms_per_step": round(1e3 * worst / a.steps, 3),
                          "gradient_bytes_per_step": 4 * nparam, "threads": torch.get_num_threads(), "engine": hvd.engine_stats()}))
    hvd.shutdown()


if __name__ == "__main__":
    main()
