#!/usr/bin/env python
"""Horovod MNIST example (counterpart of the reference's examples/v2beta1/horovod/
tensorflow_mnist.py): hvd.init, GPU pinned by local rank, LR x size, DistributedOptimizer
with Average, broadcast from rank 0, steps / size, rank-0-only checkpoint — on the
b200mpi runtime through the horovod.torch-compatible API. Synthetic MNIST-shaped data
(no dataset download on an air-gapped box)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))

import torch
import torch.nn.functional as F

import horovod.torch as hvd
from mpi_operator_b200.models import MnistConvNet


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--use-adasum", action="store_true", help="use the Adasum reduction instead of averaging (reference flag)")
    ap.add_argument("--steps", type=int, default=200, help="global steps; each rank runs steps // hvd.size()")
    ap.add_argument("--batch-size", type=int, default=100)
    ap.add_argument("--checkpoint-dir", default="")
    a = ap.parse_args()
    hvd.init()
    if torch.cuda.is_available() and os.environ.get("B200MPI_HVD_DEVICE") != "cpu":   # GPU pinned by local rank (tensorflow_mnist.py:155); the reference's YAML is a CPU job
        torch.cuda.set_device(hvd.local_rank() % torch.cuda.device_count())
        dev = torch.device("cuda")
    else:
        dev = torch.device("cpu")   # collectives then run over the libmpi shim (hvd/host_backend.py)
    torch.manual_seed(42 + hvd.rank())
    model = MnistConvNet().to(dev)
    # tensorflow_mnist.py:123-130: LR x size for Average; Adasum needs no scaling beyond the local size
    lr_scaler = hvd.size() if not a.use_adasum else (hvd.local_size() if hvd.nccl_built() else 1)
    opt = torch.optim.Adam(model.parameters(), lr=0.001 * lr_scaler)
    opt = hvd.DistributedOptimizer(opt, named_parameters=model.named_parameters(), op=hvd.Adasum if a.use_adasum else hvd.Average)
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    steps = max(1, a.steps // hvd.size())  # tensorflow_mnist.py:146
    # a fixed synthetic "dataset": class = brightest quadrant pattern, learnable
    g = torch.Generator().manual_seed(7)
    protos = torch.randn(10, 784, generator=g).to(dev)
    for i in range(steps):
        y = torch.randint(0, 10, (a.batch_size,), device=dev)
        x = protos[y] + 0.5 * torch.randn(a.batch_size, 784, device=dev)
        opt.zero_grad()
        loss = F.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
        if i % 10 == 0 and hvd.rank() == 0:
            print(f"step {i * hvd.size()} loss = {loss.item():.4f}", flush=True)
    avg = hvd.allreduce(loss.detach().reshape(1), op=hvd.Average)
    if hvd.rank() == 0:
        print(f"final loss (averaged over {hvd.size()} ranks) = {avg.item():.4f}")
        if a.checkpoint_dir:  # rank-0-only checkpoint (tensorflow_mnist.py:159)
            os.makedirs(a.checkpoint_dir, exist_ok=True)
            torch.save(model.state_dict(), os.path.join(a.checkpoint_dir, "model.pt"))
    hvd.shutdown()


if __name__ == "__main__":
    main()
