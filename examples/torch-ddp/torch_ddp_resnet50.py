#!/usr/bin/env python
"""Plain torch.distributed + DistributedDataParallel ResNet-50, bf16 autocast, synthetic ImageNet
(BASELINE.json config #3). Nothing in this file knows about b200mpi: when the MPIJob runs under the
operator, mpirun LD_PRELOADs libb200mpi_nccl.so into every rank and ProcessGroupNCCL's allreduce /
broadcast / allgather calls execute as b200mpi kernels (set B200MPI_ALGO=nccl in the launcher env for the
stock-NCCL baseline)."""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--backend", default="nccl", choices=["nccl", "b200mpi"],
                    help="nccl: ProcessGroupNCCL (under the operator its NCCL calls are LD-injected b200mpi kernels); b200mpi: the c10d "
                         "backend of mpi_operator_b200.parallel.c10d_backend (same runtime, no preloading)")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    if a.backend == "b200mpi":
        sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))
        import mpi_operator_b200.parallel.c10d_backend  # noqa: F401  (registers the backend)
        dist.init_process_group("b200mpi", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    torch.backends.cudnn.benchmark = True
    import torchvision   # stock model definition: this script is what a user of torch DDP already has
    model = getattr(torchvision.models, a.model)(weights=None).cuda().to(memory_format=torch.channels_last)
    ddp = nn.parallel.DistributedDataParallel(model, device_ids=[dev], gradient_as_bucket_view=True)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.01 * world, momentum=0.9)
    x = torch.randn(a.batch_size, 3, 224, 224, device="cuda").contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch_size,), device="cuda")
    loss_fn = nn.CrossEntropyLoss()

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = loss_fn(ddp(x), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        injected = "libb200mpi_nccl" in os.environ.get("LD_PRELOAD", "")
        print(f"torch DDP {a.model} bs {a.batch_size}/GPU x {world} GPUs: {world * a.batch_size * a.steps / (ms.item() * 1e-3):.1f} images/sec "
              f"({ms.item() / a.steps:.2f} ms/step, max over ranks), allreduce backend: {'b200mpi (c10d backend)' if a.backend == 'b200mpi' else 'b200mpi (LD_PRELOAD)' if injected else 'stock NCCL'}, "
              f"loss {loss.item():.3f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
