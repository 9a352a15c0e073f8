"""Rank script: torch.distributed with backend "b200mpi" (mpi_operator_b200/parallel/c10d_backend.py) on CPU tensors - the
collectives torch users call and DDP training against a single-process run on the concatenated batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn as nn

import mpi_operator_b200.parallel.c10d_backend  # noqa: F401  (registers the backend)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    on_gpu = torch.cuda.is_available() and os.environ.get("B200MPI_C10D_DEVICE", "") != "cpu"
    if on_gpu:   # CUDA tensors: the same process group drives runtime.comm.Communicator (NVSwitch kernels on the current stream)
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count())
        torch.set_default_device("cuda")
    dist.init_process_group("b200mpi", rank=rank, world_size=world)
    assert dist.get_backend() == "b200mpi" and dist.get_rank() == rank and dist.get_world_size() == world
    fails = 0
    for dtype in (torch.float32, torch.float64, torch.bfloat16, torch.int64):
        x = torch.full((1003,), rank + 1, dtype=dtype)
        dist.all_reduce(x)
        fails += int(not torch.equal(x, torch.full_like(x, world * (world + 1) // 2)))
    x = torch.full((10,), float(rank + 1))
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
    fails += int(not torch.allclose(x, torch.full_like(x, (world + 1) / 2)))
    x = torch.arange(6, dtype=torch.float32) * (rank + 1)
    dist.all_reduce(x, op=dist.ReduceOp.MAX)
    fails += int(not torch.equal(x, torch.arange(6, dtype=torch.float32) * world))
    x = torch.full((4,), 2.0)
    dist.all_reduce(x, op=dist.ReduceOp.PRODUCT)
    fails += int(not torch.equal(x, torch.full((4,), 2.0 ** world)))
    nc = torch.arange(12, dtype=torch.float32).view(3, 4).t()          # non-contiguous input
    nc = nc * (rank + 1)
    dist.all_reduce(nc)
    fails += int(not torch.equal(nc, torch.arange(12, dtype=torch.float32).view(3, 4).t() * (world * (world + 1) // 2)))
    b = torch.arange(1001, dtype=torch.float64) if rank == world - 1 else torch.zeros(1001, dtype=torch.float64)
    dist.broadcast(b, src=world - 1)
    fails += int(not torch.equal(b, torch.arange(1001, dtype=torch.float64)))
    g = [torch.empty(5, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(g, torch.full((5,), rank, dtype=torch.int32))
    fails += int([int(t[0]) for t in g] != list(range(world)))
    flat = torch.empty(world * 3)
    dist.all_gather_into_tensor(flat, torch.full((3,), float(rank)))
    fails += int(flat.view(world, 3)[:, 0].tolist() != [float(r) for r in range(world)])
    out = torch.empty(4)
    dist.reduce_scatter_tensor(out, torch.arange(4 * world, dtype=torch.float32) * (rank + 1))
    want = torch.arange(4 * world, dtype=torch.float32).view(world, 4)[rank] * (world * (world + 1) // 2)
    fails += int(not torch.equal(out, want))
    r = torch.ones(9) * (rank + 1)
    dist.reduce(r, dst=0)
    if rank == 0:
        fails += int(not torch.equal(r, torch.full((9,), float(world * (world + 1) // 2))))
    a2a_in = torch.arange(world, dtype=torch.float32) + 100 * rank
    a2a_out = torch.empty(world)
    dist.all_to_all_single(a2a_out, a2a_in)
    fails += int(a2a_out.tolist() != [100.0 * k + rank for k in range(world)])
    objs = [None] * world
    dist.all_gather_object(objs, {"rank": rank})
    fails += int([o["rank"] for o in objs] != list(range(world)))
    # rooted and point-to-point operations
    mine = torch.arange(4, dtype=torch.float32) + 10 * rank
    got = [torch.empty(4) for _ in range(world)] if rank == world - 1 else None
    dist.gather(mine, got, dst=world - 1)
    if rank == world - 1:
        fails += int([t.tolist() for t in got] != [[10.0 * r + k for k in range(4)] for r in range(world)])
    part = torch.empty(3, dtype=torch.int64)
    dist.scatter(part, [torch.full((3,), 7 * r, dtype=torch.int64) for r in range(world)] if rank == 0 else None, src=0)
    fails += int(part.tolist() != [7 * rank] * 3)
    outs = [torch.empty(2) for _ in range(world)]
    dist.all_to_all(outs, [torch.full((2,), float(100 * rank + r)) for r in range(world)])
    fails += int([t.tolist() for t in outs] != [[float(100 * r + rank)] * 2 for r in range(world)])
    if world > 1:
        outs = [torch.empty(r + 1) for r in range(world)]                # uneven: rank r sends (rank + 1) elements to everyone
        dist.all_to_all(outs, [torch.full((rank + 1,), float(rank)) for _ in range(world)])
        fails += int([t.tolist() for t in outs] != [[float(r)] * (r + 1) for r in range(world)])
        nxt, prv = (rank + 1) % world, (rank - 1) % world
        tok = torch.full((1000,), float(rank))
        inc = torch.empty(1000)
        if rank % 2 == 0:
            dist.send(tok, nxt, tag=3)
            dist.recv(inc, prv, tag=3)
        else:
            dist.recv(inc, prv, tag=3)
            dist.send(tok, nxt, tag=3)
        fails += int(not torch.equal(inc, torch.full((1000,), float(prv))))
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, tok, nxt), dist.P2POp(dist.irecv, inc.zero_(), prv)])
        for q in reqs:
            q.wait()
        fails += int(not torch.equal(inc, torch.full((1000,), float(prv))))
    dist.barrier()
    # DDP training: identical to a single-process run on the concatenated batch
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 5))
    ref = nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 5))
    ref.load_state_dict(model.state_dict())
    if rank != 0:
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)                                              # DDP must broadcast rank 0's parameters
    ddp = nn.parallel.DistributedDataParallel(model)
    opt, ropt = torch.optim.SGD(ddp.parameters(), lr=0.1, momentum=0.9), torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
    gen = torch.Generator(device="cuda" if on_gpu else "cpu").manual_seed(5)
    for _ in range(4):
        xs = torch.randn(world * 8, 16, generator=gen)
        ys = torch.randint(0, 5, (world * 8,), generator=gen)
        opt.zero_grad()
        nn.functional.cross_entropy(ddp(xs[rank * 8:(rank + 1) * 8]), ys[rank * 8:(rank + 1) * 8]).backward()
        opt.step()
        ropt.zero_grad()
        nn.functional.cross_entropy(ref(xs), ys).backward()
        ropt.step()
    for p, q in zip(model.parameters(), ref.parameters()):
        fails += int(not torch.allclose(p, q, rtol=1e-4, atol=1e-5))
    if world > 2:                                                        # sub-groups get communicators of their own
        low, high = dist.new_group([0, 1]), dist.new_group(list(range(1, world))[::-1])   # overlapping groups
        z = torch.full((5,), float(rank + 1))
        if rank in (0, 1):
            dist.all_reduce(z, group=low)
            fails += int(not torch.equal(z, torch.full((5,), 3.0)))
        z = torch.full((5,), float(rank + 1))
        if rank >= 1:
            dist.all_reduce(z, group=high)
            fails += int(not torch.equal(z, torch.full((5,), float(sum(range(2, world + 1))))))
            got = [torch.empty(2) for _ in range(world - 1)]
            dist.all_gather(got, torch.full((2,), float(rank)), group=high)
            fails += int([int(t[0]) for t in got] != list(range(1, world)))           # torch orders a group by global rank
            bb = torch.full((3,), float(rank))
            dist.broadcast(bb, src=world - 1, group=high)
            fails += int(not torch.equal(bb, torch.full((3,), float(world - 1))))
    same = dist.new_group(list(range(world)))                            # a group of every rank works
    y = torch.ones(3)
    dist.all_reduce(y, group=same)
    fails += int(not torch.equal(y, torch.full((3,), float(world))))
    print(f"[rank {rank}] c10d b200mpi backend failures={fails} device={'cuda' if on_gpu else 'cpu'}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
