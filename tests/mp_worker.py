"""Multi-process (one rank per GPU) correctness worker for the b200mpi runtime.
Launched by tests/test_multigpu.py or directly:
  python tests/mp_launch.py -n 2 tests/mp_worker.py
Exits non-zero on any mismatch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpi_operator_b200.launch.env import rank_info_from_env  # noqa: E402
from mpi_operator_b200.runtime.comm import Communicator  # noqa: E402


def main():
    info = rank_info_from_env()
    ndev = torch.cuda.device_count()
    dev = info.local_rank % ndev
    torch.cuda.set_device(dev)
    comm = Communicator.create(info.rank, info.world_size, dev, info.job_id)
    comm.set_tuning(timeout_ms=10000)
    W, R = comm.world, comm.rank
    if R == 0:
        print(f"[mp_worker] world={W} multicast={comm.has_multicast} tuning={comm.get_tuning()}", flush=True)
    fails = 0

    def gen(seed, n, dtype, r):
        g = torch.Generator(device="cuda").manual_seed(seed * 131 + r)
        return torch.randn(n, device="cuda", generator=g).to(dtype)

    algos = ["oneshot", "twoshot"] + (["nvls"] if comm.has_multicast else [])
    for dtype in (torch.float32, torch.bfloat16, torch.float16):
        tol = {torch.float32: 1e-5, torch.bfloat16: 3e-2, torch.float16: 3e-3}[dtype]
        for n in (1, 1000, 65536 + 8, 3 * 1024 * 1024 + 40):
            xs = [gen(n, n, dtype, r) for r in range(W)]
            want = torch.stack([x.float() for x in xs]).mean(0)
            for algo in algos:
                if algo == "oneshot" and n * xs[0].element_size() > (1 << 20):
                    continue
                out = torch.empty_like(xs[R])
                comm.allreduce(xs[R], out, op="avg", algo=algo)
                torch.cuda.synchronize()
                comm.check_error()
                err = (out.float() - want).abs().max().item()
                if not err <= tol * max(1.0, want.abs().max().item()):
                    print(f"[rank {R}] allreduce {dtype} n={n} {algo}: max err {err}", flush=True)
                    fails += 1
    # symmetric window, in place, all algos
    n = 8 * 1024 * 1024
    win = comm.alloc_window(n * 4)
    for dtype in (torch.float32, torch.bfloat16):
        cnt = n if dtype == torch.float32 else 2 * n
        v = win.tensor(dtype, numel=cnt)
        for algo in algos[1:]:
            xs = [gen(7, cnt, dtype, r) for r in range(W)]
            v.copy_(xs[R])
            torch.cuda.synchronize()
            comm.allreduce_window(win, 0, cnt, dtype, op="sum", algo=algo)
            torch.cuda.synchronize()
            comm.check_error()
            want = torch.stack([x.float() for x in xs]).sum(0)
            err = (v.float() - want).abs().max().item()
            tol = 1e-4 if dtype == torch.float32 else 0.06 * W
            if not err <= tol:
                print(f"[rank {R}] allreduce_window {dtype} {algo}: max err {err}", flush=True)
                fails += 1
    # fused SGD (both modes)
    cnt = 4 * 1024 * 1024
    gwin, pwin = comm.alloc_window(cnt * 4), comm.alloc_window(cnt * 4)
    p0 = gen(11, cnt, torch.float32, 0)
    pv, gv = pwin.tensor(torch.float32, numel=cnt), gwin.tensor(torch.float32, numel=cnt)
    for algo in algos[1:]:
        pv.copy_(p0)
        mom = torch.zeros(comm.slice_elems(cnt, torch.float32), device="cuda")
        ref = p0.clone()
        buf = None
        for step in range(2):
            gs = [gen(20 + step, cnt, torch.float32, r) for r in range(W)]
            gv.copy_(gs[R])
            torch.cuda.synchronize()
            comm.host_barrier()
            comm.allreduce_sgd_window(gwin, 0, pwin, 0, mom, cnt, torch.float32, lr=0.05, momentum_coef=0.9,
                                      weight_decay=1e-4, first_step=(step == 0), algo=algo)
            torch.cuda.synchronize()
            comm.check_error()
            g = torch.stack(gs).mean(0) + 1e-4 * ref
            buf = g.clone() if buf is None else 0.9 * buf + g
            ref = ref - 0.05 * buf
            err = (pv - ref).abs().max().item()
            if not err <= 1e-5:
                print(f"[rank {R}] fused sgd {algo} step {step}: max err {err}", flush=True)
                fails += 1
    # broadcast / allgather / reduce_scatter / alltoall / barrier
    x = gen(3, 100003, torch.float32, R)
    want = gen(3, 100003, torch.float32, W - 1)
    comm.broadcast(x, root=W - 1)
    torch.cuda.synchronize()
    fails += int(not torch.equal(x, want))
    x = gen(4, 50001 * 2, torch.bfloat16, R)
    out = torch.empty(W * x.numel(), device="cuda", dtype=torch.bfloat16)
    comm.allgather(x, out)
    torch.cuda.synchronize()
    fails += int(not torch.equal(out, torch.cat([gen(4, 50001 * 2, torch.bfloat16, r) for r in range(W)])))
    xs = [gen(5, W * 4096, torch.float32, r) for r in range(W)]
    out = torch.empty(4096, device="cuda")
    comm.reduce_scatter(xs[R], out)
    torch.cuda.synchronize()
    fails += int(not torch.allclose(out, torch.stack(xs).sum(0)[R * 4096:(R + 1) * 4096], atol=1e-4))
    out = torch.empty(W * 4096, device="cuda")
    comm.alltoall(xs[R], out)
    torch.cuda.synchronize()
    fails += int(not torch.equal(out, torch.cat([xs[s][R * 4096:(R + 1) * 4096] for s in range(W)])))
    comm.barrier()
    torch.cuda.synchronize()
    comm.check_error()
    comm.host_barrier()
    print(f"[rank {R}] mp_worker done, failures={fails}", flush=True)
    comm.destroy()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
