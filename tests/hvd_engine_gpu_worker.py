"""Rank script (one per GPU, B200MPI_HVD_ENGINE=1): CUDA tensors through the native hvdcore engine — named async allreduces in
a rank-dependent order (fused on the engine's stream into one b200mpi kernel launch per group), broadcast, join."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import horovod.torch as hvd

hvd.init()
r, n = hvd.rank(), hvd.size()
dev = torch.device("cuda", torch.cuda.current_device())
st = hvd.engine_stats()
assert st.get("gpu") is True, st
for it in range(3):
    g = torch.Generator().manual_seed(31 * it + r)
    order = torch.randperm(24, generator=g).tolist()
    ts = {k: torch.full((1000 + 37 * k,), float(r + k), device=dev, dtype=torch.float32 if k % 3 else torch.bfloat16) for k in range(24)}
    hs = {k: hvd.allreduce_async_(ts[k], name=f"g{k}", op=hvd.Sum) for k in order}
    for k in order:
        out = hvd.synchronize(hs[k])
        want = float(n * k + n * (n - 1) / 2)
        assert torch.allclose(out.float(), torch.full_like(out, want).float(), rtol=1e-2), (k, out[:4], want)
st = hvd.engine_stats()
assert st["fused_groups"] < st["tensors"] and st["cache_hits"] >= 48, st
avg = hvd.allreduce(torch.full((5,), float(r), device=dev))
assert torch.allclose(avg, torch.full((5,), (n - 1) / 2, device=dev))
b = hvd.broadcast(torch.full((70000,), float(r), device=dev), root_rank=n - 1, name="bc")
assert torch.equal(b, torch.full((70000,), float(n - 1), device=dev))
i64 = hvd.allreduce(torch.arange(4, device=dev) * (r + 1), op=hvd.Sum)        # integers round-trip through fp32
assert torch.equal(i64, torch.arange(4, device=dev) * (n * (n + 1) // 2))
for step in range(r + 1):
    v = hvd.allreduce(torch.ones(3, device=dev), name=f"uneven.{step}", op=hvd.Sum)
    assert torch.equal(v, torch.full((3,), float(n - step), device=dev)), (step, v)
assert hvd.join() == n - 1
hvd.shutdown()
print(f"rank {r}/{n} hvd engine gpu ok", flush=True)
