"""Rank script (run under the native mpirun): the horovod.torch API surface on the CPU backend."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import horovod.torch as hvd

hvd.init()
r, n = hvd.rank(), hvd.size()
assert hvd.is_initialized() and hvd.local_rank() == r and hvd.local_size() == n and not hvd.cuda_built()

# allreduce flavours
t = torch.arange(10, dtype=torch.float32) + r
assert torch.equal(hvd.allreduce(t, op=hvd.Sum), n * torch.arange(10.) + n * (n - 1) / 2)
assert torch.allclose(hvd.allreduce(t), torch.arange(10.) + (n - 1) / 2)                 # Average is the default
assert torch.equal(hvd.allreduce(t, op=hvd.Max), torch.arange(10.) + n - 1)
assert torch.equal(hvd.allreduce(t, op=hvd.Min), torch.arange(10.))
h = (torch.ones(7, dtype=torch.bfloat16) * (r + 1))
assert torch.equal(hvd.allreduce(h, op=hvd.Sum), torch.full((7,), n * (n + 1) / 2, dtype=torch.bfloat16))
i64 = torch.tensor([r, 2 * r], dtype=torch.int64)
assert torch.equal(hvd.allreduce(i64, op=hvd.Sum), torch.tensor([n * (n - 1) // 2, n * (n - 1)]))
assert torch.allclose(hvd.allreduce(t, op=hvd.Sum, prescale_factor=0.5), 0.5 * (n * torch.arange(10.) + n * (n - 1) / 2))
import math
assert torch.allclose(hvd.allreduce(torch.full((3,), float(r + 1)), op=hvd.Product), torch.full((3,), float(math.factorial(n))))
ga, gb = torch.full((2,), float(r)), torch.full((5,), 2.0 * r)
hvd.grouped_allreduce_([ga, gb], op=hvd.Sum, name="grp.inplace")
assert torch.equal(ga, torch.full((2,), n * (n - 1) / 2)) and torch.equal(gb, torch.full((5,), float(n * (n - 1))))
gg = hvd.grouped_allgather([torch.full((1, 2), float(r)), torch.full((2,), float(-r))], name="grp.gather")
assert gg[0].shape == (n, 2) and gg[1].shape == (2 * n,) and float(gg[0][n - 1, 0]) == n - 1 and float(gg[1][-1]) == -(n - 1)
assert torch.equal(hvd.alltoall_async(torch.arange(n, dtype=torch.float32) + 100 * r).wait(), torch.tensor([100. * k + r for k in range(n)]))
grs = hvd.grouped_reducescatter([torch.ones(2 * n), torch.arange(n, dtype=torch.float32)], op=hvd.Sum)
assert torch.equal(grs[0], torch.full((2,), float(n))) and torch.equal(grs[1], torch.tensor([float(n * r)]))
assert hvd.is_homogeneous() and hvd.remove_process_set(object()) is False
# process sets (Horovod >= 0.23): sub-communicators formed by their members; root_rank stays a global rank
assert hvd.global_process_set.size() == n and hvd.global_process_set.rank() == r and hvd.global_process_set.included()
assert torch.equal(hvd.allreduce(torch.ones(2), op=hvd.Sum, process_set=hvd.global_process_set), torch.full((2,), float(n)))
if n >= 4:
    ends = hvd.add_process_set([n - 1, 0])
    evens = hvd.add_process_set(hvd.ProcessSet(range(0, n, 2)))
    assert (ends.process_set_id, evens.process_set_id) == (1, 2) and ends.ranks == [0, n - 1] and evens.size() == (n + 1) // 2
    assert ends.included() == (r in (0, n - 1)) and ends.rank() == ({0: 0, n - 1: 1}.get(r, -1))
    if ends.included():
        v = hvd.allreduce(torch.full((3,), float(r + 1)), op=hvd.Sum, process_set=ends)
        assert torch.equal(v, torch.full((3,), float(n + 1))), v
        assert torch.allclose(hvd.allreduce(torch.full((3,), float(r)), process_set=ends), torch.full((3,), (n - 1) / 2))   # Average over the SET
        bb = hvd.broadcast(torch.full((4,), float(r)), root_rank=n - 1, process_set=ends)                                      # global root rank
        assert torch.equal(bb, torch.full((4,), float(n - 1)))
        gg2 = hvd.allgather(torch.full((1 + ends.rank(), 2), float(r)), process_set=ends)
        assert gg2.shape == (3, 2) and float(gg2[0, 0]) == 0.0 and float(gg2[2, 1]) == n - 1
        hvd.barrier(process_set=ends)
        ga2 = [torch.full((2,), float(r)), torch.ones(3)]
        hvd.grouped_allreduce_(ga2, op=hvd.Sum, process_set=ends)
        assert torch.equal(ga2[0], torch.full((2,), float(n - 1))) and torch.equal(ga2[1], torch.full((3,), 2.0))
    else:
        try:
            hvd.allreduce(torch.ones(1), process_set=ends)
            raise AssertionError("a non-member must not be able to use the set")
        except ValueError:
            pass
    if evens.included():
        s_ev = hvd.allreduce(torch.tensor([float(r)]), op=hvd.Sum, process_set=evens)
        assert float(s_ev) == float(sum(range(0, n, 2)))
        rs_ev = hvd.reducescatter(torch.ones(2 * evens.size()), op=hvd.Sum, process_set=evens)
        assert torch.equal(rs_ev, torch.full((2,), float(evens.size())))
    try:
        hvd.add_process_set([0, n - 1])
        raise AssertionError("duplicate process set")
    except ValueError:
        pass
    assert torch.equal(hvd.allreduce(torch.ones(2), op=hvd.Sum), torch.full((2,), float(n)))     # the global communicator is back in place
    assert hvd.size() == n and hvd.rank() == r
    assert hvd.remove_process_set(ends) and not hvd.remove_process_set(ends) and ends.process_set_id is None
inplace = t.clone()
hvd.allreduce_(inplace, op=hvd.Sum)
assert torch.equal(inplace, n * torch.arange(10.) + n * (n - 1) / 2)

# allgather / broadcast / alltoall / reducescatter
g = hvd.allgather(torch.full((2, 3), float(r)))
assert g.shape == (2 * n, 3) and all(torch.equal(g[2 * k:2 * k + 2], torch.full((2, 3), float(k))) for k in range(n))
rag = hvd.allgather(torch.full((r + 1, 2), float(r)))                  # ragged first dimension (engine: allgatherv; direct path: pad + trim)
assert torch.equal(rag, torch.cat([torch.full((k + 1, 2), float(k)) for k in range(n)]))
assert hvd.allgather(torch.tensor(float(r))).tolist() == [float(k) for k in range(n)]      # 0-dim tensors gather to [n]
b = hvd.broadcast(torch.full((5,), float(r)), root_rank=n - 1)
assert torch.equal(b, torch.full((5,), float(n - 1)))
a2a = hvd.alltoall(torch.arange(n, dtype=torch.float32) + 100 * r)
assert torch.equal(a2a, torch.tensor([100. * k + r for k in range(n)]))
rs = hvd.reducescatter(torch.arange(2 * n, dtype=torch.float32) * (r + 1), op=hvd.Sum)
assert torch.equal(rs, torch.arange(2 * n, dtype=torch.float32).view(n, 2)[r] * (n * (n + 1) / 2))
assert hvd.broadcast_object({"epoch": 3, "who": r} if r == 0 else None, root_rank=0) == {"epoch": 3, "who": 0}
# uneven alltoall: rank r sends (dst + 1) rows to dst, each row tagged with its source
send = torch.cat([torch.full((d + 1, 2), float(r)) for d in range(n)])
got, recv_splits = hvd.alltoall(send, splits=[d + 1 for d in range(n)])
assert recv_splits.tolist() == [r + 1] * n and got.shape == (n * (r + 1), 2)
assert torch.equal(got, torch.cat([torch.full((r + 1, 2), float(src)) for src in range(n)]))
assert hvd.allgather_object({"rank": r, "blob": "x" * (r * 3)}) == [{"rank": k, "blob": "x" * (k * 3)} for k in range(n)]
assert hvd.global_process_set.size() == n and hvd.global_process_set.ranks == list(range(n)) and hvd.global_process_set.included()
assert torch.equal(hvd.allreduce(torch.ones(2), op=hvd.Sum, process_set=hvd.global_process_set), torch.full((2,), float(n)))
try:
    hvd.add_process_set([0, n + 5])
    raise SystemExit("add_process_set must refuse ranks outside the job")
except ValueError:
    pass
hvd.barrier()

# DistributedOptimizer == SGD on the rank-averaged gradient, identical parameters on every rank afterwards
torch.manual_seed(0)
ref = nn.Sequential(nn.Linear(6, 8), nn.Tanh(), nn.Linear(8, 3))
model = nn.Sequential(nn.Linear(6, 8), nn.Tanh(), nn.Linear(8, 3))
torch.manual_seed(100 + r)
for p in model.parameters():
    p.data.normal_()                                # ranks start different...
hvd.broadcast_parameters(model.state_dict(), root_rank=0)   # ...and agree after the broadcast (K3)
w0 = hvd.allgather(model[0].weight.data.flatten()[None])
assert all(torch.equal(w0[0], w0[k]) for k in range(n))
ref.load_state_dict(model.state_dict())
opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9), named_parameters=model.named_parameters(),
                               bucket_bytes=128)
ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
for step in range(3):
    xs = [torch.randn(4, 6, generator=torch.Generator().manual_seed(1000 * step + k)) for k in range(n)]
    opt.zero_grad()
    model(xs[r]).pow(2).mean().backward()
    opt.step()
    ropt.zero_grad()
    (sum(ref(x).pow(2).mean() for x in xs) / n).backward()      # the gradient every rank should have applied
    ropt.step()
for p, q in zip(model.parameters(), ref.parameters()):
    assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), (p - q).abs().max()

# backward_passes_per_step=2, Horovod usage: 2 x backward(), ONE step(); also survives model.zero_grad() (set_to_none=True)
torch.manual_seed(0)
model2 = nn.Sequential(nn.Linear(6, 8), nn.Tanh(), nn.Linear(8, 3))
ref2 = nn.Sequential(nn.Linear(6, 8), nn.Tanh(), nn.Linear(8, 3))
ref2.load_state_dict(model2.state_dict())
opt2 = hvd.DistributedOptimizer(torch.optim.SGD(model2.parameters(), lr=0.1, momentum=0.9), named_parameters=model2.named_parameters(),
                                backward_passes_per_step=2, bucket_bytes=128)
assert type(opt2).__name__ == "_DistributedOptimizer"
ropt2 = torch.optim.SGD(ref2.parameters(), lr=0.1, momentum=0.9)
for step in range(3):
    if step == 1:
        model2.zero_grad()            # drops .grad (set_to_none=True): the optimizer must re-home the fresh gradients
    else:
        opt2.zero_grad()
    ropt2.zero_grad()
    for micro in range(2):
        xs = [torch.randn(4, 6, generator=torch.Generator().manual_seed(5000 + 100 * step + 10 * micro + k)) for k in range(n)]
        model2(xs[r]).pow(2).mean().backward()
        (sum(ref2(x).pow(2).mean() for x in xs) / n / 2).backward()      # mean over ranks and over the two local passes
    opt2.step()
    ropt2.step()
for p, q in zip(model2.parameters(), ref2.parameters()):
    assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), ("passes=2", (p - q).abs().max())
if os.environ.get("B200MPI_HVD_ENGINE") == "0":   # no engine to fall back to: an unsupported argument must be refused loudly
    try:
        hvd.DistributedOptimizer(torch.optim.SGD(nn.Linear(2, 2).parameters(), lr=0.1), compression=hvd.Compression.fp16)
        raise SystemExit("bucket optimizer accepted a compression it cannot apply")
    except ValueError:
        pass

# broadcast_object beyond 16 MiB (a float32 length is only exact below 2^24) and optimizer state as tensors
big = bytes(range(256)) * ((17 << 20) // 256) + b"tail!"
got_big = hvd.broadcast_object(big if r == 0 else None, root_rank=0)
assert len(got_big) == (17 << 20) + 5 and got_big[-5:] == b"tail!" and got_big[:256] == bytes(range(256))
mom_model = nn.Linear(5, 4)
hvd.broadcast_parameters(mom_model.state_dict(), root_rank=0)
mom_opt = torch.optim.SGD(mom_model.parameters(), lr=0.1 * (r + 1), momentum=0.9)
if r == 0:
    mom_model(torch.ones(2, 5)).sum().backward()
    mom_opt.step()
hvd.broadcast_optimizer_state(mom_opt, root_rank=0)
assert mom_opt.param_groups[0]["lr"] == 0.1 and len(mom_opt.state_dict()["state"]) == 2
mb = hvd.allgather(mom_opt.state_dict()["state"][0]["momentum_buffer"].flatten()[None])
assert all(torch.equal(mb[0], mb[k]) for k in range(n)) and mb[0].abs().sum() > 0

# SyncBatchNorm == nn.BatchNorm2d over the concatenated global batch (outputs, input gradients, affine gradients, running stats)
torch.manual_seed(5)
full = torch.randn(4 * n, 3, 5, 5, dtype=torch.float64)
w_out = torch.randn(4 * n, 3, 5, 5, dtype=torch.float64)
ref_bn = nn.BatchNorm2d(3).double()
sbn = hvd.SyncBatchNorm(3).double()
with torch.no_grad():
    for m_ in (ref_bn, sbn):
        m_.weight.copy_(torch.tensor([1.5, -0.5, 2.0]))
        m_.bias.copy_(torch.tensor([0.1, 0.2, -0.3]))
xf = full.clone().requires_grad_(True)
ref_out = ref_bn(xf)
(ref_out * w_out).sum().backward()
xl = full[4 * r:4 * r + 4].clone().requires_grad_(True)
yl = sbn(xl)
(yl * w_out[4 * r:4 * r + 4]).sum().backward()
assert torch.allclose(yl, ref_out[4 * r:4 * r + 4].detach(), rtol=0, atol=1e-12)
assert torch.allclose(xl.grad, xf.grad[4 * r:4 * r + 4], rtol=0, atol=1e-12)
assert torch.allclose(hvd.allreduce(sbn.weight.grad, op=hvd.Sum), ref_bn.weight.grad, atol=1e-9)
assert torch.allclose(sbn.running_mean, ref_bn.running_mean, atol=1e-9) and torch.allclose(sbn.running_var, ref_bn.running_var, atol=1e-9)
sbn.eval()
ref_bn.eval()
assert torch.allclose(sbn(full[:2]), ref_bn(full[:2]), atol=1e-9)

# ElasticSampler: the ranks partition the epoch; after 2 recorded batches the rest is re-partitioned without repeats
ds = list(range(10 * n + 3))
samp = hvd.elastic.ElasticSampler(ds, shuffle=True, seed=3)
mine = list(samp)
everyone = hvd.allgather_object(mine)
assert len(mine) == len(samp) and set(i for part in everyone for i in part) == set(ds)
samp.record_batch(0, 4)
samp.record_batch(1, 4)
done = set(i for part in hvd.allgather_object(sorted(samp.processed_indices)) for i in part)
sd = samp.state_dict()
samp2 = hvd.elastic.ElasticSampler(ds, shuffle=True, seed=3)
samp2.load_state_dict(sd)
assert samp2.epoch == 0 and not (set(samp2) & samp.processed_indices) and len(done) == 8 * n
samp.set_epoch(1)
assert not samp.processed_indices and len(list(samp)) == len(samp)

# Adasum: orthogonal gradients add, parallel gradients average
e = torch.zeros(n)
e[r] = 1.0
assert torch.allclose(hvd.allreduce(e, op=hvd.Adasum), torch.ones(n))
assert torch.allclose(hvd.allreduce(torch.ones(4), op=hvd.Adasum), torch.ones(4))
# against the pairwise tree folded locally over the gathered vectors (with the engine: the native slice-parallel
# implementation, hvd_core.cc host_adasum; without: one allgather + the same tree), several dtypes and an odd length
from mpi_operator_b200.hvd.adasum import adasum_tree


def tree64(vs):   # the same pairwise tree in float64
    vs = [v.double() for v in vs]
    while len(vs) > 1:
        nxt = []
        for i in range(0, len(vs) - 1, 2):
            a, b = vs[i], vs[i + 1]
            d, na, nb = torch.dot(a, b), torch.dot(a, a), torch.dot(b, b)
            nxt.append((1 - d / (2 * na) if na > 0 else 1.0) * a + (1 - d / (2 * nb) if nb > 0 else 1.0) * b)
        if len(vs) % 2:
            nxt.append(vs[-1])
        vs = nxt
    return vs[0]


for dt, tol in ((torch.float32, 1e-5), (torch.float64, 1e-12), (torch.bfloat16, 2e-2)):
    gen = torch.Generator().manual_seed(100 + r)
    v = torch.randn(1237, generator=gen, dtype=torch.float32).to(dt)
    want = tree64(list(hvd.allgather(v.unsqueeze(0)).unbind(0)))
    got = hvd.allreduce(v, op=hvd.Adasum, name=f"adasum.{dt}")
    assert got.dtype == dt and torch.allclose(got.double(), want.double(), rtol=tol, atol=tol), (dt, (got.double() - want.double()).abs().max())
    same = hvd.allgather(got.unsqueeze(0))
    assert all(torch.equal(same[0], same[k]) for k in range(n))          # bit-identical on every rank
assert torch.allclose(hvd.allreduce(torch.full((3,), 2.0), op=hvd.Adasum, prescale_factor=0.5, postscale_factor=4.0), torch.full((3,), 4.0))
# DistributedOptimizer(op=Adasum) works on the model DELTAS of a local optimizer step (Horovod's semantics), so the wrapped
# optimizer keeps its meaning: start + adasum(rank deltas) per tensor
torch.manual_seed(7)
am = nn.Linear(6, 3)
hvd.broadcast_parameters(am.state_dict(), root_rank=0)
start = [p.detach().clone() for p in am.parameters()]
aopt = hvd.DistributedOptimizer(torch.optim.Adam(am.parameters(), lr=0.05), named_parameters=am.named_parameters(), op=hvd.Adasum)
xg = torch.randn(4, 6, generator=torch.Generator().manual_seed(50 + r))
aopt.zero_grad()
am(xg).pow(2).sum().backward()
ref_m = nn.Linear(6, 3)
ref_m.load_state_dict({k: v.clone() for k, v in zip(am.state_dict(), start)})
ropt = torch.optim.Adam(ref_m.parameters(), lr=0.05)
ref_m(xg).pow(2).sum().backward()
ropt.step()
aopt.step()
for p, s0, q in zip(am.parameters(), start, ref_m.parameters()):
    deltas = hvd.allgather((q.detach() - s0).reshape(1, -1))
    want = s0 + adasum_tree(list(deltas.unbind(0))).view_as(s0)
    assert torch.allclose(p.detach(), want, rtol=1e-5, atol=1e-6), (p.detach() - want).abs().max()

# elastic state: commit / restore / sync
state = hvd.elastic.TorchState(model=model, optimizer=None, epoch=r, batch=7)
state.sync()
assert state.epoch == 0
state.epoch = 5
state.commit()
state.epoch = 9
state.restore()
assert state.epoch == 5
# commit snapshots by VALUE: in-place changes after the commit are rolled back (objects, model, optimizer momentum)
em = nn.Linear(3, 2)
eo = torch.optim.SGD(em.parameters(), lr=0.1, momentum=0.9)
em(torch.ones(1, 3)).sum().backward()
eo.step()
hist = [1, 2]
samp3 = hvd.elastic.ElasticSampler(list(range(8 * n)), shuffle=False)
st2 = hvd.elastic.TorchState(model=em, optimizer=eo, hist=hist, sampler=samp3)
st2.commit()
w_commit = em.weight.detach().clone()
m_commit = eo.state_dict()["state"][0]["momentum_buffer"].clone()
hist.append(3)
em(torch.ones(1, 3)).sum().backward()
eo.step()
samp3.record_batch(0, 2)
assert not torch.equal(em.weight, w_commit)
st2.restore()
assert st2.hist == [1, 2] and torch.equal(em.weight, w_commit)
assert torch.equal(eo.state_dict()["state"][0]["momentum_buffer"], m_commit)
assert st2.sampler is samp3 and not samp3.processed_indices          # same object, rolled back through load_state_dict
# sync() unites the processed indices of all ranks and keeps every rank on its own shard
samp3.record_batch(0, 2)
mine_done = set(samp3.processed_indices)
st2.sync()
assert st2.sampler is samp3 and samp3.rank == r and samp3.num_replicas == n
assert mine_done <= samp3.processed_indices and len(samp3.processed_indices) == 2 * n
assert not (set(samp3.indices) & samp3.processed_indices)
print(f"rank {r}/{n} hvd cpu ok", flush=True)
hvd.shutdown()
