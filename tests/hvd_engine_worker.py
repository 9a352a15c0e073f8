"""Rank script (run under the native mpirun): the horovod.torch async API on the native hvdcore engine — named tensors in
rank-dependent order, fusion and response-cache counters, ragged allgather, mismatch errors, join(), timeline, the
engine-backed DistributedOptimizer with fp16 compression, stall inspector. Modes: default | stall | stall_shutdown | peer_death | autotune."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import horovod.torch as hvd

mode = sys.argv[1] if len(sys.argv) > 1 else "default"
hvd.init()
r, n = hvd.rank(), hvd.size()
assert hvd.engine_stats().get("world") == n, "the background engine did not start"

if mode == "stall":
    # rank 1 shows up late: rank 0's inspector names the missing rank and tensor, then everything completes
    if r == 1:
        time.sleep(2.5)
    out = hvd.allreduce(torch.ones(3), name="late.tensor", op=hvd.Sum)
    assert torch.equal(out, torch.full((3,), float(n)))
    assert hvd.engine_stats()["stall_warnings"] >= 1 or n == 1
    hvd.shutdown()
    print(f"rank {r}/{n} stall ok", flush=True)
    sys.exit(0)

if mode == "autotune":
    # HOROVOD_AUTOTUNE=1: rank 0 walks the {cycle time} x {fusion threshold} grid, every rank adopts its current candidate in the
    # same cycle (so fused groups always agree and results stay exact), and the search ends on the best-scoring setting
    for it in range(140):
        hs = [hvd.allreduce_async_(torch.full((2000 + 10 * k,), float(r + k)), name=f"at.{k}", op=hvd.Sum) for k in range(12)]
        for k, h in enumerate(hs):
            assert torch.equal(hvd.synchronize(h), torch.full((2000 + 10 * k,), float(n * k + n * (n - 1) / 2))), (it, k)
    st = hvd.engine_stats()
    mine = torch.tensor([st["cycle_time_ms"], float(st["fusion_threshold"])], dtype=torch.float64)
    everyone = hvd.allgather(mine.view(1, 2), name="at.params")
    assert all(torch.equal(everyone[0], everyone[k]) for k in range(n)), everyone        # every rank ended on rank 0's choice
    assert st["autotune"] == "done", st                   # followers notice that rank 0 stopped announcing candidates
    if r == 0:
        assert st["autotune_samples"] >= 21, st
        rows = [l for l in open(os.environ["HOROVOD_AUTOTUNE_LOG"]).read().splitlines() if l and l[0].isdigit()]
        assert len(rows) == 20 and len({tuple(l.split(",")[:2]) for l in rows}) == 20, rows
        assert st["cycle_time_ms"] in (0.5, 1.0, 2.5, 5.0) and st["fusion_threshold"] in (1 << 20, 4 << 20, 16 << 20, 64 << 20, 128 << 20)
    hvd.shutdown()
    print(f"rank {r}/{n} autotune ok", flush=True)
    sys.exit(0)

if mode == "peer_death":
    # rank 1 vanishes without a shutdown: the other engines notice the dead pid in the rendezvous barrier and fail the
    # outstanding handles (HorovodInternalError, transport) instead of hanging
    hvd.barrier()
    if r == 1:
        os._exit(0)
    try:
        hvd.allreduce(torch.ones(3), name="orphaned")
        raise SystemExit("allreduce completed without rank 1")
    except hvd.HorovodInternalError as e:
        assert e.code == -5 and ("died" in str(e) or "aborted" in str(e)), (e.code, str(e))
    print(f"rank {r}/{n} peer death detected", flush=True)
    os._exit(0)

if mode == "stall_shutdown":
    # rank 1 never submits: after HOROVOD_STALL_SHUTDOWN_TIME_SECONDS every engine stops and the waiters get an error
    if r != 1:
        try:
            hvd.allreduce(torch.ones(3), name="never.matched")
            raise SystemExit("allreduce completed although rank 1 never submitted")
        except hvd.HorovodInternalError as e:
            assert e.code == -8 and "stall" in str(e), (e.code, str(e))
    else:
        deadline = time.time() + 30
        while time.time() < deadline:
            try:
                hvd.allreduce_async_(torch.ones(1), name=f"probe.{time.time()}")
            except hvd.HorovodInternalError as e:   # the engine stopped under us as well
                assert e.code == -8, e.code
                break
            time.sleep(0.2)
        else:
            raise SystemExit("rank 1's engine never stopped")
    hvd.shutdown()
    print(f"rank {r}/{n} stall_shutdown ok", flush=True)
    sys.exit(0)

# ---- named async allreduces, submitted in a different order on every rank ----------------------------------------------
timeline = os.environ.get("HVD_TEST_TIMELINE")
if timeline:
    hvd.start_timeline(timeline)
names = [f"layer{i}.weight" for i in range(30)]
for it in range(3):
    g = torch.Generator().manual_seed(17 * it + r)
    order = torch.randperm(len(names), generator=g).tolist()
    tensors = {k: torch.full((5 + k,), float(r + k)) for k in range(len(names))}
    handles = {k: hvd.allreduce_async_(tensors[k], name=names[k], op=hvd.Sum) for k in order}
    assert all(hasattr(h, "wait") for h in handles.values())
    for k in reversed(order):
        out = hvd.synchronize(handles[k])
        assert out is tensors[k] and torch.equal(out, torch.full((5 + k,), float(n * k + n * (n - 1) / 2))), (k, out)
    assert all(hvd.poll(h) for h in handles.values())
st = hvd.engine_stats()
assert st["tensors"] >= 90 and st["fused_groups"] < st["tensors"], st          # tensors travelled fused
assert st["cache_hits"] >= 60, st                                                 # iterations 2 and 3 sent 4-byte ids
if timeline:
    hvd.stop_timeline()
    if r == 0:
        ev = json.load(open(timeline))
        phases = {e.get("name") for e in ev}
        assert {"NEGOTIATE_ALLREDUCE", "ALLREDUCE", "MEMCPY_IN_FUSION_BUFFER", "SHM_ALLREDUCE", "MEMCPY_OUT_FUSION_BUFFER"} <= phases, phases
        rows = {e["pid"]: e["args"]["name"] for e in ev if e.get("name") == "process_name"}
        # ranks that are already past this point may have submitted their next tensors (grp.*) before rank 0 stopped the
        # timeline: only the 30 layer rows of the loop above are complete by construction
        layers = {pid for pid, name in rows.items() if name.startswith("layer")}
        assert "layer3.weight" in rows.values() and len(layers) == 30, (len(layers), sorted(rows.values()))
        for pid in layers:                          # every B has its E
            depth = 0
            for e in (x for x in ev if x["pid"] == pid):
                depth += {"B": 1, "E": -1}.get(e["ph"], 0)
                assert depth >= 0
            assert depth == 0, pid

# grouped allreduce (one negotiation cycle), out-of-place async, integer / half dtypes, scalars and empty tensors
outs = hvd.grouped_allreduce([torch.full((4,), float(r)), torch.full((2, 2), 2.0 * r)], name="grp", op=hvd.Average)
assert torch.allclose(outs[0], torch.full((4,), (n - 1) / 2)) and torch.allclose(outs[1], torch.full((2, 2), float(n - 1)))
src = torch.arange(6, dtype=torch.int64) * (r + 1)
h = hvd.allreduce_async(src, name="oop", op=hvd.Sum)
assert torch.equal(hvd.synchronize(h), torch.arange(6) * (n * (n + 1) // 2)) and torch.equal(src, torch.arange(6) * (r + 1))
assert torch.equal(hvd.allreduce(torch.tensor(float(r)), op=hvd.Max), torch.tensor(float(n - 1)))
assert hvd.allreduce(torch.zeros(0), name="empty").numel() == 0
nc = torch.arange(12, dtype=torch.float32).view(3, 4).t()          # non-contiguous input is reduced and written back in place
hvd.allreduce_(nc, op=hvd.Sum, name="noncontig")
assert torch.equal(nc, n * torch.arange(12, dtype=torch.float32).view(3, 4).t())

# reductions fold in rank order on exactly one rank per slice: the result is bit-identical everywhere (floating-point sums too)
noisy = torch.randn(100003, generator=torch.Generator().manual_seed(99 + r)) * 10.0 ** float(r)
red = hvd.allreduce(noisy, op=hvd.Sum, name="bits")
sig = hvd.allgather(red.view(torch.int32).to(torch.int64).sum().reshape(1, 1), name="bits.sig")
assert all(int(sig[k]) == int(sig[0]) for k in range(n)), sig

# ragged allgather: rank k contributes k + 1 rows
g = hvd.allgather(torch.full((r + 1, 2), float(r)), name="ragged")
assert g.shape == (n * (n + 1) // 2, 2)
assert torch.equal(g, torch.cat([torch.full((k + 1, 2), float(k)) for k in range(n)]))
h = hvd.broadcast_async_(torch.full((3,), float(r)), root_rank=0, name="bc")
assert torch.equal(hvd.synchronize(h), torch.zeros(3))

# a mismatch is an error on every rank, and the engine keeps working afterwards
if n > 1:
    try:
        hvd.allreduce(torch.ones(4 if r else 5), name="mismatch")
        raise SystemExit("mismatched sizes went unnoticed")
    except hvd.HorovodInternalError as e:
        assert "mismatch" in str(e) and e.code == -4, (e.code, str(e))
    # a name may not be reused while it is in flight; rank 0 tries before the others have submitted "dup" at all (they are
    # held back by the barrier), so its first "dup" cannot have completed and the refusal is deterministic
    if r == 0:
        h1 = hvd.allreduce_async_(torch.ones(2), name="dup", op=hvd.Sum)
        try:
            hvd.allreduce_async_(torch.ones(2), name="dup")
            raise SystemExit("a duplicate name in flight was accepted")
        except hvd.HorovodInternalError as e:
            assert e.code == -6 and "dup" in str(e), (e.code, str(e))
    hvd.barrier()
    if r != 0:
        h1 = hvd.allreduce_async_(torch.ones(2), name="dup", op=hvd.Sum)
    assert torch.equal(hvd.synchronize(h1), torch.full((2,), float(n)))
    # a handle dropped without synchronize(): the engine still owns the buffers until the collective has run
    hvd.allreduce_async_(torch.ones(1000), name="dropped")
assert torch.equal(hvd.allreduce(torch.ones(2), op=hvd.Sum), torch.full((2,), float(n)))

# join(): rank k runs k + 1 steps; the ranks that are done contribute zeros
for step in range(r + 1):
    v = hvd.allreduce(torch.ones(3), name=f"uneven.{step}", op=hvd.Sum)
    assert torch.equal(v, torch.full((3,), float(n - step))), (step, v)
assert hvd.join() == n - 1

# engine-backed DistributedOptimizer (Horovod's per-parameter scheme) with fp16 compression == SGD on the mean gradient
torch.manual_seed(0)
model = nn.Sequential(nn.Linear(6, 8), nn.Tanh(), nn.Linear(8, 3))
ref = nn.Sequential(nn.Linear(6, 8), nn.Tanh(), nn.Linear(8, 3))
hvd.broadcast_parameters(model.state_dict(), root_rank=0)
ref.load_state_dict(model.state_dict())
opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9), named_parameters=model.named_parameters(),
                               compression=hvd.Compression.fp16, engine=True)
assert type(opt).__name__ == "_EngineDistributedOptimizer"
ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
for step in range(3):
    xs = [torch.randn(4, 6, generator=torch.Generator().manual_seed(1000 * step + k)) for k in range(n)]
    opt.zero_grad()
    model(xs[r]).pow(2).mean().backward()
    opt.step()
    ropt.zero_grad()
    (sum(ref(x).pow(2).mean() for x in xs) / n).backward()
    ropt.step()
for p, q in zip(model.parameters(), ref.parameters()):
    assert torch.allclose(p, q, rtol=2e-2, atol=2e-3), (p - q).abs().max()      # fp16 on the wire

hvd.barrier()
hvd.shutdown()
try:
    hvd.rank()
    raise SystemExit("still initialised")
except RuntimeError:
    pass
print(f"rank {r}/{n} hvd engine ok", flush=True)
