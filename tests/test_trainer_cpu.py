"""DataParallelTrainer bucket / optimizer / bf16-shadow / accumulation logic on CPU.

The trainer normally runs on a GPU against the native runtime; here the Communicator is replaced
by an in-memory fake whose fused allreduce+SGD does in PyTorch exactly what ``k_allreduce_sgd``
does per element (csrc/kernels/collectives.cu), so the host-side logic — flat layout, reverse-order
buckets, hook bookkeeping, the bf16 parameter shadow, ``no_sync`` accumulation — is verified against
``torch.optim.SGD`` without a device. GPU numerics of the kernels themselves: test_collectives_gpu.py,
test_trainer_gpu.py."""
import copy
import os

import pytest
import torch
import torch.nn as nn

from mpi_operator_b200.parallel.data_parallel import DataParallelTrainer


class FakeWindow:
    def __init__(self, wid, nbytes):
        self.id, self.nbytes = wid, nbytes
        self.buf = torch.zeros(nbytes, dtype=torch.uint8)

    def tensor(self, dtype=None, rank=-1, offset=0, numel=None):
        dtype = dtype or torch.uint8
        esz = torch.empty((), dtype=dtype).element_size()
        if numel is None:
            numel = (self.nbytes - offset) // esz
        return self.buf[offset:offset + numel * esz].view(dtype)


class FakeComm:
    device = "cpu"
    rank, world = 0, 1

    def __init__(self):
        self.windows, self.launch_count, self.hyper, self.calls = [], 0, None, []

    def alloc_window(self, nbytes):
        w = FakeWindow(len(self.windows), nbytes)
        self.windows.append(w)
        return w

    def slice_elems(self, count, dtype):
        return count

    def set_hyper(self, t):
        self.hyper = t

    def broadcast(self, tensor, root=0, stream=None):
        pass

    def host_barrier(self):
        pass

    def allreduce_window(self, win, offset, count, dtype, op="sum", scale=None, algo=None, stream=None):
        self.launch_count += 1  # world 1: avg of one rank is the identity

    def allreduce_sgd_window(self, grad_win, grad_off, param_win, param_off, momentum, count, grad_dtype, lr,
                             momentum_coef=0.0, weight_decay=0.0, nesterov=False, first_step=False, scale=None,
                             lowp_win=None, lowp_off=0, algo=None, stream=None):
        self.launch_count += 1
        self.calls.append((grad_off, count, lowp_off if lowp_win is not None else None))
        if self.hyper is not None:
            lr, momentum_coef, weight_decay = (float(v) for v in self.hyper)
        g = grad_win.tensor(torch.float32, offset=grad_off, numel=count) * (scale if scale is not None else 1.0 / self.world)
        p = param_win.tensor(torch.float32, offset=param_off, numel=count)
        g = g + weight_decay * p
        momentum.copy_(g if first_step else momentum_coef * momentum + g)
        p.sub_(lr * (g + momentum_coef * momentum if nesterov else momentum))
        if lowp_win is not None:
            lowp_win.tensor(torch.bfloat16, offset=lowp_off, numel=count).copy_(p)


def small_cnn():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(),
                         nn.Conv2d(8, 16, 3, padding=1, bias=True), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                         nn.Linear(16, 10))


def batches(n, seed=1):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(4, 3, 8, 8, generator=g), torch.randint(0, 10, (4,), generator=g)) for _ in range(n)]


def reference_steps(model, data, lr, mu, wd, nesterov=False, autocast=False, accumulate=1):
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=mu, weight_decay=wd, nesterov=nesterov)
    lossf = nn.CrossEntropyLoss()
    losses = []
    for i, (x, y) in enumerate(data):
        if i % accumulate == 0:
            opt.zero_grad()
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            loss = lossf(model(x), y)
        loss.backward()
        losses.append(float(loss.detach()))
        if i % accumulate == accumulate - 1:
            opt.step()
    return losses


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("nesterov", [False, True])
def test_fp32_trainer_matches_torch_sgd_across_buckets(fused, nesterov):
    ref = small_cnn()
    model = copy.deepcopy(ref)
    data = batches(4)
    want = reference_steps(ref, data, 0.1, 0.9, 1e-3, nesterov)
    comm = FakeComm()
    tr = DataParallelTrainer(model, nn.CrossEntropyLoss(), comm, lr=0.1, momentum=0.9, weight_decay=1e-3, nesterov=nesterov,
                             bucket_bytes=1024, autocast_dtype=None, channels_last=False, cuda_graph=False, fused_optimizer=fused)
    assert len(tr.state.buckets) > 2
    got = [float(tr.step(x, y)) for x, y in data]
    assert got == pytest.approx(want, rel=1e-5)
    for a, b in zip(model.parameters(), ref.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    if fused:  # one fused kernel per bucket per step, regions tile the flat buffer in order
        per_step = comm.calls[:len(tr.state.buckets)]
        assert sorted(c[0] for c in per_step) == [b.start * 4 for b in tr.state.buckets]
        assert tr.launches_per_step == len(tr.state.buckets)


def test_bf16_params_shadow_matches_autocast_training():
    ref = small_cnn()
    model = copy.deepcopy(ref)
    data = batches(4)
    want = reference_steps(ref, data, 0.05, 0.9, 1e-4, autocast=True)
    comm = FakeComm()
    tr = DataParallelTrainer(model, nn.CrossEntropyLoss(), comm, lr=0.05, momentum=0.9, weight_decay=1e-4, bucket_bytes=2048,
                             autocast_dtype=torch.bfloat16, channels_last=False, cuda_graph=False, bf16_params=True)
    st = tr.state
    mats = [p for p in model.parameters() if p.dim() >= 2]
    assert mats and all(p.dtype == torch.bfloat16 for p in mats)
    assert all(p.dtype == torch.float32 for p in model.parameters() if p.dim() < 2)
    got = [float(tr.step(x, y)) for x, y in data]
    # same bf16 weights in forward, same bf16 gradients widened to fp32: the trajectories coincide
    assert got == pytest.approx(want, rel=2e-3)
    masters = list(st.master_state().values())
    for m, p, r in zip(masters, model.parameters(), ref.parameters()):
        torch.testing.assert_close(m, r, rtol=2e-3, atol=2e-4)
        if p.dim() >= 2:
            assert torch.equal(p.detach(), m.to(torch.bfloat16))  # shadow == bf16(master), written by the fused kernel
            assert p.grad is None                                  # stolen bf16 gradients were gathered and released
    assert all(c[2] == c[0] // 2 for c in comm.calls)              # lowp region offset tracks the bucket offset


def test_bf16_params_unfused_optimizer_refreshes_shadow():
    model = small_cnn()
    comm = FakeComm()
    tr = DataParallelTrainer(model, nn.CrossEntropyLoss(), comm, lr=0.05, momentum=0.9, autocast_dtype=torch.bfloat16,
                             channels_last=False, cuda_graph=False, fused_optimizer=False, bf16_params=True)
    before = [p.detach().clone() for p in model.parameters()]
    for x, y in batches(2):
        tr.step(x, y)
    for p, b, m in zip(model.parameters(), before, tr.state.master_state().values()):
        assert not torch.equal(p.detach(), b)
        if p.dim() >= 2:
            assert torch.equal(p.detach(), m.to(torch.bfloat16))


@pytest.mark.parametrize("bf16", [False, True])
def test_no_sync_accumulates_then_applies_one_update(bf16):
    ref = small_cnn()
    model = copy.deepcopy(ref)
    data = batches(4)
    reference_steps(ref, data, 0.1, 0.9, 0.0, autocast=bf16, accumulate=2)
    comm = FakeComm()
    tr = DataParallelTrainer(model, nn.CrossEntropyLoss(), comm, lr=0.1, momentum=0.9, bucket_bytes=4096,
                             autocast_dtype=torch.bfloat16 if bf16 else None, channels_last=False, cuda_graph=False, bf16_params=bf16)
    for i, (x, y) in enumerate(data):
        if i % 2 == 0:
            with tr.no_sync():
                n0 = comm.launch_count
                tr.step(x, y)
                assert comm.launch_count == n0  # no collective, no update inside no_sync
        else:
            tr.step(x, y)
    tol = dict(rtol=1e-2, atol=1e-3) if bf16 else dict(rtol=1e-5, atol=1e-6)
    for m, r in zip(tr.state.master_state().values(), ref.parameters()):
        torch.testing.assert_close(m, r, **tol)


def test_env_switch_and_dtype_guard(monkeypatch):
    monkeypatch.setenv("B200MPI_BF16_PARAMS", "1")
    tr = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), lr=0.1, channels_last=False, cuda_graph=False)
    assert tr.bf16_params and tr.state.lowp_win is not None
    tr = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), lr=0.1, autocast_dtype=None, channels_last=False,
                             cuda_graph=False)
    assert not tr.bf16_params and tr.state.lowp_win is None  # no shadow without bf16 compute
    monkeypatch.delenv("B200MPI_BF16_PARAMS")
    assert not DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), lr=0.1, channels_last=False,
                                   cuda_graph=False).bf16_params


def test_graph_replay_accounting_folds_recaptures():
    """CUDA-graph replays launch collectives without the host: the trainer reports replays x per-capture counts to
    Communicator.stats() through a stat source, and keeps the totals of graphs that were re-captured since."""
    sources = []

    class CountingComm(FakeComm):
        def add_stat_source(self, fn):
            sources.append(fn)

    tr = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), CountingComm(), lr=0.1, channels_last=False, cuda_graph=False,
                             autocast_dtype=None)
    assert len(sources) == 1 and sources[0]() == []
    tr._graph_ops = [{"op": "allreduce_sgd", "algo": "nvls", "calls": 9, "bytes": 1000}]
    tr._replays = 5
    assert sources[0]() == [{"op": "allreduce_sgd", "algo": "nvls", "calls": 45, "bytes": 5000}]
    tr._folded_ops = tr._replayed_ops()                 # what _capture() does before recording a new graph
    tr._graph_ops, tr._replays = [{"op": "allreduce_sgd", "algo": "nvls", "calls": 3, "bytes": 10},
                                  {"op": "allreduce", "algo": "oneshot", "calls": 1, "bytes": 4}], 2
    got = {(o["op"], o["algo"]): (o["calls"], o["bytes"]) for o in sources[0]()}
    assert got == {("allreduce_sgd", "nvls"): (51, 5020), ("allreduce", "oneshot"): (2, 8)}


def test_prefetch_then_step_without_arguments_matches_step_with_arguments():
    """Input pipeline API: ``prefetch(x, y); step()`` trains on exactly the batch ``step(x, y)`` would have used (on CPU the
    copy is synchronous; on a GPU with async_h2d it goes through the copy stream and staging buffers)."""
    data = batches(3)
    a = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), lr=0.1, channels_last=False, cuda_graph=False, autocast_dtype=None)
    b = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), lr=0.1, channels_last=False, cuda_graph=False, autocast_dtype=None,
                            async_h2d=True)   # no CUDA here: the flag must degrade to the synchronous path
    b.prefetch(*data[0])
    for i, (x, y) in enumerate(data):
        la = float(a.step(x, y))
        lb = float(b.step())
        if i + 1 < len(data):
            b.prefetch(*data[i + 1])
        assert la == lb
    with pytest.raises(RuntimeError, match="prefetch"):
        b.step()
    for p, q in zip(a.model.parameters(), b.model.parameters()):
        assert torch.equal(p, q)


@pytest.mark.parametrize("fused,bf16", [(True, False), (True, True), (False, False)])
def test_checkpoint_resume_continues_bit_for_bit(fused, bf16):
    """state_dict() / load_state_dict(): masters, momentum (re-sharded), BN buffers and hyper-parameters; a resumed trainer takes
    exactly the steps the original would have taken."""
    data = batches(6, seed=4)
    kw = dict(lr=0.05, momentum=0.9, weight_decay=1e-4, autocast_dtype=torch.bfloat16 if bf16 else None, cuda_graph=False,
              bucket_bytes=2048, fused_optimizer=fused, bf16_params=bf16)
    a = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), **kw)
    for x, y in data[:3]:
        a.step(x, y)
    import io
    blob = io.BytesIO()
    torch.save(a.state_dict(), blob)                       # what a rank-0 checkpoint file holds
    for x, y in data[3:]:
        a.step(x, y)
    sd = torch.load(io.BytesIO(blob.getvalue()), weights_only=False)
    assert sd["format"].startswith("b200mpi.DataParallelTrainer") and set(sd["model"]) == set(sd["momentum"])
    assert all(t.dtype == torch.float32 and t.device.type == "cpu" for t in sd["model"].values())   # fp32 masters even with bf16 leaves
    torch.manual_seed(99)                                  # a differently initialised model: everything must come from the checkpoint
    fresh = small_cnn()
    with torch.no_grad():
        for p in fresh.parameters():
            p.add_(1.0)
    b = DataParallelTrainer(fresh, nn.CrossEntropyLoss(), FakeComm(), **dict(kw, lr=0.5, momentum=0.0))
    b.load_state_dict(sd)
    assert float(b.hyper[0]) == pytest.approx(0.05) and float(b.hyper[1]) == pytest.approx(0.9)
    for x, y in data[3:]:
        b.step(x, y)
    for (n, p), (_, q) in zip(a.state.master_state().items(), b.state.master_state().items()):
        assert torch.equal(p, q), n
    for (n, u), (_, v) in zip(a.model.named_buffers(), b.model.named_buffers()):
        assert torch.equal(u, v), n
    # a checkpoint written by the fused optimizer loads into the unfused one and vice versa (same momentum semantics)
    c = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), **dict(kw, fused_optimizer=not fused))
    c.load_state_dict(sd)
    for x, y in data[3:]:
        c.step(x, y)
    for p, q in zip(a.state.master_state().values(), c.state.master_state().values()):
        torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        b.load_state_dict({"format": "something else"})


def test_momentum_shards_are_rebuilt_for_another_world_size():
    """The fused optimizer keeps 1/world of each bucket's momentum per rank: loading a world-1 checkpoint into rank r of a
    world-4 trainer must give it exactly slice r of every bucket."""
    class Comm4(FakeComm):
        world = 4

        def __init__(self, rank):
            super().__init__()
            self.rank = rank

        def slice_elems(self, count, dtype):
            nvec = (count + 3) // 4
            return (nvec + self.world - 1) // self.world * 4

    src = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), lr=0.05, momentum=0.9, autocast_dtype=None,
                              cuda_graph=False, bucket_bytes=2048)
    for x, y in batches(2, seed=5):
        src.step(x, y)
    sd = src.state_dict()
    full = torch.cat([b.momentum[:b.numel] for b in src.state.buckets])
    assert float(full.abs().sum()) > 0
    pieces = {}
    for r in range(4):
        t = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), Comm4(r), lr=0.05, momentum=0.9, autocast_dtype=None,
                                cuda_graph=False, bucket_bytes=2048)
        t.load_state_dict(sd)
        pieces[r] = [b.momentum.clone() for b in t.state.buckets]
        assert len(t.state.buckets) == len(src.state.buckets)
    for k, b in enumerate(src.state.buckets):
        per = pieces[0][k].numel()
        rebuilt = torch.cat([pieces[r][k] for r in range(4)])[:b.numel]
        assert per * 4 >= b.numel and torch.equal(rebuilt, b.momentum[:b.numel])


# ---- several ranks as threads: the trainer's host logic with world > 1 (shard sizes, offsets, parameter broadcast, bf16 shadow
#      pushed to every rank, small tail bucket) against torch SGD on the GLOBAL batch ----------------------------------------
import threading


class ThreadWorld:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.windows = {}          # window id -> {rank: FakeWindow}
        self.lock = threading.Lock()


class ThreadComm(FakeComm):
    """What the runtime's kernels do, with threads for ranks: the owner of slice r averages the ranks' gradients for that
    slice, updates its fp32 parameters (momentum sharded 1/world) and writes the updated slice - and its bf16 shadow - into
    EVERY rank's window (k_allreduce_sgd, csrc/kernels/collectives.cu)."""

    def __init__(self, tw, rank):
        super().__init__()
        self.tw, self.rank, self.world = tw, rank, tw.world

    def alloc_window(self, nbytes):
        w = super().alloc_window(nbytes)
        with self.tw.lock:
            self.tw.windows.setdefault(w.id, {})[self.rank] = w
        return w

    def slice_elems(self, count, dtype):
        nvec = (count + 3) // 4
        return (nvec + self.world - 1) // self.world * 4

    def host_barrier(self):
        self.tw.barrier.wait(timeout=60)

    def broadcast(self, tensor, root=0, stream=None):       # the trainer broadcasts its flat parameter window
        self.tw.barrier.wait(timeout=60)
        if self.rank != root:
            src = self.tw.windows[0][root].tensor(torch.float32)
            tensor.copy_(src[:tensor.numel()])
        self.tw.barrier.wait(timeout=60)

    def allreduce_sgd_window(self, grad_win, grad_off, param_win, param_off, momentum, count, grad_dtype, lr,
                             momentum_coef=0.0, weight_decay=0.0, nesterov=False, first_step=False, scale=None,
                             lowp_win=None, lowp_off=0, algo=None, stream=None):
        self.launch_count += 1
        if self.hyper is not None:
            lr, momentum_coef, weight_decay = (float(v) for v in self.hyper)
        W, r = self.world, self.rank
        per = self.slice_elems(count, torch.float32)
        lo, hi = min(r * per, count), min((r + 1) * per, count)
        self.tw.barrier.wait(timeout=60)                     # every rank's gradients for this bucket are in place
        if hi > lo:
            g = sum(self.tw.windows[grad_win.id][k].tensor(torch.float32, offset=grad_off, numel=count)[lo:hi] for k in range(W))
            g = g * (scale if scale is not None else 1.0 / W)
            p = param_win.tensor(torch.float32, offset=param_off, numel=count)[lo:hi].clone()
            g = g + weight_decay * p
            m = momentum[:hi - lo]
            m.copy_(g if first_step else momentum_coef * m + g)
            p = p - lr * (g + momentum_coef * m if nesterov else m)
            for k in range(W):                                # the all-gather half: updated parameters to every rank
                self.tw.windows[param_win.id][k].tensor(torch.float32, offset=param_off, numel=count)[lo:hi] = p
                if lowp_win is not None:
                    self.tw.windows[lowp_win.id][k].tensor(torch.bfloat16, offset=lowp_off, numel=count)[lo:hi] = p.to(torch.bfloat16)
        self.tw.barrier.wait(timeout=60)                     # every slice has landed everywhere


def plain_cnn():   # no BatchNorm: per-rank batch statistics would make the 2-rank run differ from the global-batch reference
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 16, 3, padding=1), nn.ReLU(),
                         nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(16, 10))


@pytest.mark.parametrize("world,full", [(2, False), (2, True), (4, True)])
def test_multi_rank_trainer_matches_sgd_on_the_global_batch(world, full, monkeypatch):
    """`full` = bench.py's second multi-GPU candidate: bf16 weight shadow + a small tail bucket."""
    if full:
        monkeypatch.setenv("B200MPI_TAIL_BUCKET_BYTES", "1536")
    g = torch.Generator().manual_seed(11)
    data = [(torch.randn(4 * world, 3, 8, 8, generator=g), torch.randint(0, 10, (4 * world,), generator=g)) for _ in range(4)]
    ref = plain_cnn()
    base = copy.deepcopy(ref)
    want = reference_steps(ref, data, 0.05, 0.9, 1e-4, autocast=full)
    tw = ThreadWorld(world)
    out, errs = {}, []

    def run(r):
        try:
            torch.manual_seed(100 + r)
            model = copy.deepcopy(base)
            if r != 0:                                        # only rank 0's parameters may survive the initial broadcast
                with torch.no_grad():
                    for p in model.parameters():
                        p.add_(0.5)
            tr = DataParallelTrainer(model, nn.CrossEntropyLoss(), ThreadComm(tw, r), lr=0.05, momentum=0.9, weight_decay=1e-4,
                                     bucket_bytes=2048, autocast_dtype=torch.bfloat16 if full else None, channels_last=False,
                                     cuda_graph=False, bf16_params=full)
            losses = [float(tr.step(x[4 * r:4 * r + 4], y[4 * r:4 * r + 4])) for x, y in data]
            out[r] = (losses, [m.clone() for m in tr.state.master_state().values()], tr,
                      None if tr.state.flat_lowp is None else tr.state.flat_lowp.clone())
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))
            tw.barrier.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(120) for t in ts]
    assert not errs and len(out) == world, errs
    tol = dict(rtol=3e-3, atol=3e-4) if full else dict(rtol=1e-5, atol=1e-6)
    mean_loss = [sum(out[r][0][i] for r in range(world)) / world for i in range(len(data))]
    assert mean_loss == pytest.approx(want, rel=3e-3 if full else 1e-5)
    for r in range(world):
        for m, q in zip(out[r][1], ref.parameters()):
            torch.testing.assert_close(m, q.detach(), **tol)
        for m, m0 in zip(out[r][1], out[0][1]):
            assert torch.equal(m, m0)                         # bit-identical parameters on every rank
        tr = out[r][2]
        if full:
            assert len(tr.state.buckets) >= 3 and tr.state.buckets[-1].numel * 4 <= 1536
            assert torch.equal(out[r][3], tr.state.flat_param.to(torch.bfloat16))   # the check bench.py applies after a run
        for b in tr.state.buckets:
            assert b.momentum.numel() == tr.comm.slice_elems(b.numel, torch.float32)


def test_model_zoo_names_of_tf_cnn_benchmarks_build_and_train_one_step():
    """--model names of tf_cnn_benchmarks beyond the in-tree families resolve to torchvision definitions (same trainer, unfused
    BN); the *_v1.5 names are the in-tree ResNets; unknown names say what exists."""
    from mpi_operator_b200.models import build_model, model_names
    assert {"resnet101", "inception3", "googlenet", "mobilenet", "vgg16", "alexnet", "trivial"} <= set(model_names())
    assert type(build_model("resnet50_v1.5")).__name__ == "ResNet" and type(build_model("resnet50_v1.5")) is type(build_model("resnet50"))
    for name in ("mobilenet", "googlenet"):
        tr = DataParallelTrainer(build_model(name), nn.CrossEntropyLoss(), FakeComm(), lr=0.01, momentum=0.9, autocast_dtype=None,
                                 channels_last=False, cuda_graph=False)
        x, y = torch.randn(2, 3, 64, 64), torch.randint(0, 1000, (2,))
        l0 = float(tr.step(x, y))
        assert l0 == l0 and tr.launches_per_step == len(tr.state.buckets)
    with pytest.raises(KeyError) as e:
        build_model("not-a-model")
    assert "resnet101" in str(e.value)


def test_evaluate_is_forward_only_and_uses_running_statistics():
    tr = DataParallelTrainer(small_cnn(), nn.CrossEntropyLoss(), FakeComm(), lr=0.05, momentum=0.9, autocast_dtype=None,
                             channels_last=False, cuda_graph=False, bucket_bytes=2048)
    data = batches(3, seed=9)
    for x, y in data:
        tr.step(x, y)
    before = [m.clone() for m in tr.state.master_state().values()]
    bufs = [b.clone() for b in tr.model.buffers()]
    launches = tr.comm.launch_count
    x, y = data[0]
    r1 = tr.evaluate(x, y)
    r2 = tr.evaluate(x, y)
    assert r1 == r2 and set(r1) == {"loss", "top1", "top5", "examples"} and r1["examples"] == 4 and 0.0 <= r1["top1"] <= r1["top5"] <= 1.0
    assert tr.model.training and tr.comm.launch_count == launches                      # back in training mode, no collective
    assert all(torch.equal(a, b) for a, b in zip(before, tr.state.master_state().values()))
    assert all(torch.equal(a, b) for a, b in zip(bufs, tr.model.buffers()))             # running statistics untouched
    ref = copy.deepcopy(tr.model).eval()
    with torch.no_grad():
        want = float(nn.CrossEntropyLoss()(ref(x), y))
    assert r1["loss"] == pytest.approx(want, rel=1e-6)


@pytest.mark.skipif(len(os.sched_getaffinity(0)) < 4, reason="needs 4 schedulable CPUs")
def test_affinity_next_to_the_gpu_is_opt_in_and_reads_sysfs(tmp_path, monkeypatch):
    """runtime/affinity.py: B200MPI_BIND_TO=numa binds a torchrun-started rank to the CPUs of its GPU's NUMA node (same placement
    csrc/spawner/mpirun.cc gives with --bind-to numa); the default leaves the process alone."""
    from mpi_operator_b200.runtime import affinity
    before = os.sched_getaffinity(0)
    cpus = sorted(before)[:4]
    for n, part in enumerate((cpus[:2], cpus[2:])):
        d = tmp_path / f"sys/devices/system/node/node{n}"
        d.mkdir(parents=True)
        (d / "cpulist").write_text(affinity.format_cpulist(part) + "\n")
    for k, node in enumerate((0, 1, -1)):
        bus = f"0000:{0x1b + 0x20 * k:02X}:00.0"
        (tmp_path / "proc/driver/nvidia/gpus" / bus).mkdir(parents=True)
        d = tmp_path / "sys/bus/pci/devices" / bus.lower()
        d.mkdir(parents=True)
        (d / "numa_node").write_text(f"{node}\n")
    monkeypatch.setenv("B200MPI_SYSFS_ROOT", str(tmp_path))
    monkeypatch.delenv("B200MPI_BOUND_CPUS", raising=False)
    monkeypatch.delenv("B200MPI_BIND_TO", raising=False)
    assert affinity.parse_cpulist("0-2,5,7-8") == [0, 1, 2, 5, 7, 8] and affinity.format_cpulist([8, 0, 1, 2, 5, 7]) == "0-2,5,7-8"
    try:
        assert affinity.maybe_bind(1) is None and os.sched_getaffinity(0) == before            # not asked: nothing happens
        monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "1,0")
        assert affinity.gpu_numa_node(0) == 1 and affinity.gpu_numa_node(1) == 0               # CUDA ordinals follow CUDA_VISIBLE_DEVICES
        monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "")
        assert affinity.gpu_numa_node(2) is None and affinity.bind_near_gpu(2) is None         # the platform does not say: unbound
        assert affinity.gpu_numa_node(7) is None                                               # no such GPU
        monkeypatch.setenv("B200MPI_BIND_TO", "numa")
        info = affinity.maybe_bind(1)
        assert info == {"numa": 1, "cpus": affinity.format_cpulist(cpus[2:])} and os.sched_getaffinity(0) == set(cpus[2:])
        assert os.environ["B200MPI_BOUND_NUMA"] == "1"
        assert affinity.maybe_bind(0) is None and os.sched_getaffinity(0) == set(cpus[2:])     # bound once (by us or by mpirun): kept
    finally:
        os.sched_setaffinity(0, before)
        os.environ.pop("B200MPI_BOUND_CPUS", None)
        os.environ.pop("B200MPI_BOUND_NUMA", None)
