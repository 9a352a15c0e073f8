"""DataParallelTrainer (flat symmetric params/grads + fused allreduce+SGD kernel,
optionally inside a CUDA graph) must match torch.optim.SGD on the same model."""
import copy
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _small_model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(),
                         nn.Conv2d(8, 16, 3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                         nn.Linear(16, 10))


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_matches_torch_sgd(graph):
    from mpi_operator_b200.parallel.data_parallel import DataParallelTrainer
    from mpi_operator_b200.runtime.comm import Communicator
    comm = Communicator.create(0, 1, 0, f"t-trainer-{os.getpid()}-{int(graph)}")
    model = _small_model()
    ref = copy.deepcopy(model).cuda().to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    tr = DataParallelTrainer(model, nn.CrossEntropyLoss(), comm, lr=0.05, momentum=0.9, weight_decay=1e-4,
                             autocast_dtype=None, cuda_graph=graph, bucket_bytes=2048)
    assert len(tr.state.buckets) > 1
    torch.manual_seed(1)
    xs = [torch.randn(8, 3, 16, 16) for _ in range(6)]
    ys = [torch.randint(0, 10, (8,)) for _ in range(6)]
    warm = 3 if graph else 0  # graph capture runs 3 eager warm-up steps + the captured step on the first batch
    for i, (x, y) in enumerate(zip(xs, ys)):
        loss = tr.step(x.pin_memory(), y.pin_memory())
        # reference: the first graph-mode step = 3 eager warm-ups + capture (not executed) + 1 replay
        reps = (warm + 1) if (graph and i == 0) else 1
        for _ in range(reps):
            opt.zero_grad()
            l = nn.functional.cross_entropy(ref(x.cuda().contiguous(memory_format=torch.channels_last)), y.cuda())
            l.backward()
            opt.step()
    torch.cuda.synchronize()
    comm.check_error()
    for p, q in zip(model.parameters(), ref.parameters()):
        torch.testing.assert_close(p.data, q.data, rtol=2e-4, atol=2e-5)
    assert tr.launches_per_step == len(tr.state.buckets)
    comm.destroy()


@pytest.mark.parametrize("graph", [False, True])
def test_bf16_params_trainer_tracks_the_fp32_master_trainer(graph):
    """Same model, same batches: the bf16-shadow trainer (bf16 leaves, fp32 masters in the window, shadow written by
    k_allreduce_sgd's lowp output) must follow the default autocast trainer step for step."""
    from mpi_operator_b200.parallel.data_parallel import DataParallelTrainer
    from mpi_operator_b200.runtime.comm import Communicator
    comm = Communicator.create(0, 1, 0, f"t-bf16p-{os.getpid()}-{int(graph)}")
    base = _small_model()
    kw = dict(lr=0.05, momentum=0.9, weight_decay=1e-4, autocast_dtype=torch.bfloat16, cuda_graph=graph, bucket_bytes=2048)
    tr_a = DataParallelTrainer(copy.deepcopy(base), nn.CrossEntropyLoss(), comm, bf16_params=False, **kw)
    tr_b = DataParallelTrainer(copy.deepcopy(base), nn.CrossEntropyLoss(), comm, bf16_params=True, **kw)
    assert all(p.dtype == torch.bfloat16 for p in tr_b.model.parameters() if p.dim() >= 2)
    torch.manual_seed(1)
    for _ in range(5):
        x, y = torch.randn(8, 3, 16, 16).pin_memory(), torch.randint(0, 10, (8,)).pin_memory()
        la, lb = float(tr_a.step(x, y)), float(tr_b.step(x, y))
        assert lb == pytest.approx(la, rel=2e-2, abs=2e-2)
    torch.cuda.synchronize()
    comm.check_error()
    for m, p in zip(tr_b.state.master_state().values(), tr_a.model.parameters()):
        torch.testing.assert_close(m, p.data, rtol=2e-2, atol=2e-3)
    for m, p in zip(tr_b.state.master_state().values(), tr_b.model.parameters()):
        if p.dim() >= 2:
            assert torch.equal(p.data, m.to(torch.bfloat16))
    comm.destroy()


def test_collective_counters_cover_eager_launches_and_graph_replays():
    """Communicator.stats(): host launches counted natively, CUDA-graph replays added by the trainer (what the node
    agent exports as b200mpi_collective_*_total)."""
    from mpi_operator_b200.parallel.data_parallel import DataParallelTrainer
    from mpi_operator_b200.runtime.comm import Communicator
    comm = Communicator.create(0, 1, 0, f"t-stats-{os.getpid()}")
    t = torch.ones(1024, device="cuda")
    comm.allreduce(t)
    ops = {(o["op"], o["algo"]): o for o in comm.stats()["ops"]}
    assert sum(o["calls"] for o in ops.values()) == 1 and sum(o["bytes"] for o in ops.values()) == 4096
    tr = DataParallelTrainer(_small_model(), nn.CrossEntropyLoss(), comm, lr=0.05, autocast_dtype=None, cuda_graph=True, bucket_bytes=2048)
    x, y = torch.randn(8, 3, 16, 16).pin_memory(), torch.randint(0, 10, (8,)).pin_memory()
    for _ in range(4):
        tr.step(x, y)
    torch.cuda.synchronize()
    nb = len(tr.state.buckets)
    sgd = [o for o in comm.stats()["ops"] if o["op"].startswith("allreduce_sgd")]
    # 3 eager warm-ups + 1 capture are host launches; 4 replays come from the trainer's graph accounting
    assert sum(o["calls"] for o in sgd) == nb * (3 + 1 + 4)
    comm.destroy()


@pytest.mark.parametrize("graph", [False, True])
def test_async_h2d_pipeline_trains_on_the_same_batches(graph):
    """Copy-stream H2D into double-buffered staging + D2D into the graph's static input must feed exactly the batches the
    synchronous path feeds: identical parameters after the same steps, with and without the prefetch API."""
    from mpi_operator_b200.parallel.data_parallel import DataParallelTrainer
    from mpi_operator_b200.runtime.comm import Communicator
    comm = Communicator.create(0, 1, 0, f"t-h2d-{os.getpid()}-{int(graph)}")
    base = _small_model()
    kw = dict(lr=0.05, momentum=0.9, autocast_dtype=None, cuda_graph=graph, bucket_bytes=2048)
    ref = DataParallelTrainer(copy.deepcopy(base), nn.CrossEntropyLoss(), comm, async_h2d=False, **kw)
    a = DataParallelTrainer(copy.deepcopy(base), nn.CrossEntropyLoss(), comm, async_h2d=True, **kw)
    b = DataParallelTrainer(copy.deepcopy(base), nn.CrossEntropyLoss(), comm, async_h2d=True, **kw)
    torch.manual_seed(3)
    data = [(torch.randn(8, 3, 16, 16).pin_memory(), torch.randint(0, 10, (8,)).pin_memory()) for _ in range(7)]
    b.prefetch(*data[0])
    for i, (x, y) in enumerate(data):
        lr_ = ref.step(x, y)
        la = a.step(x, y)                      # no host sync in between: copies of later batches overlap earlier steps
        lb = b.step()
        if i + 1 < len(data):
            b.prefetch(*data[i + 1])
    torch.cuda.synchronize()
    comm.check_error()
    assert float(lr_) == float(la) == float(lb)
    for p, q, r in zip(ref.model.parameters(), a.model.parameters(), b.model.parameters()):
        assert torch.equal(p.data, q.data) and torch.equal(p.data, r.data)
    comm.destroy()
