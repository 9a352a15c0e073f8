"""The native hvdcore engine (csrc/hvd_core; SURVEY.md §2.2 "Horovod core": negotiation thread, fusion buffer, response
cache, timeline, stall inspector, join) — C++ semantics test on 1/3/4 ranks, then the horovod.torch async API on top of it
under the native mpirun. All CPU: the host executor reduces through the shared-memory mailboxes."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MPIRUN = os.path.join(REPO, "mpi_operator_b200/bin/mpirun")
WORKER = os.path.join(REPO, "tests/hvd_engine_worker.py")

pytestmark = pytest.mark.skipif(not os.path.exists(MPIRUN), reason="native launcher not built (run make)")


def _run(np_, *args, env=None, timeout=180):
    e = dict(os.environ, B200MPI_HVD_DEVICE="cpu")
    e.update(env or {})
    return subprocess.run([MPIRUN, "-np", str(np_), sys.executable, WORKER, *args], capture_output=True, text=True, timeout=timeout,
                          env=e, cwd="/tmp")


def test_native_engine_semantics_on_4_3_and_1_ranks():
    """csrc/tests/hvd_core_test.cc: out-of-order named submissions, fusion, response cache, every dtype / reduction,
    messages larger than a mailbox, allgatherv / broadcast / alltoallv / exchange, mismatch and duplicate errors, join."""
    r = subprocess.run(["make", "-C", REPO, "test_hvd_core"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for n in (4, 3, 1):
        assert f"all checks passed on {n} ranks" in r.stdout


def test_requests_that_arrive_in_different_cycles():
    """Same program with a rank-dependent pause before every submission and a 5 ms cycle: entries live across cycles."""
    exe = os.path.join(REPO, "build/san/hvd_core_test")
    if not os.path.exists(exe):
        pytest.skip("run after test_native_engine_semantics_on_4_3_and_1_ranks")
    env = dict(os.environ, HVD_TEST_JITTER_US="3000", HOROVOD_CYCLE_TIME="5")
    r = subprocess.run([MPIRUN, "-np", "3", exe], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
    assert r.returncode == 0 and "all checks passed on 3 ranks" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("np_", [2, 4])
def test_horovod_async_api_on_the_engine(np_, tmp_path):
    """tests/hvd_engine_worker.py: allreduce_async_/synchronize/poll with per-rank submission order, grouped allreduce,
    ragged allgather, error propagation, dropped handles, join() with uneven steps, the engine-backed DistributedOptimizer
    with fp16 compression, and a timeline whose every B has its E."""
    tl = tmp_path / "timeline.json"
    r = _run(np_, env={"HVD_TEST_TIMELINE": str(tl)})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("hvd engine ok") == np_
    ev = json.load(open(tl))
    assert any(e.get("name") == "NEGOTIATE_ALLREDUCE" for e in ev)


def test_slow_cycles_batch_more_tensors_per_negotiation():
    r = _run(2, env={"HOROVOD_CYCLE_TIME": "5"})
    assert r.returncode == 0 and r.stdout.count("hvd engine ok") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_stall_inspector_names_the_missing_rank_and_tensor():
    r = _run(2, "stall", env={"HOROVOD_STALL_CHECK_TIME_SECONDS": "1"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank 1: late.tensor" in r.stderr and "have not submitted" in r.stderr


def test_stall_shutdown_fails_the_waiters_on_every_rank():
    r = _run(3, "stall_shutdown", env={"HOROVOD_STALL_CHECK_TIME_SECONDS": "1", "HOROVOD_STALL_SHUTDOWN_TIME_SECONDS": "2"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("stall_shutdown ok") == 3


def test_a_dead_peer_fails_the_outstanding_handles_instead_of_hanging():
    r = _run(3, "peer_death", timeout=120)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("peer death detected") == 2


def test_autotune_walks_the_grid_and_every_rank_follows_rank_0(tmp_path):
    log = tmp_path / "autotune.csv"
    r = _run(3, "autotune", env={"HOROVOD_AUTOTUNE": "1", "HOROVOD_AUTOTUNE_LOG": str(log), "HOROVOD_AUTOTUNE_WARMUP_SAMPLES": "1",
                                 "HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE": "3"}, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("autotune ok") == 3 and "# best: cycle" in log.read_text()


def test_engine_can_be_disabled():
    """B200MPI_HVD_ENGINE=0: the front-end falls back to the direct (call-order) path over the libmpi shim."""
    code = ("import sys; sys.path.insert(0, %r); import torch, horovod.torch as hvd; hvd.init(); "
            "assert hvd.engine_stats() == {}; "
            "assert torch.equal(hvd.allreduce(torch.ones(3), op=hvd.Sum), torch.full((3,), float(hvd.size()))); "
            "assert hvd.synchronize(hvd.allreduce_async_(torch.ones(2), op=hvd.Sum))[0] == hvd.size(); "
            "hvd.shutdown(); print('direct ok')" % REPO)
    e = dict(os.environ, B200MPI_HVD_DEVICE="cpu", B200MPI_HVD_ENGINE="0")
    r = subprocess.run([MPIRUN, "-np", "2", sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=e, cwd="/tmp")
    assert r.returncode == 0 and r.stdout.count("direct ok") == 2, r.stdout[-2000:] + r.stderr[-3000:]
