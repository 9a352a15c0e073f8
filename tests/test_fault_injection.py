"""Fault-injection hooks (SURVEY.md §5.3 [NEW]): a rank kills / exits / hangs itself at a chosen step and the
launcher, the Job back-off and the MPIJob conditions react the way the reference's failure tests expect
(test/integration/mpi_job_controller_test.go:538-655 plays the same scenario by writing failed pod statuses)."""
import os
import signal
import subprocess
import sys
import time

import pytest

from helpers import conds, new_mpijob
from mpi_operator_b200.api import constants as C
from mpi_operator_b200.cmd.options import ServerOption
from mpi_operator_b200.cmd.server import Operator
from mpi_operator_b200.utils import fault

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MPIRUN = os.path.join(REPO, "mpi_operator_b200/bin/mpirun")
WORKER = os.path.join(REPO, "tests/fault_worker.py")
needs_native = pytest.mark.skipif(not os.path.exists(MPIRUN), reason="native launcher not built (run make)")


def test_spec_grammar():
    s = fault.FaultSpec.parse("kill_rank:3@step:50")
    assert (s.action, s.rank, s.trigger, s.at, s.once) == ("kill", 3, "step", 50, False)
    s = fault.FaultSpec.parse(" exit_rank:1@step:10:code=7 ; once ")
    assert (s.action, s.code, s.once) == ("exit", 7, True)
    assert fault.FaultSpec.parse("hang_rank:2@time:2.5").trigger == "time"
    for bad in ("", "x", "kill_rank:a@step:1", "boom_rank:1@step:1", "kill_rank:1@when:1", "kill_rank:1@step:1;twice",
                "kill_rank:1@step:-1", "exit_rank:1@step:1:rc=3"):
        with pytest.raises(ValueError):
            fault.FaultSpec.parse(bad)


def test_injector_targets_one_rank_and_once_marker(tmp_path):
    spec = fault.FaultSpec.parse("exit_rank:1@step:3;once")
    env = {"B200MPI_FAULT_DIR": str(tmp_path), "B200MPI_MPIJOB_NAME": "j"}
    other = fault.FaultInjector(spec, rank=0, env=env)
    for _ in range(10):
        other.on_step()          # not this rank: never fires
    assert other.spec is None and not other.armed
    fired = []
    inj = fault.FaultInjector(spec, rank=1, env=env)
    inj.fire = lambda: fired.append(inj.steps)
    inj.on_step(); inj.on_step()
    assert not fired and inj.armed
    inj.on_step()
    assert fired == [3]
    # the real fire() leaves a marker; a second attempt of the same job is then disarmed
    real = fault.FaultInjector(spec, rank=1, env=env)
    open(fault._marker_path(spec, env), "w").close()
    assert not real.armed
    real.on_step(5)              # would be due, but the marker says it already happened


def _run_worker(env_extra, rank=0, steps=5):
    env = dict(os.environ, B200MPI_RANK=str(rank), B200MPI_WORLD_SIZE="2", **env_extra)
    return subprocess.run([sys.executable, WORKER, str(steps)], env=env, capture_output=True, text=True, timeout=60)


def test_rank_process_dies_at_the_requested_step(tmp_path):
    r = _run_worker({"B200MPI_FAULT": "kill_rank:0@step:3"})
    assert r.returncode == -signal.SIGKILL
    assert "0/2 step 2" in r.stdout and "0/2 step 3" not in r.stdout and "injecting kill at step 3" in r.stdout
    r = _run_worker({"B200MPI_FAULT": "exit_rank:0@step:2:code=7"})
    assert r.returncode == 7
    r = _run_worker({"B200MPI_FAULT": "kill_rank:1@step:3"})   # another rank's fault
    assert r.returncode == 0 and "rank 0 done" in r.stdout
    env = {"B200MPI_FAULT": "exit_rank:0@step:2:code=9;once", "B200MPI_FAULT_DIR": str(tmp_path)}
    assert _run_worker(env).returncode == 9
    assert _run_worker(env).returncode == 0                    # second attempt runs clean


@needs_native
def test_mpirun_propagates_an_injected_rank_failure():
    env = dict(os.environ, B200MPI_FAULT="exit_rank:1@step:2:code=5", STEP_SLEEP="0.05")
    r = subprocess.run([MPIRUN, "-n", "3", sys.executable, WORKER, "50"], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 5
    assert "The first process to do so was" in r.stderr and "[[b200mpi],1]" in r.stderr
    assert "rank 0 done" not in r.stdout  # survivors were torn down, not left to finish 50 steps


@needs_native
def test_mpirun_time_triggered_kill_from_outside_the_rank(tmp_path):
    env = dict(os.environ, B200MPI_FAULT="kill_rank:1@time:0.3;once", B200MPI_FAULT_DIR=str(tmp_path), B200MPI_MPIJOB_NAME="tj")
    t0 = time.time()
    r = subprocess.run([MPIRUN, "-n", "2", "sleep", "30"], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 128 + signal.SIGKILL and time.time() - t0 < 15
    assert "fault injection: SIGKILL rank 1" in r.stderr
    assert os.path.exists(tmp_path / "tj.fault-mpirun.fired")
    r = subprocess.run([MPIRUN, "-n", "2", "sleep", "0.5"], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0  # once: the marker disarms the second attempt


def _wait(fn, timeout=40.0, what="condition"):
    t0 = time.time()
    while time.time() - t0 < timeout:
        v = fn()
        if v:
            return v
        time.sleep(0.05)
    raise AssertionError(f"timed out waiting for {what}")


@needs_native
def test_mpijob_recovers_after_an_injected_rank_failure(tmp_path):
    """Launcher restartPolicy OnFailure (the reference default, default.go:31-33): rank 1 is killed at step 3 of the
    first attempt, mpirun fails, the launcher container restarts in place, the second attempt completes -> Succeeded."""
    o = Operator(ServerOption(fake_gpus=0, leader_elect=False, state_dir=str(tmp_path)))
    o.start()
    try:
        job = new_mpijob("faulty", workers=2, launcher_cmd=("mpirun",),
                         launcher_args=("-n", "2", sys.executable, WORKER, "6"), worker_cmd=("/usr/sbin/sshd", "-De"))
        job.spec.replica("Launcher").restart_policy = "OnFailure"
        job.spec.replica("Launcher").template["spec"]["containers"][0]["env"] = [
            {"name": "B200MPI_FAULT", "value": "kill_rank:1@step:3;once"}]
        cs = o.clientset.kubeflow_v2beta1().mpijobs("default")
        cs.create(job)
        done = _wait(lambda: conds(cs.get("faulty")).get("Succeeded") == "True" and cs.get("faulty"), what="Succeeded after restart")
        launcher = [p for p in o.store.list("pods", "default") if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "launcher"][0]
        assert launcher["status"]["containerStatuses"][0]["restartCount"] == 1
        log = o.agent.logs("default", launcher["metadata"]["name"])
        assert "injecting kill at step 3" in log and log.count("rank 0 done") == 1
        assert done.status.replica_statuses["Launcher"].succeeded == 1
    finally:
        o.stop()


@needs_native
def test_mpijob_fails_when_the_fault_repeats_past_backoff_limit(tmp_path):
    o = Operator(ServerOption(fake_gpus=0, leader_elect=False, state_dir=str(tmp_path)))
    o.start()
    try:
        job = new_mpijob("doomed", workers=2, launcher_cmd=("mpirun",), launcher_args=("-n", "2", sys.executable, WORKER, "6"),
                         worker_cmd=("/usr/sbin/sshd", "-De"), backoff_limit=0)
        job.spec.replica("Launcher").restart_policy = "Never"
        job.spec.replica("Launcher").template["spec"]["containers"][0]["env"] = [
            {"name": "B200MPI_FAULT", "value": "exit_rank:0@step:2:code=3"}]
        cs = o.clientset.kubeflow_v2beta1().mpijobs("default")
        cs.create(job)
        failed = _wait(lambda: conds(cs.get("doomed")).get("Failed") == "True" and cs.get("doomed"), what="Failed")
        assert [c.reason for c in failed.status.conditions if c.type == "Failed"] == ["BackoffLimitExceeded/Error"]  # launcher Job reason + "/" + last failed pod reason (controller.go:1176-1207)
    finally:
        o.stop()
