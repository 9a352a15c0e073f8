"""mpijobctl against a live in-process daemon (the kubectl workflow of the reference README.md:63-170)."""
import io
import json
import socket
import time
from contextlib import redirect_stderr, redirect_stdout

import pytest

from mpi_operator_b200.cmd import mpijobctl
from mpi_operator_b200.cmd.options import ServerOption
from mpi_operator_b200.cmd.server import Operator

JOB = """
apiVersion: kubeflow.org/v2beta1
kind: MPIJob
metadata:
  name: cli
spec:
  runPolicy: {cleanPodPolicy: Running}
  mpiReplicaSpecs:
    Launcher:
      replicas: 1
      template: {spec: {containers: [{name: l, command: [sh, -c, "echo launcher-says-hi; sleep 0.3"]}]}}
    Worker:
      replicas: 2
      template: {spec: {containers: [{name: w, command: [/usr/sbin/sshd, -De]}]}}
"""


@pytest.fixture
def server(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    op = Operator(ServerOption(fake_gpus=4, leader_elect=False, state_dir=str(tmp_path)))
    op.serve(f"127.0.0.1:{port}")
    op.start()
    yield f"127.0.0.1:{port}"
    op.stop()


def ctl(server, *args):
    out, err = io.StringIO(), io.StringIO()
    with redirect_stdout(out), redirect_stderr(err):
        rc = mpijobctl.main(["--server", server, *args])
    return rc, out.getvalue(), err.getvalue()


def test_apply_get_describe_logs_wait_delete(server, tmp_path):
    f = tmp_path / "job.yaml"
    f.write_text(JOB)
    rc, out, _ = ctl(server, "apply", "-f", str(f))
    assert rc == 0 and "mpijob.kubeflow.org/cli created" in out
    assert "unchanged" in ctl(server, "apply", "-f", str(f))[1]
    rc, out, _ = ctl(server, "wait", "cli", "--for", "Succeeded", "--timeout", "20")
    assert rc == 0 and "condition met: Succeeded" in out
    rc, out, _ = ctl(server, "get", "mpijobs")
    assert rc == 0 and "cli" in out and "Succeeded" in out
    rc, out, _ = ctl(server, "get", "mpijob", "cli", "-o", "json")
    assert json.loads(out)["status"]["replicaStatuses"]["Launcher"]["succeeded"] == 1
    rc, out, _ = ctl(server, "describe", "cli")
    assert "MPIJobSucceeded" in out and "Conditions:" in out and "Launcher: replicas=1" in out
    assert "launcher-says-hi" in ctl(server, "logs", "cli")[1]
    rc, out, _ = ctl(server, "get", "pods")
    assert "cli-launcher-" in out
    rc, out, _ = ctl(server, "get", "events")
    assert "MPIJobCreated" in out
    assert ctl(server, "delete", "mpijob", "cli")[0] == 0
    rc, _, err = ctl(server, "get", "mpijob", "cli")
    assert rc == 1 and "NotFound" in err
    assert ctl(server, "get", "nonsense")[0] == 1


def test_scale_suspend_resume(server, tmp_path):
    f = tmp_path / "job.yaml"
    f.write_text(JOB.replace("sleep 0.3", "sleep 30").replace("name: cli", "name: longjob"))
    assert ctl(server, "apply", "-f", str(f))[0] == 0
    assert ctl(server, "wait", "longjob", "--for", "Running", "--timeout", "20")[0] == 0
    assert ctl(server, "scale", "longjob", "--replicas", "4")[0] == 0
    deadline = time.time() + 10
    while time.time() < deadline and ctl(server, "get", "pods")[1].count("longjob-worker-") < 4:
        time.sleep(0.1)
    assert ctl(server, "get", "pods")[1].count("longjob-worker-") == 4
    assert ctl(server, "suspend", "longjob")[0] == 0
    assert ctl(server, "wait", "longjob", "--for", "Suspended", "--timeout", "20")[0] == 0
    deadline = time.time() + 10
    while time.time() < deadline and "longjob-" in ctl(server, "get", "pods")[1]:
        time.sleep(0.1)
    assert "longjob-" not in ctl(server, "get", "pods")[1]  # suspended: no pods at all
    assert ctl(server, "resume", "longjob")[0] == 0
    assert ctl(server, "wait", "longjob", "--for", "Running", "--timeout", "20")[0] == 0
    rc, out, _ = ctl(server, "topology")
    assert rc == 0 and json.loads(out)["source"] == "fake"
    assert ctl(server, "wait", "longjob", "--for", "Succeeded", "--timeout", "0.5")[0] == 1  # times out -> rc 1


def test_run_standalone_and_version(tmp_path, capsys):
    f = tmp_path / "job.yaml"
    f.write_text(JOB)
    rc = mpijobctl.main(["run", "-f", str(f), "--fake-gpus", "2", "--timeout", "30"])
    out = capsys.readouterr().out
    assert rc == 0 and "launcher-says-hi" in out and "Succeeded after" in out
    assert mpijobctl.main(["version"]) == 0
