"""mpijobctl against a live in-process daemon (the kubectl workflow of the reference README.md:63-170)."""
import io
import os
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
    assert ctl(server, "logs", "cli", "--tail", "1")[1].count("\n") == 1 and ctl(server, "logs", "cli", "--tail", "0")[1] == ""
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


def test_horovodrun_maps_its_flags_onto_mpirun_and_the_engine(tmp_path, capsys):
    """`horovodrun` is Horovod's launcher CLI; here it is a front for the native mpirun and the HOROVOD_* knobs of hvdcore."""
    import subprocess
    import sys
    from mpi_operator_b200.cmd import horovodrun
    a = horovodrun.build_parser().parse_args(["-np", "4", "-H", "localhost:4", "--fusion-threshold-mb", "32", "--cycle-time-ms", "2.5",
                                              "--cache-capacity", "0", "--timeline-filename", "/tmp/t.json", "--no-stall-check",
                                              "--start-timeout", "30", "-x", "FOO=bar", "--mpi-args=-tag-output", "python", "train.py", "--lr", "0.1"])
    argv = horovodrun.mpirun_argv(a)
    assert argv[0].endswith("bin/mpirun") and argv[1:5] == ["-np", "4", "-H", "localhost:4"]
    assert argv[-4:] == ["python", "train.py", "--lr", "0.1"] and "-tag-output" in argv
    exported = {argv[i + 1] for i, v in enumerate(argv) if v == "-x"}
    assert {"HOROVOD_FUSION_THRESHOLD=33554432", "HOROVOD_CYCLE_TIME=2.5", "HOROVOD_CACHE_CAPACITY=0", "HOROVOD_TIMELINE=/tmp/t.json",
            "HOROVOD_STALL_CHECK_DISABLE=1", "B200MPI_INIT_TIMEOUT_MS=30000", "FOO=bar"} <= exported
    assert horovodrun.main(["--check-build"]) == 0 and "[X] PyTorch" in capsys.readouterr().out
    assert horovodrun.main([]) == 2
    if not horovodrun.MPIRUN.exists():
        pytest.skip("native launcher not built (run make)")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, torch, horovod.torch as hvd; hvd.init(); "
            "assert hvd.engine_stats()['cycle_time_ms'] == 3.0 and hvd.engine_stats()['cache_capacity'] == 7; "
            "print('sum', float(hvd.allreduce(torch.ones(1), op=hvd.Sum))); hvd.shutdown()")
    r = subprocess.run([sys.executable, "-m", "mpi_operator_b200.cmd.horovodrun", "-np", "2", "--cycle-time-ms", "3", "--cache-capacity", "7",
                        sys.executable, "-c", code], capture_output=True, text=True, timeout=120,
                       env=dict(os.environ, PYTHONPATH=repo, B200MPI_HVD_DEVICE="cpu"), cwd=repo)
    assert r.returncode == 0 and r.stdout.count("sum 2.0") == 2, r.stdout + r.stderr


def test_logs_follow_and_get_watch(server, tmp_path):
    """kubectl habits from the reference's README (`kubectl logs -f`, `kubectl get -w`): follow the launcher log until the job
    finishes; watch prints one line per state change."""
    f = tmp_path / "job.yaml"
    f.write_text(JOB.replace("name: cli", "name: follow").replace("echo launcher-says-hi; sleep 0.3",
                                                                     "for i in 1 2 3; do echo tick-$i; sleep 0.4; done"))
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    watch = subprocess.Popen([sys.executable, "-m", "mpi_operator_b200.cmd.mpijobctl", "--server", server, "get", "mpijobs", "-w",
                              "--watch-timeout", "6"], stdout=subprocess.PIPE, text=True, cwd=repo)
    assert ctl(server, "apply", "-f", str(f))[0] == 0
    rc, out, err = ctl(server, "logs", "follow", "-f", "--timeout", "30")
    assert rc == 0 and [l for l in out.splitlines() if l.startswith("tick-")] == ["tick-1", "tick-2", "tick-3"], (out, err)
    # the server-side stream (pods/<name>/log?follow=true) delivers the same text and ends when the pod has finished
    from mpi_operator_b200.sdk import MPIJobClient
    cli = MPIJobClient(server)
    pod = [p["metadata"]["name"] for p in cli.list_resource("pods", "default") if "follow-launcher" in p["metadata"]["name"]][0]
    assert "".join(cli.follow_pod_log(pod, timeout=10)).count("tick-") == 3
    from kubernetes import client as k8s, config as k8s_config
    k8s_config.load_kube_config(host=server)
    core = k8s.CoreV1Api()
    assert core.read_namespaced_pod_log(pod, "default", tail_lines=1) == "tick-3\n"
    assert "".join(core.read_namespaced_pod_log(pod, "default", follow=True, _preload_content=False)).count("tick-") == 3
    assert core.read_namespaced_pod(pod, "default").status.phase == "Succeeded"
    wout, _ = watch.communicate(timeout=30)
    assert watch.returncode == 0 and "NAME" in wout and "Succeeded" in wout and wout.count("follow") >= 2, wout   # several change lines


def test_patch_label_annotate(server, tmp_path):
    """`kubectl patch --type merge -p ...`, `kubectl label`, `kubectl annotate` (key=value, key- removes, --overwrite)."""
    f = tmp_path / "job.yaml"
    f.write_text(JOB.replace("sleep 0.3", "sleep 30"))
    assert ctl(server, "apply", "-f", str(f))[0] == 0
    rc, out, err = ctl(server, "patch", "mpijob", "cli", "--type", "merge", "-p", '{"spec": {"runPolicy": {"suspend": true}}}')
    assert rc == 0 and "patched" in out, err
    rc, out, _ = ctl(server, "get", "mpijob", "cli", "-o", "json")
    assert json.loads(out)["spec"]["runPolicy"]["suspend"] is True
    pf = tmp_path / "p.yaml"
    pf.write_text("spec:\n  runPolicy:\n    suspend: false\n")
    assert ctl(server, "patch", "mpijob", "cli", "--patch-file", str(pf))[0] == 0
    assert json.loads(ctl(server, "get", "mpijob", "cli", "-o", "json")[1])["spec"]["runPolicy"]["suspend"] is False
    assert ctl(server, "patch", "mpijob", "cli", "--type", "json", "-p", "[]")[0] == 2
    assert ctl(server, "patch", "mpijob", "cli", "-p", "[1, 2]")[0] == 2
    assert ctl(server, "label", "mpijob", "cli", "team=vision", "tier=batch")[0] == 0
    assert ctl(server, "annotate", "mpijob", "cli", "note=first")[0] == 0
    meta = json.loads(ctl(server, "get", "mpijob", "cli", "-o", "json")[1])["metadata"]
    assert meta["labels"]["team"] == "vision" and meta["labels"]["tier"] == "batch" and meta["annotations"]["note"] == "first"
    rc, _, err = ctl(server, "label", "mpijob", "cli", "team=speech")
    assert rc == 1 and "--overwrite" in err
    assert ctl(server, "label", "mpijob", "cli", "team=speech", "tier-", "--overwrite")[0] == 0
    meta = json.loads(ctl(server, "get", "mpijob", "cli", "-o", "json")[1])["metadata"]
    assert meta["labels"]["team"] == "speech" and "tier" not in meta["labels"]
    assert ctl(server, "delete", "mpijob", "cli")[0] == 0


def test_cordon_and_uncordon_gpus(server):
    rc, out, err = ctl(server, "cordon", "1", "3", "--reason", "swap the baseboard")
    assert rc == 0 and "gpu/1 cordoned" in out and "free GPUs: 2" in out, err
    topo = json.loads(ctl(server, "topology")[1])
    assert topo["cordoned"] == {"1": "swap the baseboard", "3": "swap the baseboard"} and topo["free_gpus"] == 2
    rc, out, _ = ctl(server, "uncordon", "1")
    assert rc == 0 and "free GPUs: 3" in out
    rc, _, err = ctl(server, "cordon", "9")
    assert rc == 1 and "no GPU 9" in err


def test_get_nodes_shows_the_box_with_capacity_and_cordons(server):
    """`kubectl get nodes` / `CoreV1Api.list_node()`: the box as a v1.Node - GPU capacity, allocatable minus cordoned GPUs, taints."""
    assert ctl(server, "cordon", "2", "--reason", "fan")[0] == 0
    rc, out, err = ctl(server, "get", "nodes")
    assert rc == 0 and "ALLOCATABLE" in out and " 4 " in out and " 3 " in out and out.strip().endswith("2"), out + err
    node = json.loads(ctl(server, "get", "nodes", "-o", "json")[1])["items"][0]
    assert node["status"]["capacity"]["nvidia.com/gpu"] == "4" and node["status"]["allocatable"]["nvidia.com/gpu"] == "3"
    assert node["spec"]["taints"] == [{"key": "b200mpi.kubeflow.org/gpu-2", "value": "fan", "effect": "NoSchedule"}]
    assert node["metadata"]["labels"]["nvidia.com/gpu.count"] == "4" and node["spec"]["unschedulable"] is False
    rc, out, _ = ctl(server, "get", "node", node["metadata"]["name"], "-o", "yaml")
    assert rc == 0 and "kind: Node" in out
    os.environ["MPIJOB_SERVER"] = server
    try:
        import kubernetes
        kubernetes.config.load_kube_config()
        nodes = kubernetes.client.CoreV1Api().list_node().items
        assert len(nodes) == 1 and nodes[0].status.allocatable["nvidia.com/gpu"] == "3"
    finally:
        os.environ.pop("MPIJOB_SERVER", None)
