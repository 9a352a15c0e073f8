"""Integration tier: the real operator (controller + node agent) driving real
CPU processes on localhost — analogue of the reference's envtest + e2e tiers
(test/integration/mpi_job_controller_test.go:50-977, test/e2e/mpi_job_test.go:92-584)."""
import os
import time

import pytest

from helpers import conds, new_mpijob
from mpi_operator_b200.api import constants as C
from mpi_operator_b200.api import yaml_io
from mpi_operator_b200.api.types import SchedulingPolicy
from mpi_operator_b200.cmd.options import ServerOption
from mpi_operator_b200.cmd.server import Operator
from mpi_operator_b200.controller import metrics

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_native = pytest.mark.skipif(not os.path.exists(os.path.join(REPO, "mpi_operator_b200/bin/mpirun")),
                                  reason="native launcher not built (run make)")


@pytest.fixture
def op(tmp_path):
    o = Operator(ServerOption(fake_gpus=8, leader_elect=False, state_dir=str(tmp_path)))
    o.start()
    yield o
    o.stop()


@pytest.fixture
def gang_op(tmp_path):
    o = Operator(ServerOption(fake_gpus=4, leader_elect=False, state_dir=str(tmp_path), gang_scheduling_name="volcano"))
    o.start()
    yield o
    o.stop()


def wait_for(fn, timeout=20.0, what="condition"):
    t0 = time.time()
    while time.time() - t0 < timeout:
        v = fn()
        if v:
            return v
        time.sleep(0.02)
    raise AssertionError(f"timed out waiting for {what}")


def events(op, name):
    return [(e["type"], e["reason"]) for e in sorted(op.store.list("events"), key=lambda e: e["metadata"]["creationTimestamp"] + e["metadata"]["name"])
            if e["involvedObject"]["name"] == name]


def submit(op, job):
    return op.clientset.kubeflow_v2beta1().mpijobs(job.namespace).create(job)


def get(op, job):
    return op.clientset.kubeflow_v2beta1().mpijobs(job.namespace).get(job.name)


@needs_native
def test_pi_job_happy_path_events_and_cleanup(op):
    job = yaml_io.load_file(os.path.join(REPO, "examples/pi/pi.yaml"))[0]
    job.metadata["namespace"] = "default"
    t0 = time.time()
    submit(op, job)
    done = wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True" and get(op, job), what="Succeeded")
    assert time.time() - t0 < 30  # reference e2e budget: 200 s per wait
    assert done.status.completion_time and done.status.start_time
    assert done.status.replica_statuses["Launcher"].succeeded == 1
    ev = [r for _, r in events(op, "pi")]
    assert ev[0] == "MPIJobCreated" and "MPIJobRunning" in ev and ev[-1] == "MPIJobSucceeded"
    pods = op.store.list("pods", "default")
    launcher = [p for p in pods if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "launcher"][0]
    log = op.agent.logs("default", launcher["metadata"]["name"])
    assert "pi is approximately 3.1" in log and "Worker 1/2 on pi-worker-1" in log
    # cleanPodPolicy: Running -> idle (Running) workers are deleted after completion
    wait_for(lambda: not [p for p in op.store.list("pods", "default") if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "worker"], what="worker cleanup")
    # hostfile + discover_hosts.sh were mounted into the launcher sandbox with the reference's modes
    root = os.path.join(op.agent.pod_dir(launcher), "rootfs", "etc", "mpi")
    assert open(os.path.join(root, "hostfile")).read() == "pi-worker-0.pi.default.svc slots=1\npi-worker-1.pi.default.svc slots=1\n"
    assert os.stat(os.path.join(root, "discover_hosts.sh")).st_mode & 0o111
    assert os.path.exists(os.path.join(op.agent.pod_dir(launcher), "rootfs", "home/mpiuser/.ssh", "id_rsa"))


@needs_native
@pytest.mark.parametrize("flavour,hostfile_env,line", [("mpich", "HYDRA_HOST_FILE", "pi-worker-0.pi.default.svc:1"),
                                                        ("intel", "I_MPI_HYDRA_HOST_FILE", "pi-worker-0.pi.default.svc:1")])
def test_pi_job_hydra_flavours_run_through_image_entrypoint(op, flavour, hostfile_env, line):
    """examples/pi/pi-{mpich,intel}.yaml (reference: examples/v2beta1/pi/pi-mpich.yaml, pi-intel.yaml; e2e
    test/e2e/mpi_job_test.go:191-283): args-only launcher -> image ENTRYPOINT (entrypoint.sh) -> Hydra-dialect mpirun."""
    job = yaml_io.load_file(os.path.join(REPO, f"examples/pi/pi-{flavour}.yaml"))[0]
    job.metadata["namespace"] = "default"
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", what="Succeeded")
    launcher = [p for p in op.store.list("pods", "default") if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "launcher"][0]
    env = {e["name"]: e.get("value") for e in launcher["spec"]["containers"][0]["env"]}
    assert env[hostfile_env] == "/etc/mpi/hostfile"
    root = os.path.join(op.agent.pod_dir(launcher), "rootfs", "etc", "mpi")
    assert open(os.path.join(root, "hostfile")).read().splitlines()[0] == line
    log = op.agent.logs("default", launcher["metadata"]["name"])
    assert "pi is approximately 3.1" in log and "Worker 1/2 on pi-worker-1" in log


@needs_native
def test_generic_mpi_program_with_point_to_point_runs_as_an_mpijob(op):
    """examples/mpi-ring/ring.yaml: a token ring + ping-pong (MPI_Send/MPI_Recv) on 4 ranks over the libmpi shim — MPI programs
    other than the reference's pi example (collectives only) run too."""
    job = yaml_io.load_file(os.path.join(REPO, "examples/mpi-ring/ring.yaml"))[0]
    job.metadata["namespace"] = "default"
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", timeout=60, what="Succeeded")
    launcher = [p for p in op.store.list("pods", "default") if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "launcher"][0]
    log = op.agent.logs("default", launcher["metadata"]["name"])
    assert "ring of 4 ranks closed on mpi-ring-worker-0: token = 10 (expected 10)" in log
    assert "pingpong         8 bytes" in log and "pingpong  16777216 bytes" in log


@needs_native
def test_pi_job_with_custom_cluster_domain(tmp_path):
    """e2e 'with custom cluster-domain' (test/e2e/mpi_job_test.go:532-583): --cluster-domain is appended to every
    hostfile FQDN and the launcher still resolves its workers."""
    o = Operator(ServerOption(fake_gpus=0, leader_elect=False, state_dir=str(tmp_path), cluster_domain="cluster.local"))
    o.start()
    try:
        job = yaml_io.load_file(os.path.join(REPO, "examples/pi/pi.yaml"))[0]
        job.metadata["namespace"] = "default"
        submit(o, job)
        wait_for(lambda: conds(get(o, job)).get("Succeeded") == "True", what="Succeeded")
        launcher = [p for p in o.store.list("pods", "default") if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "launcher"][0]
        root = os.path.join(o.agent.pod_dir(launcher), "rootfs", "etc", "mpi")
        assert open(os.path.join(root, "hostfile")).read().splitlines()[0] == "pi-worker-0.pi.default.svc.cluster.local slots=1"
        assert "Worker 1/2 on pi-worker-1" in o.agent.logs("default", launcher["metadata"]["name"])
    finally:
        o.stop()


@needs_native
def test_elastic_horovod_rescales_2_4_2_and_resumes_from_committed_steps(op, tmp_path):
    """SURVEY.md §3.4 / proposals/elastic-horovod.md end to end with real rank processes (CPU backend): `scale` changes
    Worker.replicas, the controller regenerates discover_hosts.sh, the ranks notice at their next commit and leave with
    the rescale code, the launcher restarts mpirun on the new world, state resumes from rank 0's checkpoint.
    (GPU version: test_elastic_gpu.py.)"""
    ckpt = str(tmp_path / "ckpt.pt")
    job = new_mpijob("elastic", workers=2, launcher_cmd=("mpirun",), worker_cmd=("/usr/sbin/sshd", "-De"),
                     launcher_args=("python", os.path.join(REPO, "examples/horovod/elastic_mnist.py"), "--total-steps", "120",
                                    "--commit-every", "4", "--step-sleep", "0.02", "--checkpoint", ckpt))
    job.spec.replica("Launcher").template["spec"]["containers"][0]["env"] = [{"name": "B200MPI_HVD_DEVICE", "value": "cpu"}]
    c = op.clientset.kubeflow_v2beta1().mpijobs("default")
    c.create(job)

    def logs():
        return "".join(op.agent.logs("default", p["metadata"]["name"]) for p in op.store.list("pods", "default")
                       if "launcher" in p["metadata"]["name"])

    def scale(n):
        j = c.get("elastic")
        j.spec.replica("Worker").replicas = n
        c.update(j)

    wait_for(lambda: "world size 2" in logs(), timeout=60, what="first incarnation")
    scale(4)
    wait_for(lambda: "world size 4" in logs(), timeout=90, what="scaled-up incarnation")
    scale(2)
    wait_for(lambda: logs().count("with world size 2") >= 2, timeout=90, what="scaled-down incarnation")
    wait_for(lambda: conds(c.get("elastic")).get("Succeeded") == "True", timeout=180, what="job success")
    text = logs()
    assert "world sizes seen: [2, 4, 2]" in text, text[-3000:]
    restarts = [ln for ln in text.splitlines() if "(re)started at step" in ln]
    assert len(restarts) == 3 and all(int(ln.split("step ")[1].split()[0]) > 0 for ln in restarts[1:]), restarts


@needs_native
def test_elastic_horovod_rescales_in_place_without_restarting_survivors(op, tmp_path):
    """B200MPI_ELASTIC_INPLACE=1: the native mpirun watches discover_hosts.sh itself, spawns only the ADDITIONAL ranks of a
    larger world (publishing the new world once they are ready) and lets the surplus ranks of a smaller one retire; the
    surviving ranks keep their process and re-form the communicator at their next commit (hvd.elastic, generation counter).
    2 -> 4 -> 2 on the CPU backend: mpirun starts exactly once, rank 0's pid never changes, training resumes from the
    committed step without reloading a checkpoint."""
    ckpt = str(tmp_path / "ckpt.pt")
    job = new_mpijob("inplace", workers=2, launcher_cmd=("mpirun",), worker_cmd=("/usr/sbin/sshd", "-De"),
                     launcher_args=("python", os.path.join(REPO, "examples/horovod/elastic_mnist.py"), "--total-steps", "160",
                                    "--commit-every", "4", "--step-sleep", "0.02", "--checkpoint", ckpt))
    job.spec.replica("Launcher").template["spec"]["containers"][0]["env"] = [
        {"name": "B200MPI_HVD_DEVICE", "value": "cpu"}, {"name": "B200MPI_ELASTIC_INPLACE", "value": "1"}]
    c = op.clientset.kubeflow_v2beta1().mpijobs("default")
    c.create(job)

    def logs():
        return "".join(op.agent.logs("default", p["metadata"]["name"]) for p in op.store.list("pods", "default")
                       if "launcher" in p["metadata"]["name"])

    def scale(n):
        j = c.get("inplace")
        j.spec.replica("Worker").replicas = n
        c.update(j)

    wait_for(lambda: "world size 2" in logs(), timeout=60, what="first incarnation")
    scale(4)
    wait_for(lambda: "world size 4" in logs(), timeout=90, what="scaled-up world")
    scale(2)
    wait_for(lambda: logs().count("with world size 2") >= 2, timeout=90, what="scaled-down world")
    wait_for(lambda: conds(c.get("inplace")).get("Succeeded") == "True", timeout=180, what="job success")
    text = logs()
    assert "world sizes seen: [2, 4, 2]" in text, text[-3000:]
    assert text.count("re-formed in place") == 2, text[-3000:]
    assert "elastic: world 2 -> 4, generation 1: spawned 2 ranks" in text and "elastic: world 4 -> 2, generation 2" in text
    restarts = [ln for ln in text.splitlines() if "(re)started at step" in ln]
    assert len(restarts) == 3 and all(int(ln.split("step ")[1].split()[0]) > 0 for ln in restarts[1:]), restarts
    launcher = [p for p in op.store.list("pods", "default") if "inplace-launcher" in p["metadata"]["name"]]
    assert len(launcher) == 1 and not launcher[0].get("status", {}).get("containerStatuses", [{}])[0].get("restartCount", 0)


@needs_native
def test_horovod_mnist_example_runs_as_a_cpu_job(op):
    """F3: examples/horovod/tensorflow-mnist.yaml — `mpirun -np 2 ... python /examples/tensorflow_mnist.py`,
    the image path remapped to the torch script; on a host without CUDA the hvd collectives run over the libmpi shim."""
    job = yaml_io.load_file(os.path.join(REPO, "examples/horovod/tensorflow-mnist.yaml"))[0]
    job.metadata["namespace"] = "default"
    c0 = job.spec.replica("Launcher").template["spec"]["containers"][0]
    c0["args"] = list(c0["args"]) + ["--steps", "20", "--batch-size", "32"]      # keep the CPU test short
    c0["env"] = list(c0.get("env", [])) + [{"name": "B200MPI_HVD_DEVICE", "value": "cpu"}]
    negotiated = metrics.counter_value(metrics.hvd_tensors)
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", timeout=120, what="Succeeded")
    launcher = [p for p in op.store.list("pods", "default") if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "launcher"][0]
    log = op.agent.logs("default", launcher["metadata"]["name"])
    assert "step 0 loss" in log and "final loss (averaged over 2 ranks)" in log
    # the ranks' Horovod-core engines (broadcast_parameters, the final metric allreduce) report to /metrics through the agent
    assert metrics.counter_value(metrics.hvd_tensors) >= negotiated + 2 * 9
    assert "b200mpi_hvd_tensors_total" in metrics.render().decode()


@needs_native
def test_headline_tensorflow_benchmarks_yaml_runs_on_the_host_with_device_cpu(op):
    """The reference's headline job (examples/v2beta1/tensorflow-benchmarks/tensorflow-benchmarks.yaml:17-42), same YAML and
    mpirun line, with the flags a user would edit: a small model and --device=cpu. Covers image lookup, script path, the
    Open MPI flags of that command line, hvd.init under the operator and the sample-output format (README.md:180-212)."""
    job = yaml_io.load_file(os.path.join(REPO, "examples/tensorflow-benchmarks/tensorflow-benchmarks.yaml"))[0]
    job.metadata["namespace"] = "default"
    c0 = job.spec.replica("Launcher").template["spec"]["containers"][0]
    cmd = [t for t in c0["command"] if not t.startswith("--model=") and not t.startswith("--batch_size=")]
    c0["command"] = cmd + ["--model=trivial", "--batch_size=8", "--device=cpu", "--image_size=32", "--num_batches=20", "--num_warmup_batches=2"]
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", timeout=120, what="Succeeded")
    launcher = [p for p in op.store.list("pods", "default") if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "launcher"][0]
    log = op.agent.logs("default", launcher["metadata"]["name"])
    assert "Model:       trivial" in log and "Batch size:  16 global" in log and "Variables:   horovod" in log
    assert "20\timages/sec:" in log and "total images/sec:" in log


@needs_native
def test_mpijobctl_run_keeps_the_launcher_log_of_a_job_that_failed(tmp_path):
    """A launcher that hits the backoff limit is deleted with its Job (batch/v1 semantics); `mpijobctl run` copies the log
    out while the pod exists, so the reason for the failure is still on the terminal."""
    import subprocess
    import sys
    y = tmp_path / "fail.yaml"
    y.write_text("""apiVersion: kubeflow.org/v2beta1
kind: MPIJob
metadata: {name: doomed}
spec:
  runPolicy: {backoffLimit: 0}
  mpiReplicaSpecs:
    Launcher:
      replicas: 1
      restartPolicy: Never
      template:
        spec:
          containers:
          - {name: l, image: mpioperator/mpi-pi, command: [sh, -c, "echo the-reason-it-failed; exit 3"]}
    Worker:
      replicas: 1
      template:
        spec:
          containers:
          - {name: w, image: mpioperator/mpi-pi}
""")
    r = subprocess.run([sys.executable, "-m", "mpi_operator_b200.cmd.mpijobctl", "run", "-f", str(y), "--fake-gpus", "2", "--timeout", "60"],
                       cwd=REPO, capture_output=True, text=True, timeout=120, env=dict(os.environ, B200MPI_STATE_DIR=str(tmp_path / "state")))
    assert r.returncode == 1, r.stdout + r.stderr
    assert "the-reason-it-failed" in r.stdout and "doomed: Failed" in r.stdout


@needs_native
def test_malformed_command_backoff_limit_failed(op):
    job = new_mpijob("bad", workers=1, launcher_cmd=("mpirun",), launcher_args=("-n", "1", "sh", "-c", "echo boom >&2; exit 7"),
                     worker_cmd=("/usr/sbin/sshd", "-De"), backoff_limit=1)
    job.spec.replica("Launcher").restart_policy = "Never"
    before = metrics.counter_value(metrics.mpi_jobs_failed)
    submit(op, job)
    failed = wait_for(lambda: conds(get(op, job)).get("Failed") == "True" and get(op, job), what="Failed")
    c = [c for c in failed.status.conditions if c.type == "Failed"][0]
    assert c.reason == "BackoffLimitExceeded/Error" and "boom" in c.message
    assert failed.status.replica_statuses["Launcher"].failed == 2  # backoffLimit 1 -> two failed pods
    assert metrics.counter_value(metrics.mpi_jobs_failed) == before + 1
    assert failed.status.completion_time is not None


@needs_native
def test_single_launcher_pod_failure_is_not_job_failure(op):
    marker = os.path.join(op.state_dir, "once")
    script = f"if [ ! -e {marker} ]; then touch {marker}; exit 1; fi; exit 0"
    job = new_mpijob("flaky", workers=1, launcher_cmd=("sh", "-c", script), launcher_args=None, worker_cmd=("/usr/sbin/sshd",))
    job.spec.replica("Launcher").restart_policy = "Never"
    submit(op, job)
    ok = wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True" and get(op, job), what="Succeeded after one retry")
    assert ok.status.replica_statuses["Launcher"].failed == 1 and "Failed" not in conds(ok)


@needs_native
def test_on_failure_restarts_in_place(op):
    marker = os.path.join(op.state_dir, "once2")
    script = f"if [ ! -e {marker} ]; then touch {marker}; exit 3; fi; exit 0"
    job = new_mpijob("inplace", workers=None, launcher_cmd=("sh", "-c", script), launcher_args=None)  # default launcher policy OnFailure
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", what="Succeeded")
    pods = [p for p in op.store.list("pods", "default") if p["metadata"]["name"].startswith("inplace-launcher")]
    assert len(pods) == 1 and pods[0]["status"]["containerStatuses"][0]["restartCount"] == 1


def test_created_suspended_has_no_pods_until_resumed(op):
    job = new_mpijob("susp", workers=2, launcher_cmd=("sh", "-c", "sleep 0.2"), launcher_args=None, worker_cmd=("/usr/sbin/sshd",), suspend=True)
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Suspended") == "True", what="Suspended")
    time.sleep(0.3)
    assert op.store.list("pods", "default") == []
    assert get(op, job).status.start_time is None
    j = get(op, job)
    j.spec.run_policy.suspend = False
    op.clientset.kubeflow_v2beta1().mpijobs("default").update(j)
    done = wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True" and get(op, job), what="Succeeded after resume")
    assert conds(done)["Suspended"] == "False" and done.status.start_time is not None
    assert ("Normal", "MPIJobResumed") in events(op, "susp")


def test_suspend_running_job_kills_processes(op):
    job = new_mpijob("longrun", workers=2, launcher_cmd=("sh", "-c", "sleep 60"), launcher_args=None, worker_cmd=("/usr/sbin/sshd",))
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Running") == "True", what="Running")
    j = get(op, job)
    j.spec.run_policy.suspend = True
    op.clientset.kubeflow_v2beta1().mpijobs("default").update(j)
    wait_for(lambda: not op.store.list("pods", "default"), what="all pods deleted on suspend")
    c = conds(get(op, job))
    assert c["Suspended"] == "True" and c["Running"] == "False"
    assert op.agent.alloc.free_gpus == 8


def test_managed_by_multikueue_is_untouched(op):
    job = new_mpijob("ext", workers=1, managed_by=C.MULTIKUEUE_CONTROLLER)
    submit(op, job)
    time.sleep(0.4)
    assert get(op, job).status.conditions == [] and op.store.list("pods") == [] and op.store.list("jobs") == []


def test_wait_for_workers_ready(op):
    job = new_mpijob("wfw", workers=2, launcher_cmd=("sh", "-c", "true"), launcher_args=None, worker_cmd=("/usr/sbin/sshd",))
    job.spec.launcher_creation_policy = C.LAUNCHER_CREATION_POLICY_WAIT_FOR_WORKERS_READY
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", what="Succeeded")
    lj = op.store.get("jobs", "default", "wfw-launcher")
    workers = [p for p in op.store.list("pods", "default") if "worker" in p["metadata"]["name"]]
    # the launcher Job was created only after both workers reported Ready
    assert all(lj["metadata"]["creationTimestamp"] >= w["status"]["startTime"] for w in workers) or not workers


def test_gpu_slots_gang_and_release(gang_op):
    op = gang_op
    # 4 GPUs on the box. job A: 2 workers x 2 GPUs (needs all 4); job B: 2 workers x 1 GPU must wait for A.
    a = new_mpijob("a", workers=2, launcher_cmd=("sh", "-c", "sleep 0.8"), launcher_args=None, worker_cmd=("/usr/sbin/sshd",))
    a.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = {"limits": {"nvidia.com/gpu": 2}}
    b = new_mpijob("b", workers=2, launcher_cmd=("sh", "-c", "cat $B200MPI_SLOTS_FILE"), launcher_args=None, worker_cmd=("/usr/sbin/sshd",))
    b.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = {"limits": {"nvidia.com/gpu": 1}}
    for j in (a, b):
        j.spec.run_policy.clean_pod_policy = "All"
    submit(op, a)
    wait_for(lambda: conds(get(op, a)).get("Running") == "True", what="A running")
    assert op.agent.alloc.free_gpus == 0
    pg = op.store.get("volcano-podgroups", "default", "a")
    assert pg["spec"]["minMember"] == 3 and pg["spec"]["minResources"] == {"nvidia.com/gpu": "4"}
    submit(op, b)

    def b_pods():
        return [p for p in op.store.list("pods", "default") if p["metadata"]["name"].startswith("b-")]
    wait_for(lambda: any(c.get("reason") == "Unschedulable" for p in b_pods() for c in (p.get("status") or {}).get("conditions", [])),
             timeout=10, what="B's gang reported Unschedulable")
    bw = b_pods()
    # whole gang pending (a pod created a moment ago may not have a status yet), nothing partially started
    assert bw and all((p.get("status") or {}).get("phase") in (None, "Pending") for p in bw)
    wait_for(lambda: conds(get(op, b)).get("Succeeded") == "True", timeout=30, what="B runs after A released its GPUs")
    launcher = [p for p in op.store.list("pods", "default") if p["metadata"]["name"].startswith("b-launcher")][0]
    assert '"b-worker-0": [' in op.agent.logs("default", launcher["metadata"]["name"])
    wait_for(lambda: op.agent.alloc.free_gpus == 4, what="all GPUs released")
    wait_for(lambda: op.store.list("volcano-podgroups", "default") == [], what="PodGroups deleted on finish-with-cleanup")


def test_pending_gangs_are_admitted_in_priority_order(gang_op):
    """priorityClass orders the pending list (Volcano PodGroup.spec.priorityClassName, podgroup.go:311-334).
    A holds the whole box; `low` is submitted before `high`; when A finishes, `high` must run first."""
    op = gang_op
    for name, value in (("high", 1000), ("low", 1)):
        op.store.create("priorityclasses", {"apiVersion": "scheduling.k8s.io/v1", "kind": "PriorityClass", "metadata": {"name": name}, "value": value})

    def gang(name, cmd, prio=None):
        j = new_mpijob(name, workers=1, launcher_cmd=("sh", "-c", cmd), launcher_args=None, worker_cmd=("/usr/sbin/sshd",), clean="All")
        j.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = {"limits": {"nvidia.com/gpu": 4}}
        if prio:
            j.spec.run_policy.scheduling_policy = SchedulingPolicy(priority_class=prio)
        return j
    a = gang("a", "sleep 2.0")     # holds every GPU long enough for BOTH waiting gangs to be registered, also on a loaded host
    submit(op, a)
    wait_for(lambda: conds(get(op, a)).get("Running") == "True", what="A running")
    low, high = gang("lo", "date +%s.%N", "low"), gang("hi", "date +%s.%N; sleep 0.3", "high")
    submit(op, low)
    time.sleep(0.2)
    submit(op, high)
    for j in (low, high):
        wait_for(lambda: conds(get(op, j)).get("Succeeded") == "True", timeout=40, what=f"{j.name} Succeeded")
    started = {}
    for j in (low, high):
        pod = [p for p in op.store.list("pods", "default") if p["metadata"]["name"].startswith(j.name + "-launcher")]
        started[j.name] = get(op, j).status.start_time if not pod else float(op.agent.logs("default", pod[0]["metadata"]["name"]).split()[0])
    assert started["hi"] < started["lo"], started


def test_unschedulable_min_resources_then_cleared(gang_op):
    op = gang_op
    j = new_mpijob("big", workers=1, launcher_cmd=("sh", "-c", "true"), launcher_args=None, worker_cmd=("/usr/sbin/sshd",))
    j.spec.run_policy.scheduling_policy = SchedulingPolicy(min_resources={"nvidia.com/gpu": "64"})
    submit(op, j)
    time.sleep(0.4)
    assert "Running" not in conds(get(op, j)) and "Succeeded" not in conds(get(op, j))
    cur = get(op, j)
    cur.spec.run_policy.scheduling_policy = None
    op.clientset.kubeflow_v2beta1().mpijobs("default").update(cur)
    wait_for(lambda: conds(get(op, j)).get("Succeeded") == "True", what="Succeeded once minResources cleared")


@needs_native
def test_elastic_scale_updates_discover_hosts_in_launcher(op):
    script = 'for i in $(seq 1 200); do n=$(sh $B200MPI_POD_ROOTFS/etc/mpi/discover_hosts.sh | wc -l); echo hosts=$n; if [ "$n" = "4" ]; then exit 0; fi; sleep 0.05; done; exit 1'
    job = new_mpijob("elastic", workers=2, launcher_cmd=("sh", "-c", script), launcher_args=None, worker_cmd=("/usr/sbin/sshd",))
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Running") == "True", what="Running")
    j = get(op, job)
    j.spec.replica("Worker").replicas = 4
    op.clientset.kubeflow_v2beta1().mpijobs("default").update(j)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", what="launcher saw 4 hosts")
    assert get(op, job).status.replica_statuses["Launcher"].succeeded == 1


def test_active_deadline_and_ttl(op):
    job = new_mpijob("deadline", workers=None, launcher_cmd=("sh", "-c", "sleep 30"), launcher_args=None,
                     active_deadline_seconds=1, ttl_seconds_after_finished=1)
    submit(op, job)
    failed = wait_for(lambda: conds(get(op, job)).get("Failed") == "True" and get(op, job), timeout=15, what="DeadlineExceeded")
    assert [c.reason for c in failed.status.conditions if c.type == "Failed"] == ["DeadlineExceeded"]
    wait_for(lambda: not op.store.list("jobs", "default"), timeout=10, what="launcher Job removed by TTL")


@needs_native
def test_rank_collective_counters_reach_the_metrics_endpoint(op):
    """SURVEY.md §5.5 [NEW]: ranks drop Communicator.stats() into $B200MPI_STATS_DIR, the node agent folds them into
    b200mpi_collective_{calls,bytes}_total when the launcher exits (a CPU stand-in writes the file a GPU rank would)."""
    script = ("import json,os;d=os.environ['B200MPI_STATS_DIR'];os.makedirs(d,exist_ok=True);"
              "json.dump({'rank':0,'world':2,'launches':3,'ops':[{'op':'allreduce_sgd','algo':'nvls','calls':2,'bytes':4096},"
              "{'op':'broadcast','algo':'auto','calls':1,'bytes':128}]},open(os.path.join(d,'stats-rank0-1.json'),'w'))")
    job = new_mpijob("stats", workers=1, launcher_cmd=("python",), launcher_args=("-c", script), worker_cmd=("/usr/sbin/sshd", "-De"))
    before = metrics.collective_bytes.labels(op="allreduce_sgd", algo="nvls")._value.get()
    ar_before = metrics.allreduce_bytes.labels(algo="nvls")._value.get()
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", what="Succeeded")
    assert metrics.collective_bytes.labels(op="allreduce_sgd", algo="nvls")._value.get() - before == 4096
    assert metrics.allreduce_bytes.labels(algo="nvls")._value.get() - ar_before == 4096
    text = metrics.render().decode()
    assert 'b200mpi_collective_calls_total{algo="auto",op="broadcast"}' in text
    launcher = [p for p in op.store.list("pods", "default") if p["metadata"]["labels"][C.JOB_ROLE_LABEL] == "launcher"][0]
    assert os.listdir(os.path.join(op.agent.pod_dir(launcher), "stats")) == ["stats-rank0-1.json.seen"]  # consumed once


def test_metrics_text_has_reference_names(op):
    text = metrics.render().decode()
    for name in ("mpi_operator_jobs_created_total", "mpi_operator_jobs_successful_total", "mpi_operator_jobs_failed_total",
                 "mpi_operator_job_info", "mpi_operator_is_leader"):
        assert name in text


def test_job_duration_histogram_counts_finished_jobs(op):
    """Beyond the reference's five metrics: startTime -> completionTime of finished jobs, by result."""
    def count(result):
        for ln in metrics.render().decode().splitlines():
            if ln.startswith('mpi_operator_job_duration_seconds_count{result="%s"}' % result):
                return float(ln.split()[-1])
        return 0.0
    ok0, bad0 = count("Succeeded"), count("Failed")
    good = new_mpijob("dur-ok", workers=1, launcher_cmd=("true",), worker_cmd=("/usr/sbin/sshd", "-De"))
    bad = new_mpijob("dur-bad", workers=1, launcher_cmd=("false",), worker_cmd=("/usr/sbin/sshd", "-De"))
    bad.spec.run_policy.backoff_limit = 0
    submit(op, good)
    submit(op, bad)
    wait_for(lambda: conds(get(op, good)).get("Succeeded") == "True" and conds(get(op, bad)).get("Failed") == "True", what="both finished")
    assert count("Succeeded") == ok0 + 1 and count("Failed") == bad0 + 1


@needs_native
def test_collective_runtime_is_ld_injected_into_ranks(op):
    shim = os.path.join(REPO, "mpi_operator_b200/lib/libb200mpi_nccl.so")
    if not os.path.exists(shim):
        pytest.skip("shim not built")
    job = new_mpijob("inject", workers=2, launcher_cmd=("mpirun",), launcher_args=("-np", "2", "sh", "-c", "echo preload=$LD_PRELOAD"),
                     worker_cmd=("/usr/sbin/sshd",))
    job.spec.replica("Launcher").template["spec"]["containers"][0]["env"] = [{"name": "B200MPI_INJECT", "value": "1"}]
    submit(op, job)
    wait_for(lambda: conds(get(op, job)).get("Succeeded") == "True", what="Succeeded")
    launcher = [p for p in op.store.list("pods", "default") if "inject-launcher" in p["metadata"]["name"]][0]
    assert op.agent.logs("default", launcher["metadata"]["name"]).count(f"preload={shim}") == 2
    base = new_mpijob("noinject", workers=1, launcher_cmd=("mpirun",), launcher_args=("-np", "1", "sh", "-c", "echo preload=[$LD_PRELOAD]"),
                      worker_cmd=("/usr/sbin/sshd",))
    base.spec.replica("Launcher").template["spec"]["containers"][0]["env"] = [{"name": "B200MPI_INJECT", "value": "1"}, {"name": "B200MPI_ALGO", "value": "nccl"}]
    submit(op, base)
    wait_for(lambda: conds(get(op, base)).get("Succeeded") == "True", what="Succeeded")
    launcher = [p for p in op.store.list("pods", "default") if "noinject-launcher" in p["metadata"]["name"]][0]
    assert "preload=[]" in op.agent.logs("default", launcher["metadata"]["name"])
    # default: on when the box has GPUs (this fixture has fake ones), off with B200MPI_INJECT=0
    for name, env, want in (("dflt", [], f"preload=[{shim}]"), ("optout", [{"name": "B200MPI_INJECT", "value": "0"}], "preload=[]")):
        j = new_mpijob(name, workers=1, launcher_cmd=("mpirun",), launcher_args=("-np", "1", "sh", "-c", "echo preload=[$LD_PRELOAD]"),
                       worker_cmd=("/usr/sbin/sshd",))
        j.spec.replica("Launcher").template["spec"]["containers"][0]["env"] = env
        submit(op, j)
        wait_for(lambda: conds(get(op, j)).get("Succeeded") == "True", what="Succeeded")
        launcher = [p for p in op.store.list("pods", "default") if f"{name}-launcher" in p["metadata"]["name"]][0]
        assert want in op.agent.logs("default", launcher["metadata"]["name"])


def test_daemon_restart_reaps_lost_processes_and_recovers(tmp_path):
    """Persisted store + restarted daemon: idle workers are re-adopted with their GPU slots, the launcher the old
    daemon owned is reaped and marked Failed(DaemonRestarted), the Job controller retries, the job succeeds."""
    marker = tmp_path / "second-run"
    script = f"if [ -e {marker} ]; then exit 0; fi; touch {marker}; sleep 60"
    o1 = Operator(ServerOption(fake_gpus=4, leader_elect=False, state_dir=str(tmp_path / "s")))
    o1.start()
    job = new_mpijob("restart", workers=2, launcher_cmd=("sh", "-c", script), launcher_args=None, worker_cmd=("/usr/sbin/sshd",))
    job.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = {"limits": {"nvidia.com/gpu": 1}}
    job.spec.replica("Launcher").restart_policy = "Never"
    submit(o1, job)
    wait_for(lambda: conds(get(o1, job)).get("Running") == "True", what="Running")
    old_launcher = [p for p in o1.store.list("pods", "default") if "restart-launcher" in p["metadata"]["name"]][0]
    pgid = int(old_launcher["metadata"]["annotations"]["b200mpi.kubeflow.org/pgid"])
    # simulate a daemon crash: threads stop, child processes survive
    o1.controller.stop(); o1.agent._stop.set(); o1.agent._thread.join(2); o1.informers.stop()
    os.kill(pgid, 0)  # the old launcher process is still alive
    o2 = Operator(ServerOption(fake_gpus=4, leader_elect=False, state_dir=str(tmp_path / "s")))
    o2.start()
    try:
        assert o2.agent.alloc.free_gpus == 2  # worker reservations re-adopted
        lost = o2.store.get("pods", "default", old_launcher["metadata"]["name"])
        assert lost["status"]["phase"] == "Failed" and lost["status"]["reason"] == "DaemonRestarted"
        done = wait_for(lambda: conds(get(o2, job)).get("Succeeded") == "True" and get(o2, job), what="Succeeded after restart")
        assert done.status.replica_statuses["Launcher"].failed == 1
        # the orphan was SIGKILLed by the new daemon (it is a zombie child of this test process until reaped)
        old_popen = o1.agent._procs[f"default/{old_launcher['metadata']['name']}"].popen
        assert old_popen.wait(timeout=5) == -9
    finally:
        o2.stop()


def test_pod_logs_are_followed_incrementally_and_rotated(op, monkeypatch, tmp_path):
    """`logs -f` polls with an offset (log_slice) instead of re-reading the file; a running pod's log that outgrows
    B200MPI_POD_LOG_MAX_BYTES is copied to 0.log.1 and truncated in place (kubelet's containerLogMaxSize), the container keeps
    writing and a follower starts over at the new beginning."""
    stop = tmp_path / "stop"
    job = new_mpijob("chatty", workers=1, launcher_cmd=("sh", "-c"), worker_cmd=("/usr/sbin/sshd", "-De"),
                     launcher_args=(f"i=0; while [ ! -f {stop} ]; do echo line-$i-xxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxx; i=$((i+1)); sleep 0.005; done; echo last-line",))
    op.clientset.kubeflow_v2beta1().mpijobs("default").create(job)
    wait_for(lambda: any("launcher" in p["metadata"]["name"] for p in op.store.list("pods", "default")), 10)
    launcher = next(p for p in op.store.list("pods", "default") if "launcher" in p["metadata"]["name"])["metadata"]["name"]
    wait_for(lambda: len(op.agent.logs("default", launcher)) > 2000, 20)
    first, off = op.agent.log_slice("default", launcher, 0)
    assert first.startswith(b"line-0-") and off == len(first)
    again, off2 = op.agent.log_slice("default", launcher, off)
    assert off2 >= off and (again == b"" or again.startswith(b"line-"))          # nothing is delivered twice
    monkeypatch.setenv("B200MPI_POD_LOG_MAX_BYTES", "1500")
    assert op.agent.rotate_logs() == 1
    log_path = os.path.join(op.agent.state_dir, "pods", "default", launcher, "logs", "0.log")
    assert os.path.getsize(log_path + ".1") > 1500
    rotated, off3 = op.agent.log_slice("default", launcher, off2)                # offset beyond the new size: starts over
    assert rotated == b"" or rotated.startswith(b"line-")
    monkeypatch.setenv("B200MPI_POD_LOG_MAX_BYTES", "0")
    wait_for(lambda: os.path.getsize(log_path) > 200, 20)                        # the container kept writing after the truncation
    stop.write_text("x")
    wait_for(lambda: "last-line" in open(log_path).read(), 20)
    assert "line-0-" in open(log_path + ".1").read() and "line-0-" not in open(log_path).read()


def test_gpu_cordon_health_monitor_and_scheduling(op):
    """`mpijobctl cordon <gpu>` / the NVML health monitor (node/health.py): a cordoned GPU receives no new ranks, a job that
    does not fit the remaining GPUs waits, reservations that already hold the GPU keep it, the monitor lifts only its own
    cordons, every change is an Event and a metric."""
    from mpi_operator_b200.controller import metrics
    from mpi_operator_b200.node.health import GpuHealthMonitor
    alloc = op.agent.alloc
    assert alloc.free_gpus == 8
    holder = new_mpijob("holder", workers=2, launcher_cmd=("sleep", "30"), worker_cmd=("/usr/sbin/sshd", "-De"))
    holder.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = {"limits": {"nvidia.com/gpu": 1}}
    op.clientset.kubeflow_v2beta1().mpijobs("default").create(holder)
    wait_for(lambda: alloc.free_gpus == 6, 15)
    held = sorted(g for k in ("default/holder-worker-0", "default/holder-worker-1") for g in (alloc.held(k) or []))
    assert held == [0, 1]
    verdict = {g: None for g in range(8)}
    verdict[1] = "uncorrected ECC errors since the last reset"      # a GPU in use
    verdict[5] = "not reachable through NVML (NVMLError_GpuIsLost)"  # an idle GPU
    mon = GpuHealthMonitor(op.agent, probe=lambda: dict(verdict), interval=0, recorder=op.gpu_health.recorder)
    mon.check_once()
    assert alloc.cordoned == {1: "health: " + verdict[1], 5: "health: " + verdict[5]} and alloc.free_gpus == 5
    assert alloc.held("default/holder-worker-1") == [1]              # the running reservation keeps its GPU
    alloc.cordon(7, "maintenance")                                   # a manual cordon
    assert alloc.free_gpus == 4
    text = metrics.render().decode() if isinstance(metrics.render(), bytes) else metrics.render()
    assert 'b200mpi_gpu_healthy{gpu="5"} 0.0' in text and 'b200mpi_gpu_healthy{gpu="2"} 1.0' in text
    evs = [e for e in op.store.list("events") if e.get("reason") == "GPUUnhealthy"]
    assert len(evs) == 2 and any("GPU 5" in e["message"] for e in evs)
    big = new_mpijob("big", workers=5, launcher_cmd=("true",), worker_cmd=("/usr/sbin/sshd", "-De"))
    big.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = {"limits": {"nvidia.com/gpu": 1}}
    op.clientset.kubeflow_v2beta1().mpijobs("default").create(big)   # needs 5, only 4 usable GPUs are free
    time.sleep(1.0)
    running_big = [p for p in op.store.list("pods", "default") if p["metadata"]["name"].startswith("big-worker") and (p.get("status") or {}).get("phase") == "Running"]
    assert len(running_big) < 5
    verdict[5] = None
    mon.check_once()                                                 # GPU 5 recovered: the monitor lifts ITS cordon, not the manual one
    assert 5 not in alloc.cordoned and alloc.cordoned[7] == "maintenance" and 1 in alloc.cordoned
    wait_for(lambda: get(op, big).status and any(c.type == "Succeeded" and c.status == "True" for c in get(op, big).status.conditions or []), 30)
    assert 7 not in {g for k in list(alloc._held) for g in alloc._held[k]}   # nothing was placed on the manually cordoned GPU
    op.clientset.kubeflow_v2beta1().mpijobs("default").delete("holder")
    op.clientset.kubeflow_v2beta1().mpijobs("default").delete("big")   # cleanPodPolicy None: its workers still hold their GPUs
    wait_for(lambda: alloc.held("default/holder-worker-1") is None and not alloc._held, 15)
    assert 1 not in alloc._free                                      # released while cordoned: stays out of the pool
    verdict[1] = None
    mon.check_once()
    assert 1 in alloc._free and alloc.uncordon(7) and alloc.free_gpus == 8
    assert any(e.get("reason") == "GPUHealthy" for e in op.store.list("events"))


def test_a_pod_the_agent_cannot_digest_fails_alone(op):
    """kubelet's CreateContainerConfigError: a pod spec that cannot be turned into a process (no containers, a GPU quantity that
    is not a number - written past the REST admission, straight into the store) fails or stays Pending with the reason; the
    pods next to it are scheduled and run. One such object used to stop the node agent's whole loop."""
    pods = op.clientset.kube.pods("default") if hasattr(op.clientset, "kube") else None
    mk = lambda name, spec: op.store.create("pods", {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": name, "namespace": "default"}, "spec": spec})  # noqa: E731
    mk("no-containers", {"containers": [], "restartPolicy": "Never"})
    with pytest.raises(Exception, match="quantities must match"):           # the store itself refuses impossible quantities ...
        mk("bad-gpus", {"containers": [{"name": "c", "command": ["true"], "resources": {"limits": {"nvidia.com/gpu": "several"}}}], "restartPolicy": "Never"})
    with pytest.raises(ValueError, match="Invalid value: 'several'"):        # ... and the scheduler would name the field if one got through
        op.agent._gpu_request({"spec": {"containers": [{"resources": {"limits": {"nvidia.com/gpu": "several"}}}]}})
    mk("nameless-env", {"containers": [{"name": "c", "command": ["true"], "env": [{"value": "v"}]}], "restartPolicy": "Never"})
    mk("fine", {"containers": [{"name": "c", "command": ["sh", "-c", "echo ok"], "resources": {"limits": {"nvidia.com/gpu": 1}}}], "restartPolicy": "Never"})
    phase = lambda n: (op.store.get("pods", "default", n).get("status") or {}).get("phase")  # noqa: E731
    wait_for(lambda: phase("fine") == "Succeeded", what="the healthy pod ran")
    wait_for(lambda: phase("no-containers") == "Failed" and phase("nameless-env") == "Failed", what="indigestible pods failed")
    st = op.store.get("pods", "default", "no-containers")["status"]
    assert st["reason"] == "CreateContainerConfigError"
    wait_for(lambda: op.agent.alloc.free_gpus == 8, what="GPU slots returned")
    del pods
