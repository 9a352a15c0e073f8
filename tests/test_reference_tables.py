"""The reference's unit-test TABLES, ported as data (not code): every case below names the Go subtest it mirrors.

* pkg/controller/podgroup_test.go:48-309     TestNewPodGroup            (3 scenarios x 2 schedulers)
* pkg/controller/podgroup_test.go:311-385    TestCalcPriorityClassName  (4 cases)
* pkg/controller/podgroup_test.go:442-801    TestCalculatePGMinResources (4 Volcano + 5 scheduler-plugins cases)
* pkg/controller/podgroup_test.go:929-964    TestReplicasOrder          (2 cases)
* pkg/controller/mpi_job_controller_test.go:737-767,1226-1255   foreign Service (Intel) / foreign worker
* pkg/controller/mpi_job_controller_test.go:1049-1224           resume with an existing launcher (two variants)
* pkg/controller/mpi_job_controller_test.go:1257-1422           launcher active / workers (not) ready / 16 workers
* pkg/controller/mpi_job_controller_test.go:1424-1890           TestNewLauncherAndWorker: defaults, launcher-as-worker, overrides
"""
import copy

import pytest

from helpers import conds, new_mpijob, template
from mpi_operator_b200.api import constants as C
from mpi_operator_b200.api import meta as M
from mpi_operator_b200.api.defaults import set_defaults_mpijob
from mpi_operator_b200.api.types import MPIJob, MPIJobSpec, ReplicaSpec, RunPolicy, SchedulingPolicy
from mpi_operator_b200.client import FakeClientset, SharedInformerFactory
from mpi_operator_b200.controller import builders as B
from mpi_operator_b200.controller.controller import SyncError
from mpi_operator_b200.controller.podgroup import (SchedulerPluginsCtrl, VolcanoCtrl, cal_pg_min_resource,
                                                   calculate_priority_class_name)
from test_controller import Fixture

MIN_RESOURCES = {"cpu": "100", "memory": "512Gi", "example.com/gpu": "40"}        # podgroup_test.go:35-40
MIN_RESOURCES_NO_MIN_MEMBER = {"cpu": "1", "memory": "2Gi"}                       # podgroup_test.go:42-45
QUEUE_ANNOTATION = "scheduling.volcano.sh/queue-name"


def _replica(replicas, cpu, mem, pc="", extra_containers=()):
    cs = [{"name": "main", "resources": {"requests": {"cpu": cpu, "memory": mem}}}]
    for c, m in extra_containers:
        cs.append({"name": "extra", "resources": {"requests": {"cpu": c, "memory": m}}})
    spec = {"containers": cs}
    if pc:
        spec["priorityClassName"] = pc
    return ReplicaSpec(replicas=replicas, template={"spec": spec})


def _pg_job(launcher=None, worker=None, sp=None, annotations=None, law=None):
    specs = {}
    if launcher is not None:
        specs["Launcher"] = launcher
    if worker is not None:
        specs["Worker"] = worker
    return MPIJob(metadata={"name": "test", "namespace": "default", "uid": "uid-test", "annotations": annotations or {}},
                  spec=MPIJobSpec(mpi_replica_specs=specs, run_launcher_as_worker=law, run_policy=RunPolicy(scheduling_policy=sp)))


def _ctrls(priority_classes=()):
    cs = FakeClientset()
    for name, value in priority_classes:
        cs.store.create("priorityclasses", {"apiVersion": "scheduling.k8s.io/v1", "kind": "PriorityClass", "metadata": {"name": name}, "value": value})
    inf = SharedInformerFactory(cs.store)
    pcl = inf.lister_for("priorityclasses")
    v, s = VolcanoCtrl(cs.kube(), inf, pcl), SchedulerPluginsCtrl(cs.kube(), inf, "default-scheduler", pcl)
    inf.start()
    return v, s


# ---------------------------------------------------------------- TestNewPodGroup (podgroup_test.go:48-309) ----
NEW_POD_GROUP = {
    "all schedulingPolicy fields are set": dict(
        job=lambda: _pg_job(_replica(1, "1", "2Gi"), _replica(1000, "10", "20Gi"),
                            SchedulingPolicy(min_available=2, queue="project-y", priority_class="high", min_resources=dict(MIN_RESOURCES),
                                             schedule_timeout_seconds=100), {QUEUE_ANNOTATION: "project-x"}),
        volcano={"minMember": 2, "queue": "project-y", "priorityClassName": "high", "minResources": MIN_RESOURCES},
        sched={"minMember": 2, "minResources": MIN_RESOURCES, "scheduleTimeoutSeconds": 100}),
    "schedulingPolicy is nil": dict(
        job=lambda: _pg_job(_replica(1, "1", "2Gi", pc="high"), _replica(2, "10", "20Gi"), None, {QUEUE_ANNOTATION: "project-x"}),
        volcano={"minMember": 3, "queue": "project-x", "priorityClassName": "high", "minResources": {"cpu": "21", "memory": "42Gi"}},
        sched={"minMember": 3, "scheduleTimeoutSeconds": 0, "minResources": {"cpu": "21", "memory": "42Gi"}}),
    "no worker no MinResources": dict(
        job=lambda: _pg_job(_replica(1, "1", "2Gi"), None,
                            SchedulingPolicy(min_available=1, queue="project-y", priority_class="high", schedule_timeout_seconds=100),
                            {QUEUE_ANNOTATION: "project-x"}, law=True),
        volcano={"minMember": 1, "queue": "project-y", "priorityClassName": "high", "minResources": MIN_RESOURCES_NO_MIN_MEMBER},
        sched={"minMember": 1, "minResources": MIN_RESOURCES_NO_MIN_MEMBER, "scheduleTimeoutSeconds": 100}),
}


@pytest.mark.parametrize("name", list(NEW_POD_GROUP))
def test_new_pod_group_table(name):
    tc = NEW_POD_GROUP[name]
    v, s = _ctrls()
    vpg = v.new_pod_group(tc["job"]())
    assert (vpg["apiVersion"], vpg["kind"], vpg["metadata"]["name"]) == ("scheduling.volcano.sh/v1beta1", "PodGroup", "test")
    assert vpg["spec"] == tc["volcano"]
    spg = s.new_pod_group(tc["job"]())
    assert (spg["apiVersion"], spg["kind"], spg["metadata"]["name"]) == ("scheduling.x-k8s.io/v1alpha1", "PodGroup", "test")
    assert spg["spec"] == tc["sched"]
    for pg in (vpg, spg):   # owned by the MPIJob (the Go test ignores the reference; here it is checked)
        ref = pg["metadata"]["ownerReferences"][0]
        assert (ref["kind"], ref["name"], ref["controller"]) == ("MPIJob", "test", True)


# ------------------------------------------------------ TestCalcPriorityClassName (podgroup_test.go:311-385) ----
@pytest.mark.parametrize("name,replicas,sp,want", [
    ("use schedulingPolicy", {}, SchedulingPolicy(priority_class="high"), "high"),
    ("use launcher", {"Launcher": _replica(1, "1", "1Gi", pc="high"), "Worker": _replica(1, "1", "1Gi", pc="low")}, None, "high"),
    ("use worker", {"Launcher": _replica(1, "1", "1Gi"), "Worker": _replica(1, "1", "1Gi", pc="low")}, None, "low"),
    ("nothing", {"Launcher": _replica(1, "1", "1Gi"), "Worker": _replica(1, "1", "1Gi")}, None, ""),
])
def test_calc_priority_class_name_table(name, replicas, sp, want):
    assert calculate_priority_class_name(replicas, sp) == want


# -------------------------------------------------- TestCalculatePGMinResources (podgroup_test.go:442-801) ----
VOLCANO_MIN_RESOURCES = {
    "minResources is not empty": (None, lambda: _pg_job(sp=SchedulingPolicy(min_resources=dict(MIN_RESOURCES))), (), MIN_RESOURCES),
    "schedulingPolicy is nil": (None, lambda: _pg_job(), (), None),
    "without priorityClass": (3, lambda: _pg_job(_replica(1, "2", "1Gi"), _replica(2, "10", "32Gi")), (), {"cpu": "22", "memory": "65Gi"}),
    "without worker without priorityClass": (1, lambda: _pg_job(_replica(1, "2", "1Gi")), (), {"cpu": "2", "memory": "1Gi"}),
}
SCHED_MIN_RESOURCES = {
    "schedulingPolicy.minResources isn't empty": (None, lambda: _pg_job(sp=SchedulingPolicy(min_resources=dict(MIN_RESOURCES))), (), MIN_RESOURCES),
    "schedulingPolicy.minMember is 0": (0, lambda: _pg_job(), (), None),
    "without priorityClass": (None, lambda: _pg_job(_replica(1, "2", "1Gi"), _replica(2, "10", "32Gi", extra_containers=[("50", "512Gi")])), (),
                              {"cpu": "122", "memory": "1089Gi"}),
    "with non-existence priorityClass": (2, lambda: _pg_job(_replica(1, "2", "2Gi", pc="non-existence"), _replica(2, "5", "16Gi", pc="non-existence")), (),
                                         {"cpu": "7", "memory": "18Gi"}),
    "with existence priorityClass": (2, lambda: _pg_job(_replica(1, "2", "4Gi", pc="high"), _replica(100, "20", "64Gi", pc="low")),
                                     (("high", 100_010), ("low", 10_010)), {"cpu": "22", "memory": "68Gi"}),
}


@pytest.mark.parametrize("name", list(VOLCANO_MIN_RESOURCES))
def test_volcano_min_resources_table(name):
    min_member, job, pcs, want = VOLCANO_MIN_RESOURCES[name]
    v, _ = _ctrls(pcs)
    assert v.calculate_pg_min_resources(min_member, job()) == want


@pytest.mark.parametrize("name", list(SCHED_MIN_RESOURCES))
def test_scheduler_plugins_min_resources_table(name):
    min_member, job, pcs, want = SCHED_MIN_RESOURCES[name]
    _, s = _ctrls(pcs)
    assert s.calculate_pg_min_resources(min_member, job()) == want


# --------------------------------------------------------------- TestReplicasOrder (podgroup_test.go:929-964) ----
def test_replicas_order_is_stable_and_by_priority():
    """`sort.Sort(sort.Reverse(order))` in the reference: higher priority first; on a tie the launcher is trimmed LAST, i.e. with
    minMember below the total the workers give way first (observable through the sums)."""
    v, _ = _ctrls((("p1", 1),))
    # launcher higher priority: launcher (1) + 1 of 2 workers for minMember 2
    job = _pg_job(_replica(1, "1", "1Gi", pc="p1"), _replica(2, "10", "10Gi"))
    assert cal_pg_min_resource(2, job, v.pc_lister) == {"cpu": "11", "memory": "11Gi"}
    # equal priority: same outcome (workers are trimmed, the launcher stays)
    job = _pg_job(_replica(1, "1", "1Gi"), _replica(2, "10", "10Gi"))
    assert cal_pg_min_resource(2, job, v.pc_lister) == {"cpu": "11", "memory": "11Gi"}


# ------------------------------------------------------------- controller scenarios missing from test_controller.py ----
def _seed(f, job, impl_default=True):
    """Objects the controller would have created (the Go tests build them with the builders and seed the informer caches)."""
    j = set_defaults_mpijob(copy.deepcopy(job))
    return j


def test_launcher_service_not_controlled_by_us_intel():
    """TestLauncherServiceNotControlledByUs (:737-767): an Intel job whose Service exists without our owner reference."""
    f = Fixture()
    job = f.create_job(new_mpijob("test", workers=2, impl="Intel"))
    svc = B.new_job_service(set_defaults_mpijob(copy.deepcopy(job)))
    svc["metadata"]["ownerReferences"] = []
    f.cs.store.create("services", svc)
    with pytest.raises(SyncError):
        f.sync(job)
    assert any("ErrResourceExists" in e for e in f.recorder.events)


def test_worker_not_controlled_by_us():
    """TestWorkerNotControlledByUs (:1226-1255): worker pod 0 exists, owned by nobody."""
    f = Fixture()
    job = f.create_job(new_mpijob("test", workers=8))
    w = B.new_worker(set_defaults_mpijob(copy.deepcopy(job)), 0)
    w["metadata"]["ownerReferences"] = []
    f.cs.store.create("pods", w)
    with pytest.raises(SyncError):
        f.sync(job)
    assert any("ErrResourceExists" in e for e in f.recorder.events)


KUEUE_DIRECTIVES = dict(node_selector={"foo": "bar"},
                        tolerations=[{"key": "gpu", "operator": "Equal", "value": "true", "effect": "NoSchedule"}],
                        gates=[{"name": "kueue.x-k8s.io/topology"}], annotations={"kueue.x-k8s.io/workload": "my-workload"})


@pytest.mark.parametrize("launcher_started", [False, True], ids=["TestResumeMPIJobWithExistingLauncher", "TestResumeMPIJobClearsStartTime"])
def test_resume_with_existing_launcher_syncs_scheduling_directives(launcher_started):
    """(:1049-1224) running -> suspended -> resumed with a launcher Job that already exists. Kueue has since injected
    nodeSelector / tolerations / schedulingGates / annotations into the MPIJob's launcher template: the existing launcher is
    updated IN PLACE (not re-created); when it had started before, its status.startTime is cleared through the status
    sub-resource first; 8 workers are created; Suspended=False (MPIJobResumed) and StartTime = now."""
    f = Fixture()
    job = f.create_job(new_mpijob("test", workers=8, suspend=True))
    f.sync(job)                      # creates the suspended launcher, no workers
    lj = f.cs.store.get("jobs", "default", "test-launcher")
    assert lj["spec"]["suspend"] is True
    if launcher_started:
        lj.setdefault("status", {})["startTime"] = "2020-01-01T00:00:00Z"
        f.cs.store.update_status("jobs", lj)
    j = f.get(job)
    t = j.spec.replica("Launcher").template
    t["spec"]["nodeSelector"] = dict(KUEUE_DIRECTIVES["node_selector"])
    t["spec"]["tolerations"] = copy.deepcopy(KUEUE_DIRECTIVES["tolerations"])
    t["spec"]["schedulingGates"] = copy.deepcopy(KUEUE_DIRECTIVES["gates"])
    t.setdefault("metadata", {}).setdefault("annotations", {}).update(KUEUE_DIRECTIVES["annotations"])
    j.spec.run_policy.suspend = False
    f.cs.kubeflow_v2beta1().mpijobs("default").update(j)
    f.clock.set_time(1_700_000_123)
    acts = f.sync(j)
    assert not any(a[:2] == ("create", "jobs") for a in acts)                      # updated in place
    assert ("update", "jobs", "test-launcher", "") in acts
    assert (("update", "jobs", "test-launcher", "status") in acts) == launcher_started
    if launcher_started:
        assert acts.index(("update", "jobs", "test-launcher", "status")) < acts.index(("update", "jobs", "test-launcher", ""))
    assert sorted(a[2] for a in acts if a[:2] == ("create", "pods")) == sorted(f"test-worker-{i}" for i in range(8))
    lj = f.cs.store.get("jobs", "default", "test-launcher")
    tmpl = lj["spec"]["template"]
    assert lj["spec"]["suspend"] is False and "startTime" not in lj.get("status", {})
    assert tmpl["spec"]["nodeSelector"] == KUEUE_DIRECTIVES["node_selector"]
    assert tmpl["spec"]["tolerations"] == KUEUE_DIRECTIVES["tolerations"]
    assert tmpl["spec"]["schedulingGates"] == KUEUE_DIRECTIVES["gates"]
    assert tmpl["metadata"]["annotations"]["kueue.x-k8s.io/workload"] == "my-workload"
    got = f.get(job)
    sus = [c for c in got.status.conditions if c.type == "Suspended"][0]
    assert (sus.status, sus.reason, sus.message) == ("False", "MPIJobResumed", "MPIJob resumed")
    assert got.status.start_time == M.now_rfc3339(1_700_000_123)


def test_launcher_active_worker_not_ready():
    """TestLauncherActiveWorkerNotReady (:1257-1308): launcher pod Running, all 8 workers Pending => no Running condition,
    replicaStatuses Launcher.active=1, Worker.active=0."""
    f = Fixture()
    job = f.create_job(new_mpijob("test", workers=8))
    f.sync(job)
    f.launcher_pod(job, "Running")
    for i in range(8):
        f.set_pod_phase("default", f"test-worker-{i}", "Pending")
    f.sync(job)
    got = f.get(job)
    assert "Running" not in conds(got)
    assert (got.status.replica_statuses["Launcher"].active, got.status.replica_statuses["Worker"].active) == (1, 0)


def test_launcher_active_worker_ready():
    """TestLauncherActiveWorkerReady (:1310-1367): launcher Running and all 8 workers Running => Running=True with the
    reference's message, Worker.active=8."""
    f = Fixture()
    job = f.create_job(new_mpijob("test", workers=8))
    f.sync(job)
    f.launcher_pod(job, "Running")
    for i in range(8):
        f.set_pod_phase("default", f"test-worker-{i}", "Running")
    f.sync(job)
    got = f.get(job)
    run = [c for c in got.status.conditions if c.type == "Running"][0]
    assert (run.status, run.reason, run.message) == ("True", "MPIJobRunning", "MPIJob default/test is running.")
    assert (got.status.replica_statuses["Launcher"].active, got.status.replica_statuses["Worker"].active) == (1, 8)


def test_worker_ready_creates_launcher():
    """TestWorkerReady (:1369-1422): 16 workers Running, no launcher yet => the launcher Job is created, Worker.active=16."""
    f = Fixture()
    job = new_mpijob("test", workers=16)
    job.spec.launcher_creation_policy = C.LAUNCHER_CREATION_POLICY_WAIT_FOR_WORKERS_READY
    job = f.create_job(job)
    f.sync(job)
    for i in range(16):
        f.set_pod_phase("default", f"test-worker-{i}", "Running", ready=True)
    acts = f.sync(job)
    assert ("create", "jobs", "test-launcher", "") in acts
    assert f.get(job).status.replica_statuses["Worker"].active == 16


# --------------------------------------------- TestNewLauncherAndWorker (mpi_job_controller_test.go:1424-1890) ----
def _golden_job(name, ns, **spec):
    launcher = ReplicaSpec(template={"spec": {"containers": [{}]}})
    worker = ReplicaSpec(template={"spec": {"containers": [{}]}})
    return MPIJob(metadata={"name": name, "namespace": ns, "uid": "uid-" + name},
                  spec=MPIJobSpec(mpi_replica_specs={"Launcher": launcher, "Worker": worker}, **spec))


SSH_ITEMS = [{"key": "ssh-privatekey", "path": "id_rsa"}, {"key": "ssh-publickey", "path": "id_rsa.pub"},
             {"key": "ssh-publickey", "path": "authorized_keys"}]
CONFIG_ITEMS = [{"key": "hostfile", "path": "hostfile", "mode": 0o444}, {"key": "discover_hosts.sh", "path": "discover_hosts.sh", "mode": 0o555}]


def test_golden_defaults_full_objects():
    """"defaults" (:1431-1561): the complete launcher Job and worker Pod for an empty job foo/bar."""
    job = set_defaults_mpijob(_golden_job("foo", "bar"))
    lj = B.new_launcher_job(job)
    assert lj["metadata"]["name"] == "foo-launcher" and lj["metadata"]["namespace"] == "bar" and lj["metadata"]["labels"] == {"app": "foo"}
    assert set(lj["spec"]) == {"template"}     # no ttl / deadline / backoffLimit / suspend keys when the run policy is empty
    t = lj["spec"]["template"]
    assert t["metadata"]["labels"] == {C.OPERATOR_NAME_LABEL: "mpi-operator", C.JOB_NAME_LABEL: "foo", C.JOB_ROLE_LABEL: "launcher"}
    assert t["spec"] == {
        "hostname": "foo-launcher", "subdomain": "foo", "restartPolicy": "OnFailure",
        "containers": [{
            "env": [{"name": "K_MPI_JOB_ROLE", "value": "launcher"},
                    {"name": "OMPI_MCA_orte_keep_fqdn_hostnames", "value": "true"},
                    {"name": "OMPI_MCA_orte_default_hostfile", "value": "/etc/mpi/hostfile"},
                    {"name": "OMPI_MCA_plm_rsh_args", "value": "-o ConnectionAttempts=10"},
                    {"name": "OMPI_MCA_orte_set_default_slots", "value": "1"},
                    {"name": "NVIDIA_VISIBLE_DEVICES"}, {"name": "NVIDIA_DRIVER_CAPABILITIES"}],
            "volumeMounts": [{"name": "ssh-auth", "mountPath": "/root/.ssh"}, {"name": "mpi-job-config", "mountPath": "/etc/mpi"}]}],
        "volumes": [{"name": "ssh-auth", "secret": {"secretName": "foo-ssh", "defaultMode": 0o600, "items": SSH_ITEMS}},
                    {"name": "mpi-job-config", "configMap": {"name": "foo-config", "items": CONFIG_ITEMS}}]}
    w = B.new_worker(job, 0)
    assert w["metadata"]["name"] == "foo-worker-0" and w["metadata"]["namespace"] == "bar"
    assert w["metadata"]["labels"] == {C.OPERATOR_NAME_LABEL: "mpi-operator", C.JOB_NAME_LABEL: "foo", C.JOB_ROLE_LABEL: "worker",
                                        C.REPLICA_INDEX_LABEL: "0"}
    assert w["spec"] == {
        "hostname": "foo-worker-0", "subdomain": "foo", "restartPolicy": "Never",
        "dnsConfig": {"searches": ["foo.bar.svc.cluster.local"]},
        "containers": [{"command": ["/usr/sbin/sshd", "-De"], "env": [{"name": "K_MPI_JOB_ROLE", "value": "worker"}],
                        "volumeMounts": [{"name": "ssh-auth", "mountPath": "/root/.ssh"}]}],
        "volumes": [{"name": "ssh-auth", "secret": {"secretName": "foo-ssh", "defaultMode": 0o600, "items": SSH_ITEMS}}]}


def test_golden_launcher_as_worker_full_objects():
    """"launcher-as-worker" (:1562-1698): index labels shift by one, the launcher keeps its GPUs (no NVIDIA_* blanking)."""
    job = set_defaults_mpijob(_golden_job("foo", "bar", run_launcher_as_worker=True))
    t = B.new_launcher_job(job)["spec"]["template"]
    assert t["metadata"]["labels"] == {C.OPERATOR_NAME_LABEL: "mpi-operator", C.JOB_NAME_LABEL: "foo", C.JOB_ROLE_LABEL: "launcher",
                                        C.REPLICA_INDEX_LABEL: "0"}
    assert t["spec"]["containers"][0]["env"] == [
        {"name": "K_MPI_JOB_ROLE", "value": "launcher"},
        {"name": "OMPI_MCA_orte_keep_fqdn_hostnames", "value": "true"},
        {"name": "OMPI_MCA_orte_default_hostfile", "value": "/etc/mpi/hostfile"},
        {"name": "OMPI_MCA_plm_rsh_args", "value": "-o ConnectionAttempts=10"},
        {"name": "OMPI_MCA_orte_set_default_slots", "value": "1"}]
    w = B.new_worker(job, 0)
    assert w["metadata"]["labels"][C.REPLICA_INDEX_LABEL] == "1" and w["metadata"]["name"] == "foo-worker-0"
    assert B.new_job_service(job)["spec"]["publishNotReadyAddresses"] is True


def test_golden_overrides_full_objects():
    """"overrides" (:1699-1882): Intel, slots 5, custom ssh path (no 0600 mode), run-policy fields copied to the Job, hostNetwork
    => ClusterFirstWithHostNet, user labels / env / extra containers preserved, only container[0] decorated."""
    job = _golden_job("bar", "foo", slots_per_worker=5, mpi_implementation="Intel", ssh_auth_mount_path="/home/mpiuser/.ssh",
                      run_policy=RunPolicy(ttl_seconds_after_finished=1, active_deadline_seconds=2, backoff_limit=3, suspend=True))
    l = job.spec.replica("Launcher")
    l.restart_policy = "Never"
    l.template = {"metadata": {"labels": {"foo": "bar"}},
                  "spec": {"hostNetwork": True, "containers": [{"env": [{"name": "FOO", "value": "bar"}],
                                                                "securityContext": {"runAsUser": 1000},
                                                                "volumeMounts": [{"name": "fool-vol", "mountPath": "/mnt/foo"}]},
                                                               {}],
                           "volumes": [{"name": "foo-vol"}]}}
    w = job.spec.replica("Worker")
    w.template = {"metadata": {"labels": {"foo": "bar"}},
                  "spec": {"hostNetwork": True, "containers": [{"command": ["/entrypoint.sh"], "env": [{"name": "FOO", "value": "bar"}],
                                                                "securityContext": {"runAsUser": 1000},
                                                                "volumeMounts": [{"name": "fool-vol", "mountPath": "/mnt/foo"}]},
                                                               {}],
                           "volumes": [{"name": "foo-vol"}]}}
    set_defaults_mpijob(job)
    lj = B.new_launcher_job(job)
    assert {k: lj["spec"][k] for k in ("ttlSecondsAfterFinished", "activeDeadlineSeconds", "backoffLimit", "suspend")} == {
        "ttlSecondsAfterFinished": 1, "activeDeadlineSeconds": 2, "backoffLimit": 3, "suspend": True}
    t = lj["spec"]["template"]
    assert t["metadata"]["labels"] == {"foo": "bar", C.OPERATOR_NAME_LABEL: "mpi-operator", C.JOB_NAME_LABEL: "bar", C.JOB_ROLE_LABEL: "launcher"}
    s = t["spec"]
    assert (s["hostNetwork"], s["dnsPolicy"], s["hostname"], s["subdomain"], s["restartPolicy"]) == (True, "ClusterFirstWithHostNet", "bar-launcher", "bar", "Never")
    c0 = s["containers"][0]
    assert c0["securityContext"] == {"runAsUser": 1000}
    assert c0["env"] == [{"name": "FOO", "value": "bar"}, {"name": "K_MPI_JOB_ROLE", "value": "launcher"},
                         {"name": "I_MPI_HYDRA_HOST_FILE", "value": "/etc/mpi/hostfile"},
                         {"name": "I_MPI_HYDRA_BOOTSTRAP_EXEC_EXTRA_ARGS", "value": "-o ConnectionAttempts=10"},
                         {"name": "I_MPI_PERHOST", "value": "5"},
                         {"name": "NVIDIA_VISIBLE_DEVICES"}, {"name": "NVIDIA_DRIVER_CAPABILITIES"}]
    assert c0["volumeMounts"] == [{"name": "fool-vol", "mountPath": "/mnt/foo"}, {"name": "ssh-auth", "mountPath": "/home/mpiuser/.ssh"},
                                  {"name": "mpi-job-config", "mountPath": "/etc/mpi"}]
    assert s["containers"][1] == {}
    assert s["volumes"] == [{"name": "foo-vol"},
                            {"name": "ssh-auth", "secret": {"secretName": "bar-ssh", "items": SSH_ITEMS}},     # no defaultMode: custom ssh path
                            {"name": "mpi-job-config", "configMap": {"name": "bar-config", "items": CONFIG_ITEMS}}]
    wp = B.new_worker(job, 12)
    assert wp["metadata"]["name"] == "bar-worker-12"
    assert wp["metadata"]["labels"] == {"foo": "bar", C.OPERATOR_NAME_LABEL: "mpi-operator", C.JOB_NAME_LABEL: "bar", C.JOB_ROLE_LABEL: "worker",
                                         C.REPLICA_INDEX_LABEL: "12"}
    ws = wp["spec"]
    assert (ws["hostNetwork"], ws["dnsPolicy"], ws["hostname"], ws["subdomain"], ws["restartPolicy"]) == (True, "ClusterFirstWithHostNet", "bar-worker-12", "bar", "Never")
    assert ws["containers"][0]["command"] == ["/entrypoint.sh"]
    assert ws["containers"][0]["env"] == [{"name": "FOO", "value": "bar"}, {"name": "K_MPI_JOB_ROLE", "value": "worker"}]
    assert ws["containers"][0]["volumeMounts"] == [{"name": "fool-vol", "mountPath": "/mnt/foo"}, {"name": "ssh-auth", "mountPath": "/home/mpiuser/.ssh"}]
    assert ws["containers"][1] == {}
    assert ws["volumes"] == [{"name": "foo-vol"}, {"name": "ssh-auth", "secret": {"secretName": "bar-ssh", "items": SSH_ITEMS}}]
