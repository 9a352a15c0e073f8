"""Gang-scheduling math and PodGroup objects (reference tables:
pkg/controller/podgroup_test.go:48-964)."""
from helpers import new_mpijob, template
from mpi_operator_b200.api import constants as C
from mpi_operator_b200.api.defaults import set_defaults_mpijob
from mpi_operator_b200.api.quantity import Quantity
from mpi_operator_b200.api.types import ReplicaSpec, SchedulingPolicy
from mpi_operator_b200.client import FakeClientset, SharedInformerFactory
from mpi_operator_b200.controller.podgroup import (SchedulerPluginsCtrl, VolcanoCtrl, add_resources, cal_pg_min_resource,
                                                   calculate_min_available, calculate_priority_class_name)


def _ctrls():
    cs = FakeClientset()
    inf = SharedInformerFactory(cs.store)
    pcl = inf.lister_for("priorityclasses")
    v, s = VolcanoCtrl(cs.kube(), inf, pcl), SchedulerPluginsCtrl(cs.kube(), inf, "default-scheduler", pcl)
    inf.start()
    return cs, v, s


def _res(cpu=None, mem=None, gpu=None, kind="requests"):
    d = {}
    if cpu: d["cpu"] = cpu
    if mem: d["memory"] = mem
    if gpu: d["nvidia.com/gpu"] = gpu
    return {kind: d}


def _job(workers=1000, lres=None, wres=None, sp=None, lpc="", wpc=""):
    job = new_mpijob("test", workers=workers)
    job.spec.replica("Launcher").template["spec"]["containers"][0]["resources"] = lres or _res("1", "2Gi")
    job.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = wres or _res("10", "32Gi")
    if lpc: job.spec.replica("Launcher").template["spec"]["priorityClassName"] = lpc
    if wpc: job.spec.replica("Worker").template["spec"]["priorityClassName"] = wpc
    job.spec.run_policy.scheduling_policy = sp
    return set_defaults_mpijob(job)


def test_new_pod_group_with_scheduling_policy():
    cs, v, s = _ctrls()
    sp = SchedulingPolicy(min_available=2, queue="project-y", priority_class="high", min_resources={"cpu": "100", "memory": "512Gi"},
                          schedule_timeout_seconds=100)
    job = _job(sp=sp)
    pg = v.new_pod_group(job)
    assert pg["apiVersion"] == "scheduling.volcano.sh/v1beta1" and pg["metadata"]["name"] == "test"
    assert pg["spec"] == {"minMember": 2, "queue": "project-y", "priorityClassName": "high", "minResources": {"cpu": "100", "memory": "512Gi"}}
    pg = s.new_pod_group(job)
    assert pg["apiVersion"] == "scheduling.x-k8s.io/v1alpha1"
    assert pg["spec"] == {"minMember": 2, "scheduleTimeoutSeconds": 100, "minResources": {"cpu": "100", "memory": "512Gi"}}


def test_new_pod_group_defaults():
    cs, v, s = _ctrls()
    job = _job(workers=2, lpc="high")
    job.metadata["annotations"] = {C.VOLCANO_QUEUE_NAME_ANNOTATION: "project-x"}
    pg = v.new_pod_group(job)
    assert pg["spec"]["minMember"] == 3 and pg["spec"]["queue"] == "project-x" and pg["spec"]["priorityClassName"] == "high"
    assert pg["spec"]["minResources"] == {"cpu": "21", "memory": "66Gi"}
    pg = s.new_pod_group(job)
    assert pg["spec"] == {"minMember": 3, "scheduleTimeoutSeconds": 0, "minResources": {"cpu": "21", "memory": "66Gi"}}
    assert pg["metadata"]["ownerReferences"][0]["name"] == "test"


def test_priority_class_precedence():
    r = {"Launcher": ReplicaSpec(template=template(priorityClassName="l")), "Worker": ReplicaSpec(template=template(priorityClassName="w"))}
    assert calculate_priority_class_name(r, SchedulingPolicy(priority_class="p")) == "p"
    assert calculate_priority_class_name(r, None) == "l"
    r["Launcher"] = ReplicaSpec(template=template())
    assert calculate_priority_class_name(r, SchedulingPolicy()) == "w"
    assert calculate_priority_class_name({}, None) == ""


def test_decorate_pod_template():
    cs, v, s = _ctrls()
    t = {"spec": {"schedulerName": "default-scheduler", "containers": [{}]}}
    v.decorate_pod_template_spec(t, "test-mpijob")
    assert t["spec"]["schedulerName"] == "volcano" and t["metadata"]["annotations"] == {C.VOLCANO_GROUP_NAME_ANNOTATION: "test-mpijob"}
    t = {"spec": {"containers": [{}]}}
    s.decorate_pod_template_spec(t, "test-mpijob")
    assert t["spec"]["schedulerName"] == "default-scheduler" and t["metadata"]["labels"] == {C.SCHED_PLUGINS_POD_GROUP_LABEL: "test-mpijob"}


def test_min_resources_trims_workers_when_priorities_tie():
    # minMember 3 over launcher + 1000 workers => launcher + 2 workers
    job = _job(workers=1000, sp=SchedulingPolicy(min_available=3))
    assert cal_pg_min_resource(3, job, None) == {"cpu": "21", "memory": "66Gi"}


def test_min_resources_priority_ordering():
    cs, v, s = _ctrls()
    cs.store.create("priorityclasses", {"apiVersion": "scheduling.k8s.io/v1", "kind": "PriorityClass", "metadata": {"name": "high"}, "value": 100})
    cs.store.create("priorityclasses", {"apiVersion": "scheduling.k8s.io/v1", "kind": "PriorityClass", "metadata": {"name": "low"}, "value": 1})
    # workers outrank the launcher: all minMember-? slots go to workers first, launcher trimmed to minMember-1... per reference: order[1] gets minMember-1
    job = _job(workers=2, lpc="low", wpc="high")
    got = cal_pg_min_resource(2, job, v.pc_lister)
    assert got == {"cpu": "21", "memory": "66Gi"}  # workers (2) kept, launcher replicas -> minMember-1 = 1
    job = _job(workers=4, lpc="high", wpc="low")
    assert cal_pg_min_resource(3, job, v.pc_lister) == {"cpu": "21", "memory": "66Gi"}  # launcher + 2 workers
    # unknown priority class is ignored (priority 0)
    job = _job(workers=2, lpc="does-not-exist")
    assert cal_pg_min_resource(3, job, v.pc_lister) == {"cpu": "21", "memory": "66Gi"}


def test_min_resources_policy_and_zero_member():
    cs, v, s = _ctrls()
    job = _job(workers=2, sp=SchedulingPolicy(min_resources={"cpu": "5"}))
    assert v.calculate_pg_min_resources(3, job) == {"cpu": "5"}
    assert s.calculate_pg_min_resources(0, _job(workers=2)) is None


def test_add_resources_requests_over_limits():
    total = {}
    add_resources(total, {"requests": {"cpu": "1"}, "limits": {"cpu": "4", "memory": "1Gi", "nvidia.com/gpu": "2"}}, 3)
    assert {k: str(v) for k, v in total.items()} == {"cpu": "3", "memory": "3Gi", "nvidia.com/gpu": "6"}
    add_resources(total, {}, 5)
    add_resources(total, None, 5)
    add_resources(total, {"limits": {"cpu": "500m"}}, 1)
    assert str(total["cpu"]) == "3500m"


def test_min_available():
    assert calculate_min_available(_job(workers=2)) == 3
    assert calculate_min_available(_job(workers=2, sp=SchedulingPolicy(min_available=7))) == 7
    j = new_mpijob("x", workers=None)
    assert calculate_min_available(set_defaults_mpijob(j)) == 1


def test_quantity_roundtrip():
    for s in ["1", "100m", "2Gi", "512Mi", "10", "1500m", "3k"]:
        assert str(Quantity.parse(s)) == s
    assert str(Quantity.parse("1Gi") + Quantity.parse("1Gi")) == "2Gi"
    assert str(Quantity.parse("250m") * 4) == "1"
