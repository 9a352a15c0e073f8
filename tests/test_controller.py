"""Controller unit tier: ``sync_handler`` against a fake clientset with recorded
actions — the analogue of the reference's fixture
(pkg/controller/mpi_job_controller_test.go:63-437) and its scenarios
(:493-1422).  No processes, no GPUs: tests play kubelet by writing statuses."""
import copy

import pytest

from helpers import conds, new_mpijob
from mpi_operator_b200.api import constants as C
from mpi_operator_b200.api import meta as M
from mpi_operator_b200.api.types import MPIJob
from mpi_operator_b200.client import FakeClientset, SharedInformerFactory
from mpi_operator_b200.controller import metrics
from mpi_operator_b200.controller.clock import FakeClock
from mpi_operator_b200.controller.controller import MPIJobController, SyncError
from mpi_operator_b200.controller.events import FakeRecorder


class Fixture:
    def __init__(self, gang=""):
        self.cs = FakeClientset()
        self.kube = self.cs.kube()
        self.informers = SharedInformerFactory(self.cs.store)
        self.recorder = FakeRecorder()
        self.clock = FakeClock()
        self.ctrl = MPIJobController(self.kube, self.cs, self.informers, gang_scheduling=gang, recorder=self.recorder, clock=self.clock)
        self.informers.start()

    def create_job(self, job: MPIJob) -> MPIJob:
        out = self.cs.kubeflow_v2beta1().mpijobs(job.namespace).create(job)
        self.cs.clear_actions()
        return out

    def sync(self, job: MPIJob):
        self.cs.clear_actions()
        self.ctrl.sync_handler(job.key)
        return [(a.verb, a.resource, a.name, a.subresource) for a in self.cs.actions if a.verb in ("create", "update", "delete", "patch")]

    def get(self, job: MPIJob) -> MPIJob:
        return self.cs.kubeflow_v2beta1().mpijobs(job.namespace).get(job.name)

    def set_pod_phase(self, ns, name, phase, ready=None, reason=None, message=None):
        p = self.cs.store.get("pods", ns, name)
        p.setdefault("status", {})["phase"] = phase
        if ready is not None:
            p["status"]["conditions"] = [{"type": "Ready", "status": "True" if ready else "False"}]
        if reason:
            p["status"]["reason"] = reason
        if message:
            p["status"]["message"] = message
        self.cs.store.update_status("pods", p)

    def launcher_pod(self, job: MPIJob, phase="Running", name=None, **kw):
        lj = self.cs.store.get("jobs", job.namespace, job.name + "-launcher")
        pod = {"apiVersion": "v1", "kind": "Pod",
               "metadata": {"name": name or job.name + "-launcher-abc12", "namespace": job.namespace,
                            "labels": {"job-name": job.name + "-launcher"}, "ownerReferences": [M.new_controller_ref(lj, "batch/v1", "Job")]},
               "spec": {"containers": [{}]}, "status": {"phase": phase, **kw}}
        return self.cs.store.create("pods", pod)

    def set_launcher_condition(self, job: MPIJob, ctype, reason="", message="", failed=0, completion=None):
        lj = self.cs.store.get("jobs", job.namespace, job.name + "-launcher")
        st = lj.setdefault("status", {})
        st["conditions"] = [{"type": ctype, "status": "True", "reason": reason, "message": message}]
        st["failed"] = failed
        if completion:
            st["completionTime"] = completion
        self.cs.store.update_status("jobs", lj)


def test_invalid_key_missing_job_do_nothing():
    f = Fixture()
    f.ctrl.sync_handler("foo/bar/baz")
    f.ctrl.sync_handler("default/missing")
    assert [a for a in f.cs.actions if a.verb in ("create", "update", "delete")] == []


def test_invalid_job_emits_validation_event_no_requeue():
    f = Fixture()
    job = new_mpijob("1-bad-name")
    job = f.create_job(job)
    assert f.sync(job) == []
    assert len(f.recorder.events) == 1 and f.recorder.events[0].startswith("Warning ValidationError Found validation errors: metadata.name")


def test_externally_managed_job_is_ignored():
    f = Fixture()
    job = f.create_job(new_mpijob("ext", managed_by=C.MULTIKUEUE_CONTROLLER))
    assert f.sync(job) == [] and f.recorder.events == []
    job2 = f.create_job(new_mpijob("own", managed_by=C.KUBEFLOW_JOB_CONTROLLER))
    assert ("create", "jobs", "own-launcher", "") in f.sync(job2)


@pytest.mark.parametrize("impl", ["OpenMPI", "Intel", "MPICH"])
def test_all_resources_created_in_order(impl):
    f = Fixture()
    before = metrics.counter_value(metrics.mpi_jobs_created)
    job = f.create_job(new_mpijob("foo", workers=5, impl=impl))
    acts = f.sync(job)
    assert acts == ([("create", "services", "foo", ""), ("create", "configmaps", "foo-config", ""), ("create", "secrets", "foo-ssh", "")]
                    + [("create", "pods", f"foo-worker-{i}", "") for i in range(5)]
                    + [("create", "jobs", "foo-launcher", ""), ("update", "mpijobs", "foo", "status")])
    got = f.get(job)
    assert conds(got) == {"Created": "True"} and got.status.start_time is not None
    assert got.status.replica_statuses["Launcher"].active == 0
    assert metrics.counter_value(metrics.mpi_jobs_created) == before + 1
    assert f.recorder.events[0] == "Normal MPIJobCreated MPIJob default/foo is created."
    # second sync: nothing left to do
    assert f.sync(job) == []


@pytest.mark.parametrize("res,name,kind", [("jobs", "foo-launcher", "Job"), ("configmaps", "foo-config", "ConfigMap"),
                                           ("services", "foo", "Service"), ("secrets", "foo-ssh", "Secret"), ("pods", "foo-worker-0", "Pod")])
def test_foreign_owned_resource_is_an_error(res, name, kind):
    f = Fixture()
    api_version = {"jobs": "batch/v1"}.get(res, "v1")
    foreign = {"apiVersion": api_version, "kind": kind, "metadata": {"name": name, "namespace": "default"}}
    if res == "pods":
        foreign["spec"] = {"containers": [{}]}
    if res == "jobs":
        foreign["spec"] = {"template": {"spec": {"containers": [{}]}}}
    f.cs.store.create(res, foreign)
    job = f.create_job(new_mpijob("foo"))
    with pytest.raises(SyncError) as ei:
        f.ctrl.sync_handler(job.key)
    assert str(ei.value) == f'Resource "{name}" of Kind "{kind}" already exists and is not managed by MPIJob'
    assert any(e.startswith("Warning ErrResourceExists") for e in f.recorder.events)


def test_launcher_succeeded():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=2))
    f.sync(job)
    f.launcher_pod(job, "Succeeded")
    f.set_launcher_condition(job, "Complete", completion="2024-01-01T00:00:00Z")
    before = metrics.counter_value(metrics.mpi_jobs_successful)
    acts = f.sync(job)
    assert acts == [("update", "mpijobs", "foo", "status")]
    got = f.get(job)
    assert conds(got)["Succeeded"] == "True" and got.status.completion_time == "2024-01-01T00:00:00Z"
    assert got.status.replica_statuses["Launcher"].succeeded == 1
    assert metrics.counter_value(metrics.mpi_jobs_successful) == before + 1
    assert "Normal MPIJobSucceeded MPIJob default/foo successfully completed." in f.recorder.events


def test_launcher_failed_backoff_limit_appends_last_pod_reason():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=0 or None))
    f.sync(job)
    p1 = f.launcher_pod(job, "Failed", name="foo-launcher-1", reason="FailedReason1", message="first message")
    p2 = f.launcher_pod(job, "Failed", name="foo-launcher-2", reason="FailedReason2", message="second message")
    # make pod 2 the most recent
    p2["metadata"]["creationTimestamp"] = "2099-01-01T00:00:00Z"
    f.cs.store._objs["pods"]["default/foo-launcher-2"]["metadata"]["creationTimestamp"] = "2099-01-01T00:00:00Z"
    f.informers.informer_for("pods").indexer.add(f.cs.store.get("pods", "default", "foo-launcher-2"))
    f.set_launcher_condition(job, "Failed", "BackoffLimitExceeded", "Job has reached the specified backoff limit", failed=2)
    f.sync(job)
    got = f.get(job)
    c = [c for c in got.status.conditions if c.type == "Failed"][0]
    assert c.reason == "BackoffLimitExceeded/FailedReason2"
    assert c.message == "Job has reached the specified backoff limit: second message"
    assert got.status.replica_statuses["Launcher"].failed == 2 and got.status.completion_time is not None
    assert any(e.startswith("Warning BackoffLimitExceeded/FailedReason2") for e in f.recorder.events)


def test_finished_job_cleans_up_workers():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=8, clean="All"))
    f.sync(job)
    f.launcher_pod(job, "Succeeded")
    f.set_launcher_condition(job, "Complete", completion="2024-01-01T00:00:00Z")
    f.sync(job)  # marks Succeeded + completion time
    acts = f.sync(job)
    assert [a for a in acts if a[0] == "delete"] == [("delete", "pods", f"foo-worker-{i}", "") for i in range(8)]


def test_clean_pod_policy_running_keeps_finished_pods():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=3, clean="Running"))
    f.sync(job)
    f.set_pod_phase("default", "foo-worker-0", "Running")
    f.set_pod_phase("default", "foo-worker-1", "Succeeded")
    f.set_pod_phase("default", "foo-worker-2", "Pending")
    f.launcher_pod(job, "Succeeded")
    f.set_launcher_condition(job, "Complete", completion="2024-01-01T00:00:00Z")
    f.sync(job)
    acts = f.sync(job)
    assert sorted(a[2] for a in acts if a[0] == "delete") == ["foo-worker-0", "foo-worker-2"]  # Running AND Pending go


def test_create_suspended_then_resume():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=2, suspend=True))
    acts = f.sync(job)
    assert not any(a[1] == "pods" for a in acts)  # no workers while suspended
    lj = f.cs.store.get("jobs", "default", "foo-launcher")
    assert lj["spec"]["suspend"] is True
    got = f.get(job)
    assert conds(got) == {"Created": "True", "Suspended": "True", "Running": "False"} and got.status.start_time is None
    assert "Normal MPIJobSuspended MPIJob suspended" in f.recorder.events
    # the Job controller would have set startTime earlier; resume must clear it through the status sub-resource
    lj.setdefault("status", {})["startTime"] = "2020-01-01T00:00:00Z"
    f.cs.store.update_status("jobs", lj)
    j = f.get(job)
    j.spec.run_policy.suspend = False
    j.spec.replica("Launcher").template["spec"]["nodeSelector"] = {"foo": "bar"}
    j.spec.replica("Launcher").template.setdefault("metadata", {})["annotations"] = {"kueue": "x"}
    f.cs.kubeflow_v2beta1().mpijobs("default").update(j)
    f.clock.set_time(1_700_000_000)
    acts = f.sync(j)
    assert ("update", "jobs", "foo-launcher", "status") in acts and ("update", "jobs", "foo-launcher", "") in acts
    assert acts.index(("update", "jobs", "foo-launcher", "status")) < acts.index(("update", "jobs", "foo-launcher", ""))
    lj = f.cs.store.get("jobs", "default", "foo-launcher")
    assert lj["spec"]["suspend"] is False and "startTime" not in lj.get("status", {})
    assert lj["spec"]["template"]["spec"]["nodeSelector"] == {"foo": "bar"}
    assert lj["spec"]["template"]["metadata"]["annotations"] == {"kueue": "x"}
    got = f.get(job)
    assert conds(got)["Suspended"] == "False" and got.status.start_time == M.now_rfc3339(1_700_000_000)
    assert [c.reason for c in got.status.conditions if c.type == "Suspended"] == ["MPIJobResumed"]
    assert sum(1 for a in acts if a[:2] == ("create", "pods")) == 2


def test_suspend_running_job_deletes_workers():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=2))
    f.sync(job)
    for i in range(2):
        f.set_pod_phase("default", f"foo-worker-{i}", "Running")
    f.launcher_pod(job, "Running")
    f.sync(job)
    assert conds(f.get(job))["Running"] == "True"
    j = f.get(job)
    j.spec.run_policy.suspend = True
    f.cs.kubeflow_v2beta1().mpijobs("default").update(j)
    acts = f.sync(j)
    assert ("update", "jobs", "foo-launcher", "") in acts
    assert sorted(a[2] for a in acts if a[0] == "delete") == ["foo-worker-0", "foo-worker-1"]
    c = conds(f.get(job))
    assert c["Suspended"] == "True" and c["Running"] == "False"


def test_running_condition_needs_launcher_and_all_workers():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=2))
    f.sync(job)
    f.launcher_pod(job, "Running")
    f.set_pod_phase("default", "foo-worker-0", "Running")
    f.sync(job)  # worker-1 still pending
    got = f.get(job)
    assert "Running" not in conds(got) and got.status.replica_statuses["Launcher"].active == 1
    f.set_pod_phase("default", "foo-worker-1", "Running")
    f.sync(job)
    got = f.get(job)
    assert conds(got)["Running"] == "True" and got.status.replica_statuses["Worker"].active == 2
    assert "Normal MPIJobRunning MPIJob default/foo is running" in f.recorder.events


def test_wait_for_workers_ready_gates_launcher():
    f = Fixture()
    job = new_mpijob("foo", workers=16)
    job.spec.launcher_creation_policy = C.LAUNCHER_CREATION_POLICY_WAIT_FOR_WORKERS_READY
    job = f.create_job(job)
    acts = f.sync(job)
    assert not any(a[1] == "jobs" for a in acts)
    for i in range(16):
        f.set_pod_phase("default", f"foo-worker-{i}", "Running", ready=(i != 7))
    assert not any(a[1] == "jobs" for a in f.sync(job))
    f.set_pod_phase("default", "foo-worker-7", "Running", ready=True)
    assert ("create", "jobs", "foo-launcher", "") in f.sync(job)


def test_scale_down_deletes_highest_indices_and_discover_hosts_follows():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=4))
    f.sync(job)
    for i in range(4):
        f.set_pod_phase("default", f"foo-worker-{i}", "Running")
    f.sync(job)
    cm = f.cs.store.get("configmaps", "default", "foo-config")
    assert cm["data"]["discover_hosts.sh"].count("echo ") == 4
    j = f.get(job)
    j.spec.replica("Worker").replicas = 2
    f.cs.kubeflow_v2beta1().mpijobs("default").update(j)
    acts = f.sync(j)
    assert sorted(a[2] for a in acts if a[0] == "delete") == ["foo-worker-2", "foo-worker-3"]
    f.sync(j)
    cm = f.cs.store.get("configmaps", "default", "foo-config")
    assert cm["data"]["hostfile"].count("\n") == 2
    assert cm["data"]["discover_hosts.sh"] == "#!/bin/sh\necho foo-worker-0.foo.default.svc\necho foo-worker-1.foo.default.svc\n"


def test_evicted_worker_fails_job():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=2))
    f.sync(job)
    f.set_pod_phase("default", "foo-worker-1", "Failed", reason="Evicted")
    f.sync(job)
    got = f.get(job)
    c = [c for c in got.status.conditions if c.type == "Failed"][0]
    assert (c.status, c.reason, c.message) == ("True", "MPIJobEvicted", "1/2 workers are evicted")
    assert got.status.replica_statuses["Worker"].failed == 1


def test_owner_hop_pod_to_job_to_mpijob_enqueues():
    f = Fixture()
    job = f.create_job(new_mpijob("foo", workers=1))
    f.sync(job)
    import time
    time.sleep(0.2)                  # let the informer deliver the events of the objects sync() created
    while len(f.ctrl.queue):         # drain AND mark done: a key that is still "processing" would not be handed out again
        k, _ = f.ctrl.queue.get(0.01)
        if k is not None:
            f.ctrl.queue.done(k)
    f.launcher_pod(job, "Running")  # informer -> handle_object -> Job -> MPIJob
    deadline = time.time() + 2
    key = None
    while time.time() < deadline and key is None:
        key, _ = f.ctrl.queue.get(0.05)
    assert key == "default/foo"


def test_truncate_message():
    from mpi_operator_b200.controller.controller import truncate_message
    assert truncate_message("x" * 2000).endswith("...") and len(truncate_message("x" * 2000)) == 1024
    assert truncate_message("short") == "short"
