"""Condition semantics (mpi_job_controller_status.go:99-144) and golden
hostfile / discover_hosts.sh / pod objects (mpi_job_controller_test.go:1424-2402)."""
import base64

import pytest

from helpers import new_mpijob, template
from mpi_operator_b200.api import constants as C
from mpi_operator_b200.api.defaults import set_defaults_mpijob
from mpi_operator_b200.api.types import JobStatus, MPIJob, MPIJobSpec
from mpi_operator_b200.controller import builders as B
from mpi_operator_b200.controller import status as S


def test_condition_update_rules():
    st = JobStatus()
    j = MPIJob(status=st)
    assert S.update_mpijob_conditions(j, "Created", "True", "MPIJobCreated", "m", "2020-01-01T00:00:00Z")
    assert not S.update_mpijob_conditions(j, "Created", "True", "MPIJobCreated", "other message", "2021-01-01T00:00:00Z")
    # same status, new reason -> lastTransitionTime preserved, lastUpdateTime moves
    assert S.update_mpijob_conditions(j, "Created", "True", "Other", "m", "2022-01-01T00:00:00Z")
    c = S.get_condition(j.status, "Created")
    assert (c.last_transition_time, c.last_update_time) == ("2020-01-01T00:00:00Z", "2022-01-01T00:00:00Z")
    S.update_mpijob_conditions(j, "Running", "True", "MPIJobRunning", "r")
    S.update_mpijob_conditions(j, "Succeeded", "True", "MPIJobSucceeded", "s")
    assert S.get_condition(j.status, "Running").status == "False"  # flipped by Succeeded
    assert S.is_finished(j.status) and S.is_succeeded(j.status) and not S.is_failed(j.status)
    S.update_mpijob_conditions(j, "Restarting", "True", "x", "y")
    assert S.get_condition(j.status, "Running") is None  # Running and Restarting are exclusive


def _job(name, ns, impl, slots=None, law=None):
    return MPIJob(metadata={"name": name, "namespace": ns},
                  spec=MPIJobSpec(mpi_implementation=impl, slots_per_worker=slots, run_launcher_as_worker=law))


HOSTFILE_CASES = [
    (_job("openmpi-without-slots", "tenant-a", "OpenMPI", law=True), 2, "",
     "openmpi-without-slots-launcher.openmpi-without-slots.tenant-a.svc slots=1\nopenmpi-without-slots-worker-0.openmpi-without-slots.tenant-a.svc slots=1\nopenmpi-without-slots-worker-1.openmpi-without-slots.tenant-a.svc slots=1\n"),
    (_job("openmpi-without-slots", "tenant-a", "OpenMPI", law=True), 2, "cluster.local",
     "openmpi-without-slots-launcher.openmpi-without-slots.tenant-a.svc.cluster.local slots=1\nopenmpi-without-slots-worker-0.openmpi-without-slots.tenant-a.svc.cluster.local slots=1\nopenmpi-without-slots-worker-1.openmpi-without-slots.tenant-a.svc.cluster.local slots=1\n"),
    (_job("openmpi-without-slots", "tenant-a", "OpenMPI", law=True), 0, "", "openmpi-without-slots-launcher.openmpi-without-slots.tenant-a.svc slots=1\n"),
    (_job("openmpi-without-slots", "tenant-a", "OpenMPI", law=False), 2, "",
     "openmpi-without-slots-worker-0.openmpi-without-slots.tenant-a.svc slots=1\nopenmpi-without-slots-worker-1.openmpi-without-slots.tenant-a.svc slots=1\n"),
    (_job("openmpi-with-slots", "tenant-a", "OpenMPI", slots=10), 1, "", "openmpi-with-slots-worker-0.openmpi-with-slots.tenant-a.svc slots=10\n"),
    (_job("openmpi-with-slots", "tenant-a", "OpenMPI"), 1, "cluster.local", "openmpi-with-slots-worker-0.openmpi-with-slots.tenant-a.svc.cluster.local slots=1\n"),
    (_job("intelmpi-with-slots", "project-x", "Intel", slots=10), 1, "", "intelmpi-with-slots-worker-0.intelmpi-with-slots.project-x.svc:10\n"),
    (_job("intelmpi-with-slots", "project-x", "Intel"), 1, "cluster.local", "intelmpi-with-slots-worker-0.intelmpi-with-slots.project-x.svc.cluster.local:1\n"),
    (_job("mpich-with-slots", "project-x", "MPICH", slots=10), 1, "", "mpich-with-slots-worker-0.mpich-with-slots.project-x.svc:10\n"),
    (_job("mpich-with-slots", "project-x", "MPICH"), 1, "cluster.local", "mpich-with-slots-worker-0.mpich-with-slots.project-x.svc.cluster.local:1\n"),
]


@pytest.mark.parametrize("job,replicas,domain,want", HOSTFILE_CASES)
def test_new_config_map_hostfile(job, replicas, domain, want):
    cm = B.new_config_map(job, replicas, domain)
    assert cm["data"] == {"hostfile": want}
    assert cm["metadata"]["name"] == job.name + "-config" and cm["metadata"]["labels"] == {"app": job.name}
    assert cm["metadata"]["ownerReferences"][0]["kind"] == "MPIJob" and cm["metadata"]["ownerReferences"][0]["controller"] is True


def _pod(name, ns="default"):
    return {"metadata": {"name": name, "namespace": ns}}


DISCOVER_CASES = [
    (_job("test-job", "default", "OpenMPI"), [], "", "#!/bin/sh\n"),
    (_job("test-job", "default", "OpenMPI"), [_pod("test-job-worker-0"), _pod("test-job-worker-1")], "",
     "#!/bin/sh\necho test-job-worker-0.test-job.default.svc\necho test-job-worker-1.test-job.default.svc\n"),
    (_job("test-job", "default", "OpenMPI", law=True), [], "", "#!/bin/sh\necho test-job-launcher.test-job.default.svc\n"),
    (_job("test-job", "default", "OpenMPI", law=True), [_pod("test-job-worker-0"), _pod("test-job-worker-1")], "",
     "#!/bin/sh\necho test-job-launcher.test-job.default.svc\necho test-job-worker-0.test-job.default.svc\necho test-job-worker-1.test-job.default.svc\n"),
    (_job("test-job", "tenant-a", "OpenMPI"), [_pod("test-job-worker-0", "tenant-a")], "cluster.local",
     "#!/bin/sh\necho test-job-worker-0.test-job.tenant-a.svc.cluster.local\n"),
    (_job("test-job", "tenant-a", "OpenMPI", law=True), [_pod("test-job-worker-0", "tenant-a")], "cluster.local",
     "#!/bin/sh\necho test-job-launcher.test-job.tenant-a.svc.cluster.local\necho test-job-worker-0.test-job.tenant-a.svc.cluster.local\n"),
    (_job("test-job", "default", "OpenMPI"), [_pod("test-job-worker-2"), _pod("test-job-worker-0"), _pod("test-job-worker-1")], "",
     "#!/bin/sh\necho test-job-worker-0.test-job.default.svc\necho test-job-worker-1.test-job.default.svc\necho test-job-worker-2.test-job.default.svc\n"),
]


@pytest.mark.parametrize("job,pods,domain,want", DISCOVER_CASES)
def test_discover_hosts(job, pods, domain, want):
    cm = B.new_config_map(job, 0, domain)
    B.update_discover_hosts_in_config_map(cm, job, pods, domain)
    assert cm["data"]["discover_hosts.sh"] == want


def test_golden_worker_and_launcher_defaults():
    job = set_defaults_mpijob(new_mpijob("foo", "bar", workers=2, launcher_cmd=None, launcher_args=None))
    w = B.new_worker(job, 1)
    assert w["metadata"]["name"] == "foo-worker-1"
    assert w["metadata"]["labels"] == {C.OPERATOR_NAME_LABEL: "mpi-operator", C.JOB_NAME_LABEL: "foo", C.JOB_ROLE_LABEL: "worker",
                                        C.REPLICA_INDEX_LABEL: "1"}
    s = w["spec"]
    assert (s["hostname"], s["subdomain"], s["restartPolicy"]) == ("foo-worker-1", "foo", "Never")
    assert s["dnsConfig"] == {"searches": ["foo.bar.svc.cluster.local"]}
    c = s["containers"][0]
    assert c["command"] == ["/usr/sbin/sshd", "-De"]  # only when command and args are both empty
    assert c["env"] == [{"name": "K_MPI_JOB_ROLE", "value": "worker"}]
    assert c["volumeMounts"] == [{"name": "ssh-auth", "mountPath": "/root/.ssh"}]
    assert s["volumes"] == [{"name": "ssh-auth", "secret": {"secretName": "foo-ssh", "defaultMode": 0o600, "items": [
        {"key": "ssh-privatekey", "path": "id_rsa"}, {"key": "ssh-publickey", "path": "id_rsa.pub"},
        {"key": "ssh-publickey", "path": "authorized_keys"}]}}]
    lj = B.new_launcher_job(job)
    assert lj["metadata"]["name"] == "foo-launcher" and lj["metadata"]["labels"] == {"app": "foo"}
    t = lj["spec"]["template"]
    assert t["metadata"]["labels"] == {C.OPERATOR_NAME_LABEL: "mpi-operator", C.JOB_NAME_LABEL: "foo", C.JOB_ROLE_LABEL: "launcher"}
    assert (t["spec"]["hostname"], t["spec"]["subdomain"], t["spec"]["restartPolicy"]) == ("foo-launcher", "foo", "OnFailure")
    assert t["spec"]["containers"][0]["env"] == [
        {"name": "K_MPI_JOB_ROLE", "value": "launcher"},
        {"name": "OMPI_MCA_orte_keep_fqdn_hostnames", "value": "true"},
        {"name": "OMPI_MCA_orte_default_hostfile", "value": "/etc/mpi/hostfile"},
        {"name": "OMPI_MCA_plm_rsh_args", "value": "-o ConnectionAttempts=10"},
        {"name": "OMPI_MCA_orte_set_default_slots", "value": "1"},
        {"name": "NVIDIA_VISIBLE_DEVICES"}, {"name": "NVIDIA_DRIVER_CAPABILITIES"}]
    assert t["spec"]["containers"][0]["volumeMounts"] == [{"name": "ssh-auth", "mountPath": "/root/.ssh"},
                                                          {"name": "mpi-job-config", "mountPath": "/etc/mpi"}]
    assert t["spec"]["volumes"][-1] == {"name": "mpi-job-config", "configMap": {"name": "foo-config", "items": [
        {"key": "hostfile", "path": "hostfile", "mode": 0o444}, {"key": "discover_hosts.sh", "path": "discover_hosts.sh", "mode": 0o555}]}}
    assert "ttlSecondsAfterFinished" not in lj["spec"] and "suspend" not in lj["spec"]


def test_golden_launcher_as_worker_and_overrides():
    job = new_mpijob("foo", "bar", workers=3, impl="Intel", slots=5)
    job.spec.run_launcher_as_worker = True
    job.spec.ssh_auth_mount_path = "/home/mpiuser/.ssh"
    job.spec.run_policy.ttl_seconds_after_finished, job.spec.run_policy.active_deadline_seconds = 1, 2
    job.spec.run_policy.backoff_limit, job.spec.run_policy.suspend = 3, True
    l = job.spec.replica("Launcher")
    l.restart_policy = "Never"
    l.template["metadata"] = {"labels": {"foo": "bar"}}
    l.template["spec"]["hostNetwork"] = True
    l.template["spec"]["containers"][0]["env"] = [{"name": "FOO", "value": "bar"}]
    l.template["spec"]["containers"].append({"name": "sidecar"})
    w = job.spec.replica("Worker")
    w.template["spec"]["hostNetwork"] = True
    w.template["spec"]["containers"][0]["command"] = ["/entrypoint.sh"]
    w.restart_policy = "ExitCode"
    set_defaults_mpijob(job)
    lj = B.new_launcher_job(job)
    assert (lj["spec"]["ttlSecondsAfterFinished"], lj["spec"]["activeDeadlineSeconds"], lj["spec"]["backoffLimit"], lj["spec"]["suspend"]) == (1, 2, 3, True)
    t = lj["spec"]["template"]
    assert t["metadata"]["labels"]["foo"] == "bar" and t["metadata"]["labels"][C.REPLICA_INDEX_LABEL] == "0"
    assert t["spec"]["dnsPolicy"] == "ClusterFirstWithHostNet" and t["spec"]["restartPolicy"] == "Never"
    env = t["spec"]["containers"][0]["env"]
    assert env[0] == {"name": "FOO", "value": "bar"} and env[1] == {"name": "K_MPI_JOB_ROLE", "value": "launcher"}
    assert {"name": "I_MPI_PERHOST", "value": "5"} in env and {"name": "I_MPI_HYDRA_HOST_FILE", "value": "/etc/mpi/hostfile"} in env
    assert not any(e["name"].startswith("NVIDIA_") for e in env)  # launcher-as-worker keeps its GPUs
    assert t["spec"]["containers"][1] == {"name": "sidecar"}  # only container[0] is decorated
    assert "defaultMode" not in t["spec"]["volumes"][0]["secret"]  # 0600 only for /root/.ssh
    wp = B.new_worker(job, 0)
    assert wp["metadata"]["labels"][C.REPLICA_INDEX_LABEL] == "1"  # padded by one
    assert wp["spec"]["restartPolicy"] == "Never" and wp["spec"]["containers"][0]["command"] == ["/entrypoint.sh"]
    assert B.new_job_service(job)["spec"] == {"clusterIP": "None", "publishNotReadyAddresses": True, "selector": {
        C.OPERATOR_NAME_LABEL: "mpi-operator", C.JOB_NAME_LABEL: "foo"}}


def test_mpich_env_and_ssh_secret():
    job = set_defaults_mpijob(new_mpijob("m", "ns", impl="MPICH"))
    env = B.new_launcher_job(job)["spec"]["template"]["spec"]["containers"][0]["env"]
    assert {"name": "HYDRA_HOST_FILE", "value": "/etc/mpi/hostfile"} in env
    assert {"name": "HYDRA_LAUNCH_EXTRA_ARGS", "value": "-o ConnectionAttempts=10"} in env
    assert not any(e["name"] in ("I_MPI_PERHOST", "OMPI_MCA_orte_set_default_slots") for e in env)
    sec = B.new_ssh_auth_secret(job)
    assert sec["type"] == "kubernetes.io/ssh-auth" and sorted(sec["data"]) == ["ssh-privatekey", "ssh-publickey"]
    assert base64.b64decode(sec["data"]["ssh-privatekey"]).startswith(b"-----BEGIN EC PRIVATE KEY-----")
    assert base64.b64decode(sec["data"]["ssh-publickey"]).startswith(b"ecdsa-sha2-nistp521 ")


def test_sync_launcher_scheduling_directives():
    launcher = {"spec": {"template": {"metadata": {"labels": {"a": "1"}}, "spec": {"nodeSelector": {"old": "x"}, "containers": [{}]}}}}
    desired = {"metadata": {"labels": {"b": "2"}, "annotations": {"k": "v"}},
               "spec": {"nodeSelector": {"foo": "bar"}, "tolerations": [{"key": "gpu"}], "schedulingGates": [{"name": "kueue"}]}}
    B.sync_launcher_scheduling_directives(launcher, desired)
    t = launcher["spec"]["template"]
    assert t["metadata"]["labels"] == {"a": "1", "b": "2"} and t["metadata"]["annotations"] == {"k": "v"}
    assert t["spec"]["nodeSelector"] == {"foo": "bar"} and t["spec"]["tolerations"] == [{"key": "gpu"}]
    assert t["spec"]["schedulingGates"] == [{"name": "kueue"}]


def test_restart_policy_replica_level_wins_exit_code_degrades_and_template_value_warns():
    """the replica-level restartPolicy wins; ExitCode -> Never; a restartPolicy in the pod template is
    overridden with a Warning SetPodTemplateRestartPolicy event (controller.go:1609-1616,1693-1699)."""
    from mpi_operator_b200.api.defaults import set_defaults_mpijob
    from mpi_operator_b200.controller.events import FakeRecorder
    job = set_defaults_mpijob(new_mpijob("foo", workers=1))
    job.spec.replica("Launcher").restart_policy = "ExitCode"
    job.spec.replica("Launcher").template["spec"]["restartPolicy"] = "Always"
    job.spec.replica("Worker").restart_policy = "OnFailure"
    rec = FakeRecorder()
    t = B.new_launcher_pod_template(job, recorder=rec)
    assert t["spec"]["restartPolicy"] == "Never"
    assert rec.events == ["Warning SetPodTemplateRestartPolicy Restart policy in pod template overridden by restart policy in replica spec"]
    assert B.new_worker(job, 0)["spec"]["restartPolicy"] == "OnFailure"
    rec2 = FakeRecorder()
    job.spec.replica("Launcher").template["spec"].pop("restartPolicy")
    B.new_launcher_pod_template(job, recorder=rec2)
    assert rec2.events == []
