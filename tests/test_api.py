"""API defaulting + validation, ported as *specs* from the reference tables
(pkg/apis/kubeflow/v2beta1/default_test.go:24-164,
pkg/apis/kubeflow/validation/validation_test.go:29-415)."""
import pytest

from mpi_operator_b200.api import constants as C
from mpi_operator_b200.api import yaml_io
from mpi_operator_b200.api.defaults import set_defaults_mpijob
from mpi_operator_b200.api.types import MPIJob, MPIJobSpec, ReplicaSpec, RunPolicy, SchedulingPolicy
from mpi_operator_b200.api.validation import (ERR_INVALID, ERR_NOT_SUPPORTED, ERR_REQUIRED, validate_mpijob)

ONE_CONTAINER = {"spec": {"containers": [{}]}}


def _spec(**kw):
    base = dict(slots_per_worker=2, run_policy=RunPolicy(clean_pod_policy="Running"), ssh_auth_mount_path="/home/mpiuser/.ssh",
                mpi_implementation="Intel",
                mpi_replica_specs={"Launcher": ReplicaSpec(replicas=1, restart_policy="Never", template=ONE_CONTAINER)})
    base.update(kw)
    return MPIJobSpec(**base)


# ------------------------------------------------------------------ defaults --
def test_defaults_base():
    j = set_defaults_mpijob(MPIJob())
    assert j.spec.slots_per_worker == 1
    assert j.spec.run_policy.clean_pod_policy == "None"
    assert j.spec.ssh_auth_mount_path == "/root/.ssh"
    assert j.spec.mpi_implementation == "OpenMPI"
    assert j.spec.launcher_creation_policy == "AtStartup"


def test_defaults_base_overridden():
    j = MPIJob(spec=MPIJobSpec(slots_per_worker=10, ssh_auth_mount_path="/home/mpiuser/.ssh", mpi_implementation="Intel",
                               launcher_creation_policy="WaitForWorkersReady",
                               run_policy=RunPolicy(clean_pod_policy="Running", ttl_seconds_after_finished=2,
                                                    active_deadline_seconds=3, backoff_limit=4)))
    want = j.deepcopy()
    assert set_defaults_mpijob(j) == want


def test_defaults_launcher_and_worker():
    j = MPIJob(spec=MPIJobSpec(mpi_replica_specs={"Launcher": ReplicaSpec(), "Worker": ReplicaSpec()}))
    set_defaults_mpijob(j)
    assert (j.spec.replica("Launcher").replicas, j.spec.replica("Launcher").restart_policy) == (1, "OnFailure")
    assert (j.spec.replica("Worker").replicas, j.spec.replica("Worker").restart_policy) == (0, "Never")


def test_defaults_replica_overrides_kept():
    j = MPIJob(spec=MPIJobSpec(mpi_replica_specs={"Launcher": ReplicaSpec(restart_policy="Never"),
                                                  "Worker": ReplicaSpec(replicas=3, restart_policy="OnFailure")}))
    set_defaults_mpijob(j)
    assert j.spec.replica("Launcher").restart_policy == "Never"
    assert (j.spec.replica("Worker").replicas, j.spec.replica("Worker").restart_policy) == (3, "OnFailure")


def test_defaults_idempotent_and_roundtrip():
    j = yaml_io.load_file("/root/repo/examples/pi/pi.yaml")[0]
    a = set_defaults_mpijob(j.deepcopy())
    assert set_defaults_mpijob(a.deepcopy()) == a
    assert MPIJob.from_dict(a.to_dict()) == a


# ---------------------------------------------------------------- validation --
def _errs(job):
    return [(e.type, e.field) for e in validate_mpijob(job)]


@pytest.mark.parametrize("impl", ["Intel", "MPICH", "OpenMPI"])
def test_valid(impl):
    assert _errs(MPIJob(metadata={"name": "foo"}, spec=_spec(mpi_implementation=impl))) == []
    specs = {"Launcher": ReplicaSpec(replicas=1, restart_policy="OnFailure", template=ONE_CONTAINER),
             "Worker": ReplicaSpec(replicas=3, restart_policy="Never", template=ONE_CONTAINER)}
    assert _errs(MPIJob(metadata={"name": "foo"}, spec=_spec(mpi_implementation=impl, mpi_replica_specs=specs))) == []


def test_empty_job():
    assert _errs(MPIJob()) == [
        (ERR_INVALID, "metadata.name"), (ERR_REQUIRED, "spec.mpiReplicaSpecs"), (ERR_REQUIRED, "spec.slotsPerWorker"),
        (ERR_REQUIRED, "spec.runPolicy.cleanPodPolicy"), (ERR_REQUIRED, "spec.sshAuthMountPath"),
        (ERR_NOT_SUPPORTED, "spec.mpiImplementation")]


def test_invalid_fields():
    specs = {"Launcher": ReplicaSpec(replicas=1, restart_policy="Never", template=ONE_CONTAINER),
             "Worker": ReplicaSpec(replicas=1000, restart_policy="Never", template=ONE_CONTAINER)}
    j = MPIJob(metadata={"name": "this-name-is-waaaaaaaay-too-long-for-a-worker-hostname"},
               spec=_spec(mpi_implementation="Unknown", mpi_replica_specs=specs,
                          run_policy=RunPolicy(clean_pod_policy="unknown", ttl_seconds_after_finished=-1,
                                               active_deadline_seconds=-1, backoff_limit=-1, managed_by="other.sigs.k8s.io/other")))
    assert _errs(j) == [
        (ERR_INVALID, "metadata.name"), (ERR_NOT_SUPPORTED, "spec.runPolicy.cleanPodPolicy"),
        (ERR_INVALID, "spec.runPolicy.ttlSecondsAfterFinished"), (ERR_INVALID, "spec.runPolicy.activeDeadlineSeconds"),
        (ERR_INVALID, "spec.runPolicy.backoffLimit"), (ERR_NOT_SUPPORTED, "spec.runPolicy.managedBy"),
        (ERR_NOT_SUPPORTED, "spec.mpiImplementation")]
    assert all(e.origin == "minimum" for e in validate_mpijob(j) if "Seconds" in e.field or "backoff" in e.field)


def test_empty_replica_specs():
    j = MPIJob(metadata={"name": "foo"}, spec=_spec(mpi_replica_specs={}, mpi_implementation="OpenMPI"))
    assert _errs(j) == [(ERR_REQUIRED, "spec.mpiReplicaSpecs[Launcher]")]


def test_missing_replica_spec_fields():
    j = MPIJob(metadata={"name": "foo"}, spec=_spec(mpi_replica_specs={"Launcher": ReplicaSpec(), "Worker": ReplicaSpec()}))
    assert _errs(j) == [
        (ERR_REQUIRED, "spec.mpiReplicaSpecs[Launcher].replicas"), (ERR_NOT_SUPPORTED, "spec.mpiReplicaSpecs[Launcher].restartPolicy"),
        (ERR_REQUIRED, "spec.mpiReplicaSpecs[Launcher].template.spec.containers"),
        (ERR_REQUIRED, "spec.mpiReplicaSpecs[Worker].replicas"), (ERR_NOT_SUPPORTED, "spec.mpiReplicaSpecs[Worker].restartPolicy"),
        (ERR_REQUIRED, "spec.mpiReplicaSpecs[Worker].template.spec.containers")]


def test_invalid_replica_fields():
    specs = {"Launcher": ReplicaSpec(replicas=2, restart_policy="Always", template=ONE_CONTAINER),
             "Worker": ReplicaSpec(replicas=0, restart_policy="Invalid", template=ONE_CONTAINER)}
    assert _errs(MPIJob(metadata={"name": "foo"}, spec=_spec(mpi_replica_specs=specs))) == [
        (ERR_NOT_SUPPORTED, "spec.mpiReplicaSpecs[Launcher].restartPolicy"), (ERR_INVALID, "spec.mpiReplicaSpecs[Launcher].replicas"),
        (ERR_NOT_SUPPORTED, "spec.mpiReplicaSpecs[Worker].restartPolicy"), (ERR_INVALID, "spec.mpiReplicaSpecs[Worker].replicas")]


def test_invalid_name():
    assert _errs(MPIJob(metadata={"name": "1-foo"}, spec=_spec())) == [(ERR_INVALID, "metadata.name")]
    msg = validate_mpijob(MPIJob(metadata={"name": "1-foo"}, spec=_spec())).to_aggregate()
    assert 'invalid DNS label "1-foo-worker-0"' in msg and "DNS-1035" in msg


def test_aggregate_format():
    agg = validate_mpijob(MPIJob(metadata={"name": "foo"}, spec=_spec(mpi_implementation="X", slots_per_worker=None))).to_aggregate()
    assert agg.startswith("[spec.slotsPerWorker: Required value: must have number of slots per worker, ")
    assert 'spec.mpiImplementation: Unsupported value: "X": supported values: "Intel", "MPICH", "OpenMPI"]' in agg


def test_reference_examples_load_and_validate():
    for f in ["pi/pi.yaml", "pi/pi-intel.yaml", "pi/pi-mpich.yaml", "horovod/tensorflow-mnist.yaml",
              "tensorflow-benchmarks/tensorflow-benchmarks.yaml"]:
        jobs = yaml_io.load_file("/root/reference/examples/v2beta1/" + f)
        assert len(jobs) == 1 and not validate_mpijob(set_defaults_mpijob(jobs[0]))


def test_scheduling_policy_roundtrip():
    sp = SchedulingPolicy(min_available=3, queue="q", min_resources={"nvidia.com/gpu": "4"}, priority_class="high", schedule_timeout_seconds=30)
    d = sp.to_dict()
    assert d == {"minAvailable": 3, "queue": "q", "minResources": {"nvidia.com/gpu": "4"}, "priorityClass": "high", "scheduleTimeoutSeconds": 30}
    assert SchedulingPolicy.from_dict(d) == sp
