"""Per-model construction tests, the role of the reference's 109 generated stubs
(sdk/python/v2beta1/test/test_v2beta1_*.py: make_instance(include_optional) for every model)."""
import pytest

import mpijob

OPTIONAL = {
    "V2beta1JobCondition": dict(type="Created", status="True", reason="r", message="m", last_update_time="2020-01-01T00:00:00Z",
                                last_transition_time="2020-01-01T00:00:00Z"),
    "V2beta1ReplicaStatus": dict(active=1, succeeded=2, failed=3, selector="a=b", label_selector=mpijob.V1LabelSelector(match_labels={"a": "b"})),
    "V2beta1JobStatus": dict(conditions=[mpijob.V2beta1JobCondition(type="Created", status="True")], start_time="t", completion_time="t",
                             last_reconcile_time="t", replica_statuses={"Worker": mpijob.V2beta1ReplicaStatus(active=1)}),
    "V2beta1SchedulingPolicy": dict(min_available=3, queue="q", priority_class="p", schedule_timeout_seconds=5, min_resources={"cpu": "1"}),
    "V2beta1RunPolicy": dict(clean_pod_policy="Running", ttl_seconds_after_finished=1, active_deadline_seconds=2, backoff_limit=3, suspend=False,
                             managed_by="kubeflow.org/mpi-operator", scheduling_policy=mpijob.V2beta1SchedulingPolicy(min_available=1)),
    "V2beta1ReplicaSpec": dict(replicas=2, restart_policy="Never", template=mpijob.V1PodTemplateSpec(spec=mpijob.V1PodSpec(containers=[mpijob.V1Container(name="c")]))),
    "V2beta1MPIJobSpec": dict(mpi_replica_specs={"Launcher": mpijob.V2beta1ReplicaSpec(replicas=1)}, slots_per_worker=1, run_launcher_as_worker=False,
                              ssh_auth_mount_path="/root/.ssh", launcher_creation_policy="AtStartup", mpi_implementation="OpenMPI",
                              run_policy=mpijob.V2beta1RunPolicy()),
    "V2beta1MPIJob": dict(api_version="kubeflow.org/v2beta1", kind="MPIJob", metadata=mpijob.V1ObjectMeta(name="n"),
                          spec=mpijob.V2beta1MPIJobSpec(mpi_replica_specs={}), status=mpijob.V2beta1JobStatus()),
    "V2beta1MPIJobList": dict(api_version="kubeflow.org/v2beta1", kind="MPIJobList", metadata=mpijob.V1ListMeta(), items=[]),
}
REQUIRED_ONLY = {
    "V2beta1JobCondition": dict(type="Created", status="True"),
    "V2beta1MPIJobSpec": dict(mpi_replica_specs={}),
    "V2beta1MPIJobList": dict(metadata=mpijob.V1ListMeta(), items=[]),
}


@pytest.mark.parametrize("name", sorted(OPTIONAL))
@pytest.mark.parametrize("include_optional", [False, True])
def test_make_instance(name, include_optional):
    cls = getattr(mpijob, name)
    kwargs = OPTIONAL[name] if include_optional else REQUIRED_ONLY.get(name, {})
    inst = cls(**kwargs)
    assert set(inst.to_dict()) == set(cls.openapi_types)
    assert set(cls.attribute_map) == set(cls.openapi_types)
    client = mpijob.ApiClient()
    body = client.sanitize_for_serialization(inst)
    assert all(k in cls.attribute_map.values() for k in body)
    again = client.deserialize(body, name)
    assert client.sanitize_for_serialization(again) == body
    assert inst == cls(**kwargs) and (inst != cls(**OPTIONAL[name])) == (not include_optional and kwargs != OPTIONAL[name])
    with pytest.raises(TypeError):
        cls(not_a_field=1)
