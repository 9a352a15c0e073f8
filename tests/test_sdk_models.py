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


# ---- the generic apimachinery models (reference: 100 generated files + their 100 test stubs), built from sdk/meta_models.py ----
import importlib  # noqa: E402
import re  # noqa: E402

from mpi_operator_b200.sdk.meta_models import META_MODELS, snake  # noqa: E402

ALL_MODELS = mpijob.models.MODEL_CLASSES


def _sample(typ: str, depth: int = 0):
    """A value of the openapi type `typ` (nested models with their required fields only below the first level)."""
    if typ == "str":
        return "s"
    if typ == "int":
        return 3
    if typ == "bool":
        return True
    if typ == "datetime":
        return "2020-01-01T00:00:00Z"
    if typ == "object":
        return {"k": "v"}
    m = re.match(r"list\[(.*)\]$", typ)
    if m:
        return [_sample(m.group(1), depth + 1)]
    m = re.match(r"dict\(str, (.*)\)$", typ)
    if m:
        return {"a": _sample(m.group(1), depth + 1)}
    cls = ALL_MODELS[typ]
    fields = cls.openapi_types if depth == 0 else {a: cls.openapi_types[a] for a in cls.required}
    return cls(**{a: _sample(t, depth + 1) for a, t in fields.items()})


def test_generated_package_inventory():
    """44 meta/v1 types under two names, 3 runtime/version types under two prefixes, 6 pre-v2beta1 leftovers, 9 MPIJob models:
    the 109 model modules of sdk/python/v2beta1/mpijob/models/."""
    meta = [n for n in ALL_MODELS if n.startswith("IoK8sApimachineryPkgApisMetaV1")]
    assert len(meta) == 44 and all("V1" + n[len("IoK8sApimachineryPkgApisMetaV1"):] in ALL_MODELS for n in meta)
    for base in ("RuntimeTypeMeta", "RuntimeUnknown", "VersionInfo"):
        assert "IoK8sApimachineryPkg" + base in ALL_MODELS and "K8sIoApimachineryPkg" + base in ALL_MODELS
    for legacy in ("V1JobCondition", "V1JobStatus", "V1ReplicaSpec", "V1ReplicaStatus", "V1RunPolicy", "V1SchedulingPolicy"):
        assert legacy in ALL_MODELS
    assert len([n for n in ALL_MODELS if n.startswith("V2beta1")]) == 9
    assert len(ALL_MODELS) >= 109
    # the hand-listed classes and the table agree on identity: one class per name
    assert ALL_MODELS["IoK8sApimachineryPkgApisMetaV1ObjectMeta"] is mpijob.V1ObjectMeta is META_MODELS["V1ObjectMeta"]
    assert {"managed_fields", "self_link", "deletion_grace_period_seconds"} <= set(mpijob.V1ObjectMeta.openapi_types)


@pytest.mark.parametrize("name,want", [("serverAddressByClientCIDRs", "server_address_by_client_cidrs"), ("clientCIDR", "client_cidr"),
                                       ("continue", "_continue"), ("fieldsV1", "fields_v1"), ("ContentEncoding", "content_encoding"),
                                       ("V1APIGroupList", "v1_api_group_list"), ("V2beta1MPIJobSpec", "v2beta1_mpi_job_spec"),
                                       ("IoK8sApimachineryPkgApisMetaV1WatchEvent", "io_k8s_apimachinery_pkg_apis_meta_v1_watch_event"),
                                       ("ignoreStoreReadErrorWithClusterBreakingPotential", "ignore_store_read_error_with_cluster_breaking_potential")])
def test_snake_names(name, want):
    assert snake(name) == want


@pytest.mark.parametrize("name", sorted(ALL_MODELS))
@pytest.mark.parametrize("include_optional", [False, True])
def test_every_model_make_instance(name, include_optional):
    cls = ALL_MODELS[name]
    assert set(cls.attribute_map) == set(cls.openapi_types) and set(cls.required) <= set(cls.openapi_types)
    fields = cls.openapi_types if include_optional else {a: cls.openapi_types[a] for a in cls.required}
    kwargs = {a: _sample(t, 1) for a, t in fields.items()}
    inst = cls(**kwargs)
    assert set(inst.to_dict()) == set(cls.openapi_types)
    assert inst == cls(**kwargs) and not (inst != cls(**kwargs))
    assert isinstance(inst.to_str(), str) and repr(inst) == inst.to_str()
    client = mpijob.ApiClient()
    body = client.sanitize_for_serialization(inst)
    assert set(body) == {cls.attribute_map[a] for a in kwargs}                  # JSON names, unset fields dropped
    again = client.deserialize(body, cls.__name__)
    assert type(again) is cls and client.sanitize_for_serialization(again) == body
    for req in cls.required:                                                     # required fields refuse None, at construction and on assignment
        with pytest.raises(ValueError):
            cls(**{k: v for k, v in kwargs.items() if k != req})
        with pytest.raises(ValueError):
            setattr(inst, req, None)
    with pytest.raises(TypeError):
        cls(not_a_field=1)
    # the generated package's import path: one module per model
    mod = importlib.import_module(f"mpijob.models.{snake(name)}")
    assert getattr(mod, name) is cls and getattr(mpijob.models, name) is cls


def test_status_and_delete_options_round_trip():
    """The two meta models a client meets in practice: the Status body of a failed call and DeleteOptions of a delete."""
    client = mpijob.ApiClient()
    st = client.deserialize({"kind": "Status", "apiVersion": "v1", "status": "Failure", "reason": "NotFound", "code": 404,
                             "message": 'mpijobs.kubeflow.org "x" not found', "metadata": {"continue": "c"},
                             "details": {"name": "x", "group": "kubeflow.org", "kind": "mpijobs", "causes": [{"reason": "r", "field": "f"}]}}, "V1Status")
    assert isinstance(st, mpijob.V1Status) and isinstance(st.details, mpijob.V1StatusDetails)
    assert isinstance(st.details.causes[0], mpijob.V1StatusCause) and st.details.causes[0].field == "f"
    assert st.metadata._continue == "c" and st.code == 404
    do = mpijob.V1DeleteOptions(propagation_policy="Foreground", grace_period_seconds=0,
                                preconditions=mpijob.V1Preconditions(uid="u"), dry_run=["All"])
    assert client.sanitize_for_serialization(do) == {"propagationPolicy": "Foreground", "gracePeriodSeconds": 0,
                                                     "preconditions": {"uid": "u"}, "dryRun": ["All"]}
