"""SDK model classes with the reference SDK's surface.

Reference: sdk/python/v2beta1/mpijob/models/v2beta1_*.py (9 MPIJob models, e.g.
v2beta1_mpi_job_spec.py:35-53 ``openapi_types``, :55-82 ``attribute_map``,
:151 required-field ``ValueError``, :245-297 ``to_dict/to_str/__eq__``) plus the
handful of apimachinery meta models jobs actually use.  The reference generates
one file per class with openapi-generator; here one declarative table builds
the same classes (same constructor kwargs, same properties, same validation).
"""
from __future__ import annotations

import pprint
from typing import Any, Dict, Tuple

from .configuration import Configuration


class _ModelMeta(type):
    def __new__(mcs, name, bases, ns):
        types: Dict[str, str] = ns.get("openapi_types", {})
        required: Tuple[str, ...] = ns.get("required", ())
        for attr in types:
            ns[attr] = mcs._make_property(attr, attr in required)
        return super().__new__(mcs, name, bases, ns)

    @staticmethod
    def _make_property(attr: str, required: bool):
        private = "_" + attr

        def getter(self):
            return getattr(self, private)

        def setter(self, value):
            if self.local_vars_configuration.client_side_validation and required and value is None:
                raise ValueError(f"Invalid value for `{attr}`, must not be `None`")
            setattr(self, private, value)
        return property(getter, setter, doc=f"Gets/sets the {attr} of this model.")


class OpenApiModel(metaclass=_ModelMeta):
    openapi_types: Dict[str, str] = {}
    attribute_map: Dict[str, str] = {}
    required: Tuple[str, ...] = ()

    def __init__(self, local_vars_configuration=None, **kwargs):
        if local_vars_configuration is None:
            local_vars_configuration = Configuration.get_default_copy()
        self.local_vars_configuration = local_vars_configuration
        self.discriminator = None
        unknown = set(kwargs) - set(self.openapi_types)
        if unknown:
            raise TypeError(f"{type(self).__name__}() got unexpected keyword arguments {sorted(unknown)}")
        for attr in self.openapi_types:
            setattr(self, "_" + attr, None)
        for attr in self.openapi_types:
            v = kwargs.get(attr)
            if v is not None or attr in self.required:
                setattr(self, attr, v)

    def to_dict(self, serialize: bool = False):
        result = {}

        def convert(x):
            if hasattr(x, "to_dict"):
                try:
                    return x.to_dict(serialize) if isinstance(x, OpenApiModel) else x.to_dict()
                except TypeError:
                    return x.to_dict()
            if isinstance(x, list):
                return [convert(i) for i in x]
            if isinstance(x, dict):
                return {k: convert(v) for k, v in x.items()}
            return x
        for attr in self.openapi_types:
            value = getattr(self, attr)
            key = self.attribute_map.get(attr, attr) if serialize else attr
            result[key] = convert(value)
        return result

    def to_str(self):
        return pprint.pformat(self.to_dict())

    def __repr__(self):
        return self.to_str()

    def __eq__(self, other):
        return isinstance(other, type(self)) and self.to_dict() == other.to_dict()

    def __ne__(self, other):
        return not self.__eq__(other)


class V2beta1JobCondition(OpenApiModel):
    openapi_types = {"last_transition_time": "datetime", "last_update_time": "datetime", "message": "str",
                     "reason": "str", "status": "str", "type": "str"}
    attribute_map = {"last_transition_time": "lastTransitionTime", "last_update_time": "lastUpdateTime",
                     "message": "message", "reason": "reason", "status": "status", "type": "type"}
    required = ("status", "type")


class V2beta1ReplicaStatus(OpenApiModel):
    openapi_types = {"active": "int", "failed": "int", "label_selector": "V1LabelSelector", "selector": "str", "succeeded": "int"}
    attribute_map = {"active": "active", "failed": "failed", "label_selector": "labelSelector", "selector": "selector",
                     "succeeded": "succeeded"}


class V2beta1JobStatus(OpenApiModel):
    openapi_types = {"completion_time": "datetime", "conditions": "list[V2beta1JobCondition]", "last_reconcile_time": "datetime",
                     "replica_statuses": "dict(str, V2beta1ReplicaStatus)", "start_time": "datetime"}
    attribute_map = {"completion_time": "completionTime", "conditions": "conditions", "last_reconcile_time": "lastReconcileTime",
                     "replica_statuses": "replicaStatuses", "start_time": "startTime"}


class V2beta1SchedulingPolicy(OpenApiModel):
    openapi_types = {"min_available": "int", "min_resources": "dict(str, object)", "priority_class": "str", "queue": "str",
                     "schedule_timeout_seconds": "int"}
    attribute_map = {"min_available": "minAvailable", "min_resources": "minResources", "priority_class": "priorityClass",
                     "queue": "queue", "schedule_timeout_seconds": "scheduleTimeoutSeconds"}


class V2beta1RunPolicy(OpenApiModel):
    openapi_types = {"active_deadline_seconds": "int", "backoff_limit": "int", "clean_pod_policy": "str", "managed_by": "str",
                     "scheduling_policy": "V2beta1SchedulingPolicy", "suspend": "bool", "ttl_seconds_after_finished": "int"}
    attribute_map = {"active_deadline_seconds": "activeDeadlineSeconds", "backoff_limit": "backoffLimit",
                     "clean_pod_policy": "cleanPodPolicy", "managed_by": "managedBy", "scheduling_policy": "schedulingPolicy",
                     "suspend": "suspend", "ttl_seconds_after_finished": "ttlSecondsAfterFinished"}


class V2beta1ReplicaSpec(OpenApiModel):
    openapi_types = {"replicas": "int", "restart_policy": "str", "template": "V1PodTemplateSpec"}
    attribute_map = {"replicas": "replicas", "restart_policy": "restartPolicy", "template": "template"}


class V2beta1MPIJobSpec(OpenApiModel):
    openapi_types = {"launcher_creation_policy": "str", "mpi_implementation": "str",
                     "mpi_replica_specs": "dict(str, V2beta1ReplicaSpec)", "run_launcher_as_worker": "bool",
                     "run_policy": "V2beta1RunPolicy", "slots_per_worker": "int", "ssh_auth_mount_path": "str"}
    attribute_map = {"launcher_creation_policy": "launcherCreationPolicy", "mpi_implementation": "mpiImplementation",
                     "mpi_replica_specs": "mpiReplicaSpecs", "run_launcher_as_worker": "runLauncherAsWorker",
                     "run_policy": "runPolicy", "slots_per_worker": "slotsPerWorker", "ssh_auth_mount_path": "sshAuthMountPath"}
    required = ("mpi_replica_specs",)


class V2beta1MPIJob(OpenApiModel):
    openapi_types = {"api_version": "str", "kind": "str", "metadata": "V1ObjectMeta", "spec": "V2beta1MPIJobSpec",
                     "status": "V2beta1JobStatus"}
    attribute_map = {"api_version": "apiVersion", "kind": "kind", "metadata": "metadata", "spec": "spec", "status": "status"}


class V2beta1MPIJobList(OpenApiModel):
    openapi_types = {"api_version": "str", "items": "list[V2beta1MPIJob]", "kind": "str", "metadata": "V1ListMeta"}
    attribute_map = {"api_version": "apiVersion", "items": "items", "kind": "kind", "metadata": "metadata"}
    required = ("items", "metadata")


# ---- apimachinery / core shims (the reference imports these from `kubernetes`) ----
class V1ObjectMeta(OpenApiModel):
    openapi_types = {"annotations": "dict(str, str)", "creation_timestamp": "datetime", "deletion_timestamp": "datetime",
                     "finalizers": "list[str]", "generate_name": "str", "generation": "int", "labels": "dict(str, str)",
                     "name": "str", "namespace": "str", "owner_references": "list[V1OwnerReference]",
                     "resource_version": "str", "uid": "str"}
    attribute_map = {"annotations": "annotations", "creation_timestamp": "creationTimestamp",
                     "deletion_timestamp": "deletionTimestamp", "finalizers": "finalizers", "generate_name": "generateName",
                     "generation": "generation", "labels": "labels", "name": "name", "namespace": "namespace",
                     "owner_references": "ownerReferences", "resource_version": "resourceVersion", "uid": "uid"}


class V1ListMeta(OpenApiModel):
    openapi_types = {"_continue": "str", "remaining_item_count": "int", "resource_version": "str", "self_link": "str"}
    attribute_map = {"_continue": "continue", "remaining_item_count": "remainingItemCount",
                     "resource_version": "resourceVersion", "self_link": "selfLink"}


class V1OwnerReference(OpenApiModel):
    openapi_types = {"api_version": "str", "block_owner_deletion": "bool", "controller": "bool", "kind": "str", "name": "str", "uid": "str"}
    attribute_map = {"api_version": "apiVersion", "block_owner_deletion": "blockOwnerDeletion", "controller": "controller",
                     "kind": "kind", "name": "name", "uid": "uid"}
    required = ("api_version", "kind", "name", "uid")


class V1LabelSelectorRequirement(OpenApiModel):
    openapi_types = {"key": "str", "operator": "str", "values": "list[str]"}
    attribute_map = {"key": "key", "operator": "operator", "values": "values"}
    required = ("key", "operator")


class V1LabelSelector(OpenApiModel):
    openapi_types = {"match_expressions": "list[V1LabelSelectorRequirement]", "match_labels": "dict(str, str)"}
    attribute_map = {"match_expressions": "matchExpressions", "match_labels": "matchLabels"}


class V1Container(OpenApiModel):
    openapi_types = {"args": "list[str]", "command": "list[str]", "env": "list[object]", "image": "str", "name": "str",
                     "resources": "object", "security_context": "object", "volume_mounts": "list[object]", "working_dir": "str",
                     "readiness_probe": "object", "image_pull_policy": "str"}
    attribute_map = {"args": "args", "command": "command", "env": "env", "image": "image", "name": "name",
                     "resources": "resources", "security_context": "securityContext", "volume_mounts": "volumeMounts",
                     "working_dir": "workingDir", "readiness_probe": "readinessProbe", "image_pull_policy": "imagePullPolicy"}
    required = ("name",)


class V1PodSpec(OpenApiModel):
    openapi_types = {"containers": "list[V1Container]", "restart_policy": "str", "host_network": "bool", "volumes": "list[object]",
                     "node_selector": "dict(str, str)", "tolerations": "list[object]", "priority_class_name": "str",
                     "scheduler_name": "str", "scheduling_gates": "list[object]", "dns_policy": "str", "hostname": "str",
                     "subdomain": "str"}
    attribute_map = {"containers": "containers", "restart_policy": "restartPolicy", "host_network": "hostNetwork",
                     "volumes": "volumes", "node_selector": "nodeSelector", "tolerations": "tolerations",
                     "priority_class_name": "priorityClassName", "scheduler_name": "schedulerName",
                     "scheduling_gates": "schedulingGates", "dns_policy": "dnsPolicy", "hostname": "hostname", "subdomain": "subdomain"}
    required = ("containers",)


class V1PodTemplateSpec(OpenApiModel):
    openapi_types = {"metadata": "V1ObjectMeta", "spec": "V1PodSpec"}
    attribute_map = {"metadata": "metadata", "spec": "spec"}


MODEL_CLASSES = {c.__name__: c for c in (
    V2beta1JobCondition, V2beta1ReplicaStatus, V2beta1JobStatus, V2beta1SchedulingPolicy, V2beta1RunPolicy,
    V2beta1ReplicaSpec, V2beta1MPIJobSpec, V2beta1MPIJob, V2beta1MPIJobList, V1ObjectMeta, V1ListMeta, V1OwnerReference,
    V1LabelSelectorRequirement, V1LabelSelector, V1Container, V1PodSpec, V1PodTemplateSpec)}

# long apimachinery names used by the reference's generated SDK (sdk/python/v2beta1/mpijob/models/io_k8s_*.py)
for _short in ("V1ObjectMeta", "V1ListMeta", "V1OwnerReference", "V1LabelSelector", "V1LabelSelectorRequirement"):
    MODEL_CLASSES["IoK8sApimachineryPkgApisMeta" + _short] = MODEL_CLASSES[_short]
    globals()["IoK8sApimachineryPkgApisMeta" + _short] = MODEL_CLASSES[_short]
