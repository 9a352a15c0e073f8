"""The part of ``kubernetes.client`` MPIJob scripts use (see the package docstring)."""
from mpi_operator_b200.sdk import ApiClient, Configuration  # noqa: F401
from mpi_operator_b200.sdk.client import MPIJobClient
from mpi_operator_b200.sdk.exceptions import ApiException  # noqa: F401
from mpi_operator_b200.sdk.models import (OpenApiModel, V1Container, V1LabelSelector, V1LabelSelectorRequirement, V1ListMeta,  # noqa: F401
                                          V1ObjectMeta, V1OwnerReference, V1PodSpec, V1PodTemplateSpec)

from . import rest  # noqa: F401
from .. import config as _config


class V1ResourceRequirements(OpenApiModel):
    openapi_types = {"limits": "dict(str, str)", "requests": "dict(str, str)"}
    attribute_map = {"limits": "limits", "requests": "requests"}


class V1EnvVar(OpenApiModel):
    openapi_types = {"name": "str", "value": "str", "value_from": "object"}
    attribute_map = {"name": "name", "value": "value", "value_from": "valueFrom"}
    required = ("name",)


class V1VolumeMount(OpenApiModel):
    openapi_types = {"mount_path": "str", "name": "str", "read_only": "bool", "sub_path": "str"}
    attribute_map = {"mount_path": "mountPath", "name": "name", "read_only": "readOnly", "sub_path": "subPath"}
    required = ("mount_path", "name")


class V1DeleteOptions(OpenApiModel):
    openapi_types = {"propagation_policy": "str", "grace_period_seconds": "int"}
    attribute_map = {"propagation_policy": "propagationPolicy", "grace_period_seconds": "gracePeriodSeconds"}


def _client(api_client=None) -> MPIJobClient:
    if api_client is not None and hasattr(api_client, "call_api"):
        return MPIJobClient(api_client=api_client)
    return MPIJobClient(_config.current_host())


class CustomObjectsApi:
    """``kubeflow.org/v2beta1`` ``mpijobs`` only (anything else raises 404 like an API server without that CRD)."""

    def __init__(self, api_client=None):
        self._c = _client(api_client)

    @staticmethod
    def _check(group, version, plural):
        if (group, version, plural) != ("kubeflow.org", "v2beta1", "mpijobs"):
            raise ApiException(status=404, reason=f"the server could not find the requested resource ({group}/{version} {plural})")

    def create_namespaced_custom_object(self, group, version, namespace, plural, body, **kw):
        self._check(group, version, plural)
        return self._c.create_namespaced_custom_object(group, version, namespace, plural, body)

    def get_namespaced_custom_object(self, group, version, namespace, plural, name, **kw):
        self._check(group, version, plural)
        return self._c.get(name, namespace)

    def get_namespaced_custom_object_status(self, group, version, namespace, plural, name, **kw):
        return self.get_namespaced_custom_object(group, version, namespace, plural, name)

    def list_namespaced_custom_object(self, group, version, namespace, plural, **kw):
        self._check(group, version, plural)
        return self._c.list_namespaced_custom_object(group, version, namespace, plural)

    def list_cluster_custom_object(self, group, version, plural, **kw):
        self._check(group, version, plural)
        return {"apiVersion": "kubeflow.org/v2beta1", "kind": "MPIJobList", "items": self._c.list(None)}

    def delete_namespaced_custom_object(self, group, version, namespace, plural, name, **kw):
        self._check(group, version, plural)
        return self._c.delete(name, namespace)

    def patch_namespaced_custom_object(self, group, version, namespace, plural, name, body, **kw):
        self._check(group, version, plural)
        return self._c.patch(name, body if isinstance(body, dict) else self._c._body(body), namespace)

    def replace_namespaced_custom_object(self, group, version, namespace, plural, name, body, **kw):
        self._check(group, version, plural)
        return self._c.api.call_api(self._c._path("mpijobs", namespace, name), "PUT", body=self._c._body(body))


class _List:
    def __init__(self, items):
        self.items = items


class CoreV1Api:
    """Read-only view of what the node agent runs: pods (dicts wrapped for attribute access), their logs, events."""

    def __init__(self, api_client=None):
        self._c = _client(api_client)

    def list_namespaced_pod(self, namespace, label_selector=None, **kw):
        return _List([_Obj(p) for p in self._c.list_resource("pods", namespace) if _matches(p, label_selector)])

    def read_namespaced_pod(self, name, namespace, **kw):
        return _Obj(self._c.get_resource("pods", namespace, name))

    def read_namespaced_pod_log(self, name, namespace, tail_lines=None, follow=False, _preload_content=True, **kw):
        """``follow=True, _preload_content=False`` returns an iterator over the log as it grows (until the pod finishes)."""
        if follow:
            stream = self._c.follow_pod_log(name, namespace, timeout=float(kw.get("_request_timeout") or 3600.0))
            return stream if not _preload_content else "".join(stream)
        return self._c.logs("", namespace, pod=name, tail=tail_lines)

    def list_node(self, **kw):
        """The single box as a one-element NodeList (capacity / allocatable nvidia.com/gpu, cordoned GPUs as taints)."""
        return _List([_Obj(n) for n in self._c.list_resource("nodes", None)])

    def read_node(self, name, **kw):
        return _Obj(self._c.get_resource("nodes", None, name))

    def list_namespaced_event(self, namespace, **kw):
        return _List([_Obj(e) for e in self._c.list_resource("events", namespace)])


def _matches(obj, selector) -> bool:
    if not selector:
        return True
    labels = obj.get("metadata", {}).get("labels") or {}
    for term in selector.split(","):
        k, _, v = term.partition("=")
        if labels.get(k.strip()) != v.strip().lstrip("="):
            return False
    return True


class _Obj(dict):
    """dict with snake_case attribute access (``pod.metadata.name``, ``pod.status.phase``, ``pod.status.container_statuses``)."""

    def __getattr__(self, name):
        parts = name.split("_")
        camel = parts[0] + "".join(p.title() for p in parts[1:])
        for k in (name, camel):
            if k in self:
                v = self[k]
                if isinstance(v, dict):
                    return _Obj(v)
                if isinstance(v, list):
                    return [_Obj(i) if isinstance(i, dict) else i for i in v]
                return v
        return None
