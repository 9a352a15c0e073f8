"""High-level MPIJob client: what ``kubernetes.client.CustomObjectsApi`` is to the
reference's SDK example (sdk/python/v2beta1/tensorflow-mnist.py:91-128:
``create_namespaced_custom_object(group="kubeflow.org", version="v2beta1",
plural="mpijobs", body=job)``), pointed at the single-box daemon."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Tuple, Union

from .api_client import ApiClient
from .configuration import Configuration
from .exceptions import ApiException, NotFoundException
from .models import V2beta1MPIJob

GROUP, VERSION, PLURAL = "kubeflow.org", "v2beta1", "mpijobs"
_PATHS = {
    "mpijobs": "/apis/kubeflow.org/v2beta1", "pods": "/api/v1", "services": "/api/v1", "configmaps": "/api/v1",
    "secrets": "/api/v1", "events": "/api/v1", "jobs": "/apis/batch/v1", "leases": "/apis/coordination.k8s.io/v1",
    "podgroups": "/apis/scheduling.volcano.sh/v1beta1",
}


class MPIJobClient:
    def __init__(self, server: Optional[str] = None, api_client: Optional[ApiClient] = None):
        if api_client is None:
            cfg = Configuration(host=("http://" + server) if server and not server.startswith("http") else server)
            api_client = ApiClient(cfg)
        self.api = api_client

    # ---------------------------------------------------------- plumbing --
    def _path(self, resource: str, namespace: Optional[str], name: str = "", sub: str = "") -> str:
        p = _PATHS[resource]
        if namespace:
            p += f"/namespaces/{namespace}"
        p += f"/{resource}"
        if name:
            p += f"/{name}"
        if sub:
            p += f"/{sub}"
        return p

    def _body(self, job: Union[V2beta1MPIJob, Dict[str, Any]]) -> Dict[str, Any]:
        return job if isinstance(job, dict) else self.api.sanitize_for_serialization(job)

    def raw_get(self, path: str):
        return self.api.call_api(path, "GET")

    # --------------------------------------------- CustomObjectsApi parity --
    def create_namespaced_custom_object(self, group, version, namespace, plural, body, **kw):
        assert (group, version, plural) == (GROUP, VERSION, PLURAL), "only kubeflow.org/v2beta1 mpijobs are served"
        return self.api.call_api(self._path("mpijobs", namespace), "POST", body=self._body(body))

    def get_namespaced_custom_object(self, group, version, namespace, plural, name, **kw):
        return self.get(name, namespace)

    def list_namespaced_custom_object(self, group, version, namespace, plural, **kw):
        return self.api.call_api(self._path("mpijobs", namespace), "GET")

    def delete_namespaced_custom_object(self, group, version, namespace, plural, name, **kw):
        return self.delete(name, namespace)

    def patch_namespaced_custom_object(self, group, version, namespace, plural, name, body, **kw):
        return self.patch(name, body, namespace)

    # ------------------------------------------------------------- verbs --
    def create(self, job, namespace: str = "default") -> Dict[str, Any]:
        return self.api.call_api(self._path("mpijobs", namespace), "POST", body=self._body(job))

    def get(self, name: str, namespace: str = "default") -> Dict[str, Any]:
        return self.api.call_api(self._path("mpijobs", namespace, name), "GET")

    def get_model(self, name: str, namespace: str = "default") -> V2beta1MPIJob:
        return self.api.call_api(self._path("mpijobs", namespace, name), "GET", response_type="V2beta1MPIJob")

    def list(self, namespace: Optional[str] = "default") -> List[Dict[str, Any]]:
        return self.api.call_api(self._path("mpijobs", namespace), "GET")["items"]

    def delete(self, name: str, namespace: str = "default"):
        return self.api.call_api(self._path("mpijobs", namespace, name), "DELETE")

    def patch(self, name: str, patch: Dict[str, Any], namespace: str = "default"):
        return self.api.call_api(self._path("mpijobs", namespace, name), "PATCH", body=patch)

    def apply(self, job, namespace: Optional[str] = None) -> Tuple[Dict[str, Any], str]:
        body = self._body(job)
        ns = namespace or body.get("metadata", {}).get("namespace", "default")
        name = body["metadata"]["name"]
        try:
            cur = self.get(name, ns)
        except NotFoundException:
            return self.create(body, ns), "created"
        if cur.get("spec") == body.get("spec"):
            return cur, "unchanged"
        body = dict(body)
        body["metadata"] = dict(body.get("metadata", {}))
        body["metadata"]["resourceVersion"] = cur["metadata"]["resourceVersion"]
        return self.api.call_api(self._path("mpijobs", ns, name), "PUT", body=body), "configured"

    def scale(self, name: str, replicas: int, namespace: str = "default"):
        return self.patch(name, {"spec": {"mpiReplicaSpecs": {"Worker": {"replicas": replicas}}}}, namespace)

    def suspend(self, name: str, namespace: str = "default"):
        return self.patch(name, {"spec": {"runPolicy": {"suspend": True}}}, namespace)

    def resume(self, name: str, namespace: str = "default"):
        return self.patch(name, {"spec": {"runPolicy": {"suspend": False}}}, namespace)

    def wait_for_condition(self, name: str, condition: str = "Succeeded", namespace: str = "default", timeout: float = 300,
                           poll: float = 0.1) -> Dict[str, Any]:
        deadline = time.time() + timeout
        want = {condition} | ({"Failed"} if condition == "Succeeded" else set())
        while True:
            j = self.get(name, namespace)
            for c in j.get("status", {}).get("conditions", []) or []:
                if c["type"] in want and c["status"] == "True":
                    if c["type"] != condition:
                        raise RuntimeError(f"MPIJob {namespace}/{name} reached {c['type']}: {c.get('reason')}: {c.get('message')}")
                    return j
            if time.time() > deadline:
                raise TimeoutError(f"timed out waiting for condition {condition} on mpijob/{name}")
            time.sleep(poll)

    def logs(self, name: str, namespace: str = "default", worker: Optional[int] = None, pod: Optional[str] = None,
             tail: Optional[int] = None) -> str:
        if pod is None:
            pods = self.list_resource("pods", namespace)
            role = "worker" if worker is not None else "launcher"
            cands = [p for p in pods if (p["metadata"].get("labels") or {}).get("training.kubeflow.org/job-name") == name
                     and (p["metadata"].get("labels") or {}).get("training.kubeflow.org/job-role") == role]
            if worker is not None:
                cands = [p for p in cands if p["metadata"]["name"].endswith(f"-worker-{worker}")]
            if not cands:
                raise ApiException(status=404, reason=f"no {role} pod found for mpijob {name}")
            pod = sorted(cands, key=lambda p: p["metadata"].get("creationTimestamp", ""))[-1]["metadata"]["name"]
        path = self._path("pods", namespace, pod, "log")
        if tail is not None:
            path += f"?tailLines={int(tail)}"
        data = self.api.call_api(path, "GET", response_type="raw")
        return data.decode(errors="replace")

    def follow_pod_log(self, pod: str, namespace: str = "default", timeout: float = 3600.0):
        """Generator over `pods/<pod>/log?follow=true`: text pieces as the container writes them, until the pod finishes."""
        import http.client
        import urllib.parse
        u = urllib.parse.urlparse(self.api.configuration.host)
        conn = http.client.HTTPConnection(u.hostname, u.port or 80, timeout=timeout + 10)
        try:
            conn.request("GET", self._path("pods", namespace, pod, "log") + f"?follow=true&timeoutSeconds={float(timeout)!r}",
                         headers=self.api.configuration.auth_headers())
            resp = conn.getresponse()
            if resp.status != 200:
                raise ApiException(status=resp.status, reason=resp.reason, body=resp.read())
            while True:
                chunk = resp.read1(65536)
                if not chunk:
                    return
                yield chunk.decode(errors="replace")
        finally:
            conn.close()

    def watch(self, resource: str = "mpijobs", namespace: Optional[str] = "default", timeout: float = 300.0,
              label_selector: Optional[str] = None, name: Optional[str] = None):
        """Generator over the server's `?watch=true` stream: ``{"type": "ADDED" | "MODIFIED" | "DELETED", "object": {...}}``,
        existing objects first, until ``timeout`` seconds have passed (then the generator ends)."""
        import http.client
        import json
        import urllib.parse
        u = urllib.parse.urlparse(self.api.configuration.host)
        q = {"watch": "true", "timeoutSeconds": repr(float(timeout))}
        if label_selector:
            q["labelSelector"] = label_selector
        if name:
            q["fieldSelector"] = f"metadata.name={name}"
        conn = http.client.HTTPConnection(u.hostname, u.port or 80, timeout=timeout + 10)
        try:
            conn.request("GET", self._path(resource, namespace) + "?" + urllib.parse.urlencode(q), headers=self.api.configuration.auth_headers())
            resp = conn.getresponse()
            if resp.status != 200:
                raise ApiException(status=resp.status, reason=resp.reason, body=resp.read())
            buf = b""
            while True:
                chunk = resp.read1(65536)
                if not chunk:
                    return
                buf += chunk
                while b"\n" in buf:
                    line, buf = buf.split(b"\n", 1)
                    if line.strip():
                        yield json.loads(line)
        finally:
            conn.close()

    # ------------------------------------------------------- generic access --
    def list_resource(self, resource: str, namespace: Optional[str] = "default") -> List[Dict[str, Any]]:
        if resource == "nodes":   # cluster-scoped, synthesised by the daemon (the box as a v1.Node)
            return self.api.call_api("/api/v1/nodes", "GET")["items"]
        return self.api.call_api(self._path(resource, namespace), "GET")["items"]

    def get_resource(self, resource: str, namespace: str, name: str) -> Dict[str, Any]:
        if resource == "nodes":
            return self.api.call_api(f"/api/v1/nodes/{name}", "GET")
        return self.api.call_api(self._path(resource, namespace, name), "GET")

    def delete_resource(self, resource: str, namespace: str, name: str):
        return self.api.call_api(self._path(resource, namespace, name), "DELETE")
