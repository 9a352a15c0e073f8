"""Structural (type) validation of MPIJob objects against the CRD schema - what kube-apiserver does with the
``openAPIV3Schema`` of manifests/base/kubeflow.org_mpijobs.yaml before an object ever reaches the controller
(the reference gets it for free from the apiserver; its own validation, pkg/apis/kubeflow/validation/validation.go, assumes
well-typed input just like ``api/validation.py`` here does).

Only TYPES are checked here ("spec.slotsPerWorker in body must be of type integer"), plus enum membership; required fields and
semantic rules stay where the reference has them (admission for ``spec.mpiReplicaSpecs``, the controller's validation for the
rest). The pod templates are ``x-kubernetes-preserve-unknown-fields`` in the CRD, so the schema says nothing about them; the
fields the builders and the node agent read are checked with the core/v1 types they have there.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Any, Dict, List

_JSON_TYPES = {"object": dict, "array": list, "string": str, "boolean": bool}


def _type_name(v: Any) -> str:
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "boolean"
    if isinstance(v, int):
        return "integer"
    if isinstance(v, float):
        return "number"
    if isinstance(v, str):
        return "string"
    if isinstance(v, list):
        return "array"
    if isinstance(v, dict):
        return "object"
    return type(v).__name__


def _is(v: Any, want: str) -> bool:
    if want == "integer":
        return isinstance(v, int) and not isinstance(v, bool)
    if want == "number":
        return isinstance(v, (int, float)) and not isinstance(v, bool)
    return isinstance(v, _JSON_TYPES[want]) if want in _JSON_TYPES else True


def _wrong(path: str, v: Any, want: str) -> str:
    shown = f'"{_type_name(v)}"'
    return f"{path}: Invalid value: {shown}: {path} in body must be of type {want}: {shown}"


_INT_BOUNDS = {"int32": (-(1 << 31), (1 << 31) - 1), "int64": (-(1 << 63), (1 << 63) - 1)}


def _walk(v: Any, schema: Dict[str, Any], path: str, errs: List[str]) -> None:
    if v is None:                      # a null is an absent field (the apiserver prunes it)
        return
    want = schema.get("type")
    int_or_string = schema.get("x-kubernetes-int-or-string") or path.startswith("spec.runPolicy.schedulingPolicy.minResources.")
    if int_or_string:
        if not (_is(v, "integer") or isinstance(v, str)):
            errs.append(_wrong(path, v, "integer or string"))
        return
    if want and not _is(v, want):
        errs.append(_wrong(path, v, want))
        return
    if want == "integer" and schema.get("format") in _INT_BOUNDS:
        lo, hi = _INT_BOUNDS[schema["format"]]
        if not lo <= v <= hi:
            errs.append(f"{path}: Invalid value: {v}: {path} in body must be of type {schema['format']}")
    if "enum" in schema and v not in schema["enum"] and v != "":
        errs.append(f'{path}: Unsupported value: "{v}": supported values: ' + ", ".join(f'"{e}"' for e in schema["enum"]))
    if want == "object" and not schema.get("x-kubernetes-preserve-unknown-fields"):
        props = schema.get("properties") or {}
        extra = schema.get("additionalProperties")
        for k, item in v.items():
            if k in props:
                _walk(item, props[k], f"{path}.{k}" if path else k, errs)
            elif isinstance(extra, dict):
                _walk(item, extra, f"{path}.{k}" if path else k, errs)
    elif want == "array" and isinstance(schema.get("items"), dict):
        for i, item in enumerate(v):
            _walk(item, schema["items"], f"{path}[{i}]", errs)


@lru_cache(maxsize=1)
def _crd_schema() -> Dict[str, Any]:
    from . import openapi
    return openapi.crd()["spec"]["versions"][0]["schema"]["openAPIV3Schema"]


def _string_map(v: Any, path: str, errs: List[str]) -> None:
    if v is None:
        return
    if not isinstance(v, dict):
        errs.append(_wrong(path, v, "object"))
        return
    for k, item in v.items():
        if item is not None and not isinstance(item, str):
            errs.append(_wrong(f"{path}.{k}", item, "string"))


def _object_meta(md: Any, path: str, errs: List[str]) -> None:
    if md is None:
        return
    if not isinstance(md, dict):
        errs.append(_wrong(path, md, "object"))
        return
    for k in ("name", "namespace", "generateName", "uid", "resourceVersion"):
        if md.get(k) is not None and not isinstance(md[k], str):
            errs.append(_wrong(f"{path}.{k}", md[k], "string"))
    _string_map(md.get("labels"), f"{path}.labels", errs)
    _string_map(md.get("annotations"), f"{path}.annotations", errs)
    for k in ("ownerReferences", "finalizers"):
        if md.get(k) is not None and not isinstance(md[k], list):
            errs.append(_wrong(f"{path}.{k}", md[k], "array"))


def _string_list(v: Any, path: str, errs: List[str]) -> None:
    if v is None:
        return
    if not isinstance(v, list):
        errs.append(_wrong(path, v, "array"))
        return
    for i, item in enumerate(v):
        if not isinstance(item, str):
            errs.append(_wrong(f"{path}[{i}]", item, "string"))


def _containers(v: Any, path: str, errs: List[str]) -> None:
    if v is None:
        return
    if not isinstance(v, list):
        errs.append(_wrong(path, v, "array"))
        return
    for i, c in enumerate(v):
        p = f"{path}[{i}]"
        if not isinstance(c, dict):
            errs.append(_wrong(p, c, "object"))
            continue
        for k in ("name", "image", "workingDir", "imagePullPolicy"):
            if c.get(k) is not None and not isinstance(c[k], str):
                errs.append(_wrong(f"{p}.{k}", c[k], "string"))
        _string_list(c.get("command"), f"{p}.command", errs)
        _string_list(c.get("args"), f"{p}.args", errs)
        env = c.get("env")
        if env is not None and not isinstance(env, list):
            errs.append(_wrong(f"{p}.env", env, "array"))
        for j, e in enumerate(env if isinstance(env, list) else []):
            if not isinstance(e, dict):
                errs.append(_wrong(f"{p}.env[{j}]", e, "object"))
            else:
                for k in ("name", "value"):
                    if e.get(k) is not None and not isinstance(e[k], str):
                        errs.append(_wrong(f"{p}.env[{j}].{k}", e[k], "string"))
        res = c.get("resources")
        if res is not None and not isinstance(res, dict):
            errs.append(_wrong(f"{p}.resources", res, "object"))
        for k in ("limits", "requests"):
            q = res.get(k) if isinstance(res, dict) else None
            if q is not None and not isinstance(q, dict):
                errs.append(_wrong(f"{p}.resources.{k}", q, "object"))
            for rk, rv in (q.items() if isinstance(q, dict) else ()):
                if rv is not None and not (isinstance(rv, str) or _is(rv, "number")):
                    errs.append(_wrong(f"{p}.resources.{k}.{rk}", rv, "integer or string"))
        for k in ("volumeMounts", "ports", "envFrom"):
            if c.get(k) is not None and not isinstance(c[k], list):
                errs.append(_wrong(f"{p}.{k}", c[k], "array"))


def _pod_template(t: Any, path: str, errs: List[str]) -> None:
    if t is None or not isinstance(t, dict):       # the CRD schema already said "object"
        return
    _object_meta(t.get("metadata"), f"{path}.metadata", errs)
    spec = t.get("spec")
    if spec is None:
        return
    if not isinstance(spec, dict):
        errs.append(_wrong(f"{path}.spec", spec, "object"))
        return
    _containers(spec.get("containers"), f"{path}.spec.containers", errs)
    _containers(spec.get("initContainers"), f"{path}.spec.initContainers", errs)
    for k in ("restartPolicy", "schedulerName", "priorityClassName", "hostname", "subdomain", "serviceAccountName", "nodeName"):
        if spec.get(k) is not None and not isinstance(spec[k], str):
            errs.append(_wrong(f"{path}.spec.{k}", spec[k], "string"))
    _string_map(spec.get("nodeSelector"), f"{path}.spec.nodeSelector", errs)
    for k in ("volumes", "tolerations", "imagePullSecrets"):
        if spec.get(k) is not None and not isinstance(spec[k], list):
            errs.append(_wrong(f"{path}.spec.{k}", spec[k], "array"))
    for k in ("hostNetwork", "hostPID", "hostIPC"):
        if spec.get(k) is not None and not isinstance(spec[k], bool):
            errs.append(_wrong(f"{path}.spec.{k}", spec[k], "boolean"))
    for k in ("terminationGracePeriodSeconds", "activeDeadlineSeconds", "priority"):
        if spec.get(k) is not None and not _is(spec[k], "integer"):
            errs.append(_wrong(f"{path}.spec.{k}", spec[k], "integer"))


def structural_errors(obj: Any) -> List[str]:
    """Type errors of an MPIJob in the apiserver's wording; [] when the object is structurally sound."""
    errs: List[str] = []
    if not isinstance(obj, dict):
        return [_wrong("", obj, "object").lstrip(": ")]
    _walk({k: v for k, v in obj.items() if k != "metadata"}, _crd_schema(), "", errs)
    _object_meta(obj.get("metadata"), "metadata", errs)
    spec = obj.get("spec")
    specs = spec.get("mpiReplicaSpecs") if isinstance(spec, dict) else None
    if isinstance(specs, dict):
        for rtype, rs in specs.items():
            if isinstance(rs, dict):
                _pod_template(rs.get("template"), f"spec.mpiReplicaSpecs.{rtype}.template", errs)
                _template_quantities(rs.get("template"), f"spec.mpiReplicaSpecs.{rtype}.template", errs)
    rp = spec.get("runPolicy") if isinstance(spec, dict) else None
    sp = rp.get("schedulingPolicy") if isinstance(rp, dict) else None
    _quantities(sp.get("minResources") if isinstance(sp, dict) else None, "spec.runPolicy.schedulingPolicy.minResources", errs)
    return errs


_QUANTITY = __import__("re").compile(r"^[+-]?(\d+\.?\d*|\.\d+)(([KMGTPE]i)|[numkMGTPE]|([eE][+-]?\d+))?$")


def _quantities(q: Any, path: str, errs: List[str]) -> None:
    """resource.Quantity syntax for the values of a limits / requests / minResources map; extended resources
    (nvidia.com/gpu) must be whole numbers."""
    if not isinstance(q, dict):
        return
    for k, v in q.items():
        if v is None or isinstance(v, bool) or isinstance(v, (list, dict)):
            continue                                   # the type check already reported it
        text = str(v).strip()
        if not _QUANTITY.match(text):
            errs.append(f'{path}.{k}: Invalid value: "{text}": quantities must match the regular expression '
                        "'^([+-]?[0-9.]+)([eEinumkKMGTP]*[-+]?[0-9]*)$'")
        elif "/" in k and not k.startswith("kubernetes.io/") and not text.lstrip("+").isdigit():
            errs.append(f'{path}.{k}: Invalid value: "{text}": must be an integer')


def _template_quantities(t: Any, path: str, errs: List[str]) -> None:
    spec = t.get("spec") if isinstance(t, dict) else None
    for field in ("containers", "initContainers"):
        cs = spec.get(field) if isinstance(spec, dict) else None
        for i, c in enumerate(cs if isinstance(cs, list) else []):
            res = c.get("resources") if isinstance(c, dict) else None
            for k in ("limits", "requests"):
                _quantities(res.get(k) if isinstance(res, dict) else None, f"{path}.spec.{field}[{i}].resources.{k}", errs)


def core_structural_errors(resource: str, obj: Any, required: bool = False) -> List[str]:
    """The same service for the core objects the REST API accepts (pods, jobs, services, configmaps, secrets, pod groups ...):
    the shapes the node agent and the controller index into. kube-apiserver decodes these into typed structs; a map where a
    list belongs never gets past it. ``required`` adds the required fields (a pod needs a spec with containers): the REST
    admission asks for them, the store itself - like the fake clientset of the reference's unit tests - only checks types."""
    errs: List[str] = []
    if not isinstance(obj, dict):
        return [f"body must be of type object: \"{_type_name(obj)}\""]
    _object_meta(obj.get("metadata"), "metadata", errs)
    for k in ("spec", "status", "data", "stringData"):
        if obj.get(k) is not None and not isinstance(obj[k], dict):
            errs.append(_wrong(k, obj[k], "object"))
    if errs:
        return errs
    if resource == "pods":
        _pod_template({"spec": obj.get("spec")}, "", errs)       # a pod IS a template: paths come out with a leading dot
        _template_quantities(obj, "", errs)
        errs[:] = [e.replace(": .", ": ").lstrip(".") for e in errs]
        if required:
            spec = obj.get("spec")
            if not isinstance(spec, dict):
                errs.append("spec: Required value")
            elif not isinstance(spec.get("containers"), list) or not spec["containers"]:
                if not any(e.startswith("spec.containers:") for e in errs):
                    errs.append("spec.containers: Required value")
    elif resource == "jobs":
        spec = obj.get("spec") or {}
        for k in ("backoffLimit", "activeDeadlineSeconds", "ttlSecondsAfterFinished", "parallelism", "completions"):
            if spec.get(k) is not None and not _is(spec[k], "integer"):
                errs.append(_wrong(f"spec.{k}", spec[k], "integer"))
        if spec.get("suspend") is not None and not isinstance(spec["suspend"], bool):
            errs.append(_wrong("spec.suspend", spec["suspend"], "boolean"))
        t = spec.get("template")
        if t is not None and not isinstance(t, dict):
            errs.append(_wrong("spec.template", t, "object"))
        _pod_template(t, "spec.template", errs)
        _template_quantities(t, "spec.template", errs)
    elif resource in ("configmaps", "secrets"):
        _string_map(obj.get("data"), "data", errs)
    elif resource in ("volcano-podgroups", "sched-podgroups"):
        spec = obj.get("spec") or {}
        for k in ("minMember", "scheduleTimeoutSeconds"):
            if spec.get(k) is not None and not _is(spec[k], "integer"):
                errs.append(_wrong(f"spec.{k}", spec[k], "integer"))
        mr = spec.get("minResources")
        if mr is not None and not isinstance(mr, dict):
            errs.append(_wrong("spec.minResources", mr, "object"))
        _quantities(mr, "spec.minResources", errs)
    return errs
