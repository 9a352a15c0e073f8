"""ObjectMeta / OwnerReference / time helpers for dict-shaped API objects.

Every non-MPIJob object in the local store (Pod, Service, ConfigMap, Secret,
Job, PodGroup, PriorityClass, Event, Lease) is a plain dict in Kubernetes JSON
shape, which keeps golden-object tests (reference:
pkg/controller/mpi_job_controller_test.go:1424-1890) direct comparisons.
"""
from __future__ import annotations

import copy
import datetime as _dt
import uuid
from typing import Any, Dict, List, Optional

from . import constants as C


def now_rfc3339(t: Optional[float] = None) -> str:
    d = _dt.datetime.now(_dt.timezone.utc) if t is None else _dt.datetime.fromtimestamp(t, _dt.timezone.utc)
    return d.replace(microsecond=0).strftime("%Y-%m-%dT%H:%M:%SZ")


def parse_rfc3339(s: str) -> float:
    return _dt.datetime.strptime(s, "%Y-%m-%dT%H:%M:%SZ").replace(tzinfo=_dt.timezone.utc).timestamp()


import re as _re

_DNS1123_SUBDOMAIN = _re.compile(r"^[a-z0-9]([-a-z0-9]*[a-z0-9])?(\.[a-z0-9]([-a-z0-9]*[a-z0-9])?)*$")


def name_problem(name: str, what: str = "metadata.name") -> Optional[str]:
    """The apiserver's generic object-name rule (apimachinery validation.IsDNS1123Subdomain): lower-case alphanumerics, '-' and
    '.', alphanumeric at both ends of every label, at most 253 characters. Here it is also what keeps object names - which become
    directory names under the node agent's state dir - from containing '/' or '..'. None when the name is fine."""
    if not isinstance(name, str) or not name:
        return f"{what}: Required value"
    if len(name) > 253:
        return f"{what}: Invalid value: must be no more than 253 characters"
    if not _DNS1123_SUBDOMAIN.match(name):
        return (f"{what}: Invalid value: \"{name}\": a lowercase RFC 1123 subdomain must consist of lower case alphanumeric "
                "characters, '-' or '.', and must start and end with an alphanumeric character")
    return None


def new_uid() -> str:
    return str(uuid.uuid4())


def deepcopy(obj):
    return copy.deepcopy(obj)


def meta(obj: Dict[str, Any]) -> Dict[str, Any]:
    return obj.setdefault("metadata", {})


def name_of(obj) -> str:
    return meta(obj).get("name", "")


def namespace_of(obj) -> str:
    return meta(obj).get("namespace", "")


def key_of(obj) -> str:
    ns = namespace_of(obj)
    return f"{ns}/{name_of(obj)}" if ns else name_of(obj)


def split_key(key: str):
    """cache.SplitMetaNamespaceKey: 'ns/name' | 'name'; anything else is invalid."""
    parts = key.split("/")
    if len(parts) == 1:
        return "", parts[0]
    if len(parts) == 2:
        return parts[0], parts[1]
    raise ValueError(f"unexpected key format: {key!r}")


def new_controller_ref(owner: Dict[str, Any], api_version: str = C.API_VERSION, kind: str = C.KIND) -> Dict[str, Any]:
    """metav1.NewControllerRef."""
    return {
        "apiVersion": api_version, "kind": kind, "name": name_of(owner), "uid": meta(owner).get("uid", ""),
        "controller": True, "blockOwnerDeletion": True,
    }


def get_controller_of(obj) -> Optional[Dict[str, Any]]:
    for ref in meta(obj).get("ownerReferences", []) or []:
        if ref.get("controller"):
            return ref
    return None


def is_controlled_by(obj, owner) -> bool:
    ref = get_controller_of(obj)
    return ref is not None and ref.get("uid", "") == meta(owner).get("uid", "")


def label_selector_matches(selector: Dict[str, str], labels: Optional[Dict[str, str]]) -> bool:
    labels = labels or {}
    return all(labels.get(k) == v for k, v in (selector or {}).items())
