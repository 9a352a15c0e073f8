"""Scheme registration: group/version/kind plumbing and the decoder registry.

Reference: pkg/apis/kubeflow/v2beta1/register.go:23-52 (+ doc.go,
zz_generated.defaults.go:29-44 which wires SetDefaults_MPIJob into the scheme).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Tuple

from . import constants as C
from .defaults import set_defaults_mpijob
from .types import MPIJob, MPIJobList

GroupVersionKind = Tuple[str, str, str]
SCHEME_GROUP_VERSION = (C.GROUP_NAME, C.GROUP_VERSION)
SCHEME_GROUP_VERSION_KIND: GroupVersionKind = (C.GROUP_NAME, C.GROUP_VERSION, C.KIND)


def resource(name: str) -> Tuple[str, str]:
    """register.go Resource(): qualified GroupResource."""
    return (C.GROUP_NAME, name)


class Scheme:
    """Maps (group, version, kind) to decoders and defaulting functions."""

    def __init__(self):
        self._types: Dict[GroupVersionKind, Callable[[dict], Any]] = {}
        self._defaulters: Dict[type, Callable[[Any], Any]] = {}

    def add_known_type(self, gvk: GroupVersionKind, decoder: Callable[[dict], Any]) -> None:
        self._types[gvk] = decoder

    def add_defaulting_func(self, typ: type, fn: Callable[[Any], Any]) -> None:
        self._defaulters[typ] = fn

    def recognizes(self, api_version: str, kind: str) -> bool:
        g, _, v = api_version.partition("/")
        return (g, v, kind) in self._types

    def decode(self, doc: dict):
        api_version, kind = doc.get("apiVersion", ""), doc.get("kind", "")
        g, _, v = api_version.partition("/")
        dec = self._types.get((g, v, kind))
        if dec is None:
            raise ValueError(f'no kind "{kind}" is registered for version "{api_version}"')
        return dec(doc)

    def default(self, obj):
        fn = self._defaulters.get(type(obj))
        return fn(obj) if fn else obj


def add_to_scheme(scheme: Scheme) -> Scheme:
    scheme.add_known_type(SCHEME_GROUP_VERSION_KIND, MPIJob.from_dict)
    scheme.add_known_type((C.GROUP_NAME, C.GROUP_VERSION, "MPIJobList"), MPIJobList.from_dict)
    scheme.add_defaulting_func(MPIJob, set_defaults_mpijob)
    return scheme


scheme = add_to_scheme(Scheme())
