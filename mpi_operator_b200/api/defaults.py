"""Defaulting of MPIJob objects.

Reference: pkg/apis/kubeflow/v2beta1/default.go:27-80 (SetDefaults_MPIJob and
helpers) — behaviour reproduced, including the quirk that cleanPodPolicy
defaults to ``None`` although its field doc says "Default to Running"
(SURVEY.md §3.7).
"""
from __future__ import annotations

from . import constants as C
from .types import MPIJob, ReplicaSpec


def set_defaults_launcher(spec: ReplicaSpec | None) -> None:
    """default.go:27-37."""
    if spec is None:
        return
    if spec.restart_policy == "":
        spec.restart_policy = C.DEFAULT_LAUNCHER_RESTART_POLICY
    if spec.replicas is None:
        spec.replicas = 1


def set_defaults_worker(spec: ReplicaSpec | None) -> None:
    """default.go:40-50."""
    if spec is None:
        return
    if spec.restart_policy == "":
        spec.restart_policy = C.DEFAULT_RESTART_POLICY
    if spec.replicas is None:
        spec.replicas = 0


def set_defaults_run_policy(job: MPIJob) -> None:
    """default.go:52-58: TTL, deadline, backoff, scheduling policy stay nil."""
    if job.spec.run_policy.clean_pod_policy is None:
        job.spec.run_policy.clean_pod_policy = C.CLEAN_POD_POLICY_NONE


def normalize_pod_template(spec: ReplicaSpec | None) -> None:
    """Not in the reference (there the API server's schema rejects it): a container ``command`` / ``args`` given as one
    string becomes a one-element list. The reference's own SDK example passes ``command="mpirun"``
    (sdk/python/v2beta1/tensorflow-mnist.py:30,44-47); without an API-server schema in front of us, accept what it means."""
    if spec is None or not isinstance(spec.template, dict):
        return
    for c in ((spec.template.get("spec") or {}).get("containers") or []):
        if isinstance(c, dict):
            for key in ("command", "args"):
                if isinstance(c.get(key), str):
                    c[key] = [c[key]]


def set_defaults_mpijob(job: MPIJob) -> MPIJob:
    """default.go:60-80. Mutates and returns ``job``."""
    set_defaults_run_policy(job)
    if job.spec.slots_per_worker is None:
        job.spec.slots_per_worker = 1
    if job.spec.ssh_auth_mount_path == "":
        job.spec.ssh_auth_mount_path = "/root/.ssh"
    if job.spec.mpi_implementation == "":
        job.spec.mpi_implementation = C.MPI_IMPLEMENTATION_OPENMPI
    if job.spec.launcher_creation_policy == "":
        job.spec.launcher_creation_policy = C.LAUNCHER_CREATION_POLICY_AT_STARTUP
    specs = job.spec.mpi_replica_specs or {}
    set_defaults_launcher(specs.get(C.REPLICA_TYPE_LAUNCHER))
    set_defaults_worker(specs.get(C.REPLICA_TYPE_WORKER))
    for rs in specs.values():
        normalize_pod_template(rs)
    return job
