"""Shared fixtures for control-plane tests."""
import copy

from mpi_operator_b200.api import constants as C
from mpi_operator_b200.api.types import MPIJob, MPIJobSpec, ReplicaSpec, RunPolicy


def template(cmd=None, args=None, gpus=0, **spec_extra):
    c = {"name": "main", "image": "test-image"}
    if cmd:
        c["command"] = list(cmd)
    if args:
        c["args"] = list(args)
    if gpus:
        c["resources"] = {"limits": {"nvidia.com/gpu": gpus}}
    spec = {"containers": [c]}
    spec.update(spec_extra)
    return {"spec": spec}


def new_mpijob(name="test", namespace="default", workers=2, launcher_cmd=("mpirun",), launcher_args=("-n", "2", "true"),
               worker_cmd=None, impl="", slots=None, clean=None, **run_policy) -> MPIJob:
    specs = {C.REPLICA_TYPE_LAUNCHER: ReplicaSpec(template=template(launcher_cmd, launcher_args))}
    if workers is not None:
        specs[C.REPLICA_TYPE_WORKER] = ReplicaSpec(replicas=workers, template=template(worker_cmd))
    return MPIJob(metadata={"name": name, "namespace": namespace, "uid": f"uid-{name}"},
                  spec=MPIJobSpec(slots_per_worker=slots, mpi_implementation=impl, mpi_replica_specs=specs,
                                  run_policy=RunPolicy(clean_pod_policy=clean, **run_policy)))


def conds(job_dict_or_obj):
    d = job_dict_or_obj.to_dict() if hasattr(job_dict_or_obj, "to_dict") else job_dict_or_obj
    return {c["type"]: c["status"] for c in d.get("status", {}).get("conditions", []) or []}
