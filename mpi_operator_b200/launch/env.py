"""Per-rank environment contract.

The reference hands the workload its identity through the MPI launcher's
environment (``OMPI_COMM_WORLD_*`` for Open MPI,
``PMI_*`` for Hydra-based MPICH / Intel MPI) and through ``K_MPI_JOB_ROLE``
(pkg/controller/mpi_job_controller.go:169-180).  Our spawner sets *all* of
those dialects plus the torch.distributed one, and this module is the single
place that reads them back.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Mapping, Optional


@dataclass(frozen=True)
class RankInfo:
    rank: int
    world_size: int
    local_rank: int
    local_size: int
    node_rank: int
    job_id: str


_RANK_VARS = ("B200MPI_RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "PMIX_RANK", "RANK")
_SIZE_VARS = ("B200MPI_WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "WORLD_SIZE")
_LRANK_VARS = ("B200MPI_LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "LOCAL_RANK")
_LSIZE_VARS = ("B200MPI_LOCAL_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "LOCAL_WORLD_SIZE")


def _first(env: Mapping[str, str], names, default: Optional[int] = None) -> Optional[int]:
    for n in names:
        v = env.get(n)
        if v is not None and v != "":
            return int(v)
    return default


def job_id_from_env(env: Mapping[str, str]) -> str:
    """Rendezvous key shared by all ranks of one job."""
    if env.get("B200MPI_JOB_ID"):
        return env["B200MPI_JOB_ID"]
    # torchrun / torch.distributed style launch: the master endpoint identifies the job
    if env.get("MASTER_PORT"):
        run_id = env.get("TORCHELASTIC_RUN_ID", "")
        return f"torch-{env.get('MASTER_ADDR', '127.0.0.1')}-{env['MASTER_PORT']}-{run_id}"
    return f"solo-{os.getpid()}"


def rank_info_from_env(env: Optional[Mapping[str, str]] = None) -> RankInfo:
    env = os.environ if env is None else env
    rank = _first(env, _RANK_VARS, 0)
    world = _first(env, _SIZE_VARS, 1)
    lrank = _first(env, _LRANK_VARS, rank)
    lsize = _first(env, _LSIZE_VARS, world)
    node = _first(env, ("OMPI_COMM_WORLD_NODE_RANK", "GROUP_RANK"), 0)
    return RankInfo(rank, world, lrank, lsize, node, job_id_from_env(env))


def build_rank_env(*, rank: int, world_size: int, local_rank: int, local_size: int, node_rank: int = 0,
                   job_id: str, gpu: Optional[int] = None, master_addr: str = "127.0.0.1", master_port: int = 29500,
                   hostname: Optional[str] = None, role: str = "worker") -> Dict[str, str]:
    """Environment for one spawned rank, in every dialect workloads read.

    ``gpu`` pins the rank to one device via CUDA_VISIBLE_DEVICES (so that
    ``local_rank`` indexing inside frameworks stays valid we export the full
    slot list separately in B200MPI_GPU).
    """
    e = {
        # ours
        "B200MPI_JOB_ID": job_id, "B200MPI_RANK": str(rank), "B200MPI_WORLD_SIZE": str(world_size),
        "B200MPI_LOCAL_RANK": str(local_rank), "B200MPI_LOCAL_SIZE": str(local_size),
        # Open MPI (orted) dialect
        "OMPI_COMM_WORLD_RANK": str(rank), "OMPI_COMM_WORLD_SIZE": str(world_size),
        "OMPI_COMM_WORLD_LOCAL_RANK": str(local_rank), "OMPI_COMM_WORLD_LOCAL_SIZE": str(local_size),
        "OMPI_COMM_WORLD_NODE_RANK": str(node_rank),
        # Hydra (MPICH / Intel MPI) dialect
        "PMI_RANK": str(rank), "PMI_SIZE": str(world_size), "MPI_LOCALRANKID": str(local_rank),
        "MPI_LOCALNRANKS": str(local_size),
        # torch.distributed dialect
        "RANK": str(rank), "WORLD_SIZE": str(world_size), "LOCAL_RANK": str(local_rank),
        "LOCAL_WORLD_SIZE": str(local_size), "GROUP_RANK": str(node_rank),
        "MASTER_ADDR": master_addr, "MASTER_PORT": str(master_port),
        # Horovod dialect
        "HOROVOD_RANK": str(rank), "HOROVOD_SIZE": str(world_size), "HOROVOD_LOCAL_RANK": str(local_rank),
        "HOROVOD_LOCAL_SIZE": str(local_size), "HOROVOD_CROSS_RANK": str(node_rank), "HOROVOD_CROSS_SIZE": "1",
        # reference contract (pkg/controller/mpi_job_controller.go:169-180)
        "K_MPI_JOB_ROLE": role,
    }
    if hostname:
        e["HOSTNAME"] = hostname
        e["B200MPI_HOSTNAME"] = hostname
    if gpu is not None:
        e["B200MPI_GPU"] = str(gpu)
    return e
