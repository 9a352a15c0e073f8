"""Prometheus metrics with the reference's metric names (part of the surface;
README.md:227-239, pkg/controller/mpi_job_controller.go:125-140,
cmd/mpi-operator/app/server.go:73-77) plus data-plane additions
(SURVEY.md §5.5 [NEW])."""
from __future__ import annotations

from prometheus_client import CollectorRegistry, Counter, Gauge, Histogram, generate_latest

REGISTRY = CollectorRegistry()

# prometheus_client appends "_total" to counters: declare them without the suffix
mpi_jobs_created = Counter("mpi_operator_jobs_created", "Counts number of MPI jobs created", registry=REGISTRY)
mpi_jobs_successful = Counter("mpi_operator_jobs_successful", "Counts number of MPI jobs successful", registry=REGISTRY)
mpi_jobs_failed = Counter("mpi_operator_jobs_failed", "Counts number of MPI jobs failed", registry=REGISTRY)
mpi_job_info = Gauge("mpi_operator_job_info", "Information about MPIJob", ["launcher", "namespace"], registry=REGISTRY)
is_leader = Gauge("mpi_operator_is_leader", "Is this client the leader of this mpi-operator client set?", registry=REGISTRY)

# data plane (new)
allreduce_bytes = Counter("b200mpi_allreduce_bytes", "Bytes reduced by b200mpi allreduce kernels", ["algo"], registry=REGISTRY)
collective_calls = Counter("b200mpi_collective_calls", "Collective kernels launched by finished ranks", ["op", "algo"], registry=REGISTRY)
collective_bytes = Counter("b200mpi_collective_bytes", "Payload bytes moved by collective kernels of finished ranks", ["op", "algo"],
                           registry=REGISTRY)
hvd_tensors = Counter("b200mpi_hvd_tensors", "Collectives negotiated by the Horovod-core engine of finished ranks", registry=REGISTRY)
hvd_fused_groups = Counter("b200mpi_hvd_fused_groups", "Fused allreduce groups executed by the engine", registry=REGISTRY)
hvd_cache_hits = Counter("b200mpi_hvd_response_cache_hits", "Submissions sent as cached ids", registry=REGISTRY)
hvd_negotiation_bytes = Counter("b200mpi_hvd_negotiation_bytes", "Bytes exchanged by the engine's negotiation", registry=REGISTRY)
hvd_stall_warnings = Counter("b200mpi_hvd_stall_warnings", "Stall-inspector warnings", registry=REGISTRY)
ranks_active = Gauge("b200mpi_ranks_active", "Ranks currently running under the node agent", registry=REGISTRY)
gpu_slots_free = Gauge("b200mpi_gpu_slots_free", "Unallocated GPU slots on this box", registry=REGISTRY)
gpu_healthy = Gauge("b200mpi_gpu_healthy", "1 when the GPU passed its last health probe (node/health.py), 0 when it is cordoned by it", ["gpu"],
                    registry=REGISTRY)
job_duration_seconds = Histogram("mpi_operator_job_duration_seconds", "startTime -> completionTime of finished MPIJobs", ["result"],
                                 buckets=(1, 5, 15, 30, 60, 120, 300, 600, 1800, 3600, 4 * 3600, 24 * 3600), registry=REGISTRY)
reconcile_seconds = Histogram("mpi_operator_reconcile_duration_seconds", "Wall time of one syncHandler call",
                              buckets=(0.001, 0.005, 0.01, 0.05, 0.1, 0.5, 1, 5), registry=REGISTRY)


def observe_rank_stats(stats: dict) -> None:
    """Fold one rank's ``Communicator.stats()`` dump into the data-plane counters (node agent, on pod exit)."""
    for o in stats.get("ops", []):
        op, algo, calls, nbytes = str(o.get("op", "")), str(o.get("algo", "")), int(o.get("calls", 0)), int(o.get("bytes", 0))
        if not op or calls <= 0:
            continue
        collective_calls.labels(op=op, algo=algo).inc(calls)
        collective_bytes.labels(op=op, algo=algo).inc(nbytes)
        if op.startswith("allreduce"):
            allreduce_bytes.labels(algo=algo).inc(nbytes)
    h = stats.get("hvd") or {}     # hvd.engine_stats() of the rank (csrc/hvd_core), if the job used the Horovod front-end
    for counter, key in ((hvd_tensors, "tensors"), (hvd_fused_groups, "fused_groups"), (hvd_cache_hits, "cache_hits"),
                         (hvd_negotiation_bytes, "negotiation_bytes"), (hvd_stall_warnings, "stall_warnings")):
        if int(h.get(key, 0)) > 0:
            counter.inc(int(h[key]))


def render() -> bytes:
    return generate_latest(REGISTRY)


def counter_value(counter) -> float:
    return counter._value.get()  # test helper


def observe_job_duration(job, result: str) -> None:
    """``result``: Succeeded | Failed. Uses the status' own timestamps (RFC 3339); silently skips jobs without a start time."""
    try:
        from ..api.meta import parse_rfc3339
        st, ct = job.status.start_time, job.status.completion_time
        if st and ct:
            job_duration_seconds.labels(result=result).observe(max(0.0, parse_rfc3339(ct) - parse_rfc3339(st)))
    except Exception:  # noqa: BLE001
        pass
