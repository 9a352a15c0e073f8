"""MPIJob status / condition bookkeeping.

Reference: pkg/controller/mpi_job_controller_status.go:24-144 — reasons, and
the exact update rules: a condition is untouched when type+status+reason are
unchanged; lastTransitionTime survives when only the reason changes; setting
Succeeded/Failed flips existing Running/Failed conditions to False; Running
and Restarting are mutually exclusive.
"""
from __future__ import annotations

from typing import Optional

from ..api import constants as C
from ..api import meta as M
from ..api.types import JobCondition, JobStatus, MPIJob, ReplicaStatus

MPIJOB_CREATED_REASON = "MPIJobCreated"
MPIJOB_SUCCEEDED_REASON = "MPIJobSucceeded"
MPIJOB_RUNNING_REASON = "MPIJobRunning"
MPIJOB_SUSPENDED_REASON = "MPIJobSuspended"
MPIJOB_RESUMED_REASON = "MPIJobResumed"
MPIJOB_FAILED_REASON = "MPIJobFailed"
MPIJOB_EVICT = "MPIJobEvicted"


def initialize_replica_statuses(job: MPIJob, rtype: str) -> None:
    job.status.replica_statuses[rtype] = ReplicaStatus()


def new_condition(ctype: str, status: str, reason: str, message: str, now: Optional[str] = None) -> JobCondition:
    now = now or M.now_rfc3339()
    return JobCondition(type=ctype, status=status, reason=reason, message=message, last_update_time=now,
                        last_transition_time=now)


def get_condition(status: JobStatus, ctype: str) -> Optional[JobCondition]:
    for c in status.conditions:
        if c.type == ctype:
            return c
    return None


def has_condition(status: JobStatus, ctype: str) -> bool:
    return any(c.type == ctype and c.status == C.CONDITION_TRUE for c in status.conditions)


def is_succeeded(status: JobStatus) -> bool:
    return has_condition(status, C.JOB_SUCCEEDED)


def is_failed(status: JobStatus) -> bool:
    return has_condition(status, C.JOB_FAILED)


def is_finished(status: JobStatus) -> bool:
    return is_succeeded(status) or is_failed(status)


def _filter_out(conditions, ctype: str):
    out = []
    for c in conditions:
        if ctype == C.JOB_RESTARTING and c.type == C.JOB_RUNNING:
            continue
        if ctype == C.JOB_RUNNING and c.type == C.JOB_RESTARTING:
            continue
        if c.type == ctype:
            continue
        if ctype in (C.JOB_FAILED, C.JOB_SUCCEEDED) and c.type in (C.JOB_RUNNING, C.JOB_FAILED):
            c = JobCondition(**{**c.__dict__, "status": C.CONDITION_FALSE})
        out.append(c)
    return out


def set_condition(status: JobStatus, cond: JobCondition) -> bool:
    cur = get_condition(status, cond.type)
    if cur is not None and cur.status == cond.status and cur.reason == cond.reason:
        return False
    if cur is not None and cur.status == cond.status:
        cond.last_transition_time = cur.last_transition_time
    status.conditions = _filter_out(status.conditions, cond.type) + [cond]
    return True


def update_mpijob_conditions(job: MPIJob, ctype: str, status: str, reason: str, message: str,
                             now: Optional[str] = None) -> bool:
    return set_condition(job.status, new_condition(ctype, status, reason, message, now))
