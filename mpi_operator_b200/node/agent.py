"""NodeAgent = scheduler + kubelet + batch/v1 Job controller for one box.

What Kubernetes does for the reference between "controller created a Pod/Job
object" and "controller observes its phase" (the k8s boundary
line) happens here:

* Job controller: one active pod per launcher Job, ``backoffLimit`` (default 6),
  ``activeDeadlineSeconds``, ``suspend``, ``ttlSecondsAfterFinished``,
  ``Complete`` / ``Failed`` conditions (reasons BackoffLimitExceeded /
  DeadlineExceeded) — the fields the reference copies into the launcher Job at
  pkg/controller/mpi_job_controller.go:1540-1545 and reads back at :1103-1130.
* Scheduler: GPU slots from the discovered topology; PodGroup members are
  granted all-or-nothing (gang); unschedulable pods stay Pending.
* Kubelet: volumes (ConfigMap/Secret) materialised under the pod directory,
  container[0] started as a process group, phases Pending -> Running ->
  Succeeded/Failed, Ready condition, restartPolicy OnFailure/Never, logs.
  A worker whose command is ``sshd`` (the reference default,
  mpi_job_controller.go:1503-1505) is an idle placeholder: on one box ranks are
  spawned by the native mpirun, no ssh needed.
"""
from __future__ import annotations

import base64
import collections
import copy
import json
import logging
import os
import random
import signal
import string
import subprocess
import threading
import time
from typing import Dict, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..client import errors
from ..client.store import DELETED, MODIFIED, ObjectStore
from ..controller import metrics
from .allocator import GangAllocator, SlotRequest
from .topology import Topology, discover_topology

log = logging.getLogger("node-agent")

NODE_NAME = "localhost"
GPU_ANNOTATION = "b200mpi.kubeflow.org/gpus"
PGID_ANNOTATION = "b200mpi.kubeflow.org/pgid"
JOB_UID_LABEL = "batch.kubernetes.io/controller-uid"
JOB_NAME_LABEL = "batch.kubernetes.io/job-name"
DEFAULT_BACKOFF_LIMIT = 6
PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN_DIR = os.path.join(PKG_DIR, "bin")


REPO_ROOT = os.path.dirname(PKG_DIR)
# "Images" on a single box: a container image is a directory + command remaps. The reference's
# example images resolve to the in-repo equivalents of what those images contain.
# An image may carry an "entrypoint" (used when the container sets no `command`, k8s ENTRYPOINT/CMD semantics) —
# the reference's MPICH and Intel images use build/base/entrypoint.sh (build/base/mpich.Dockerfile, intel.Dockerfile).
_ENTRYPOINT_SH = os.path.join(REPO_ROOT, "build", "base", "entrypoint.sh")
IMAGE_REGISTRY = {
    "mpioperator/tensorflow-benchmarks": {"workingDir": os.path.join(REPO_ROOT, "examples", "tensorflow-benchmarks")},
    "mpioperator/mpi-pi": {},
    "mpioperator/mpi-pi:mpich": {"entrypoint": ["bash", _ENTRYPOINT_SH]},
    "mpioperator/mpi-pi:intel": {"entrypoint": ["bash", _ENTRYPOINT_SH]},
    "mpioperator/mpich": {"entrypoint": ["bash", _ENTRYPOINT_SH]},
    "mpioperator/intel": {"entrypoint": ["bash", _ENTRYPOINT_SH]},
    "mpioperator/torch-ddp": {"workingDir": REPO_ROOT},
    "docker.io/kubeflow/mpi-horovod-mnist": {"remap": {"/examples/tensorflow_mnist.py": os.path.join(REPO_ROOT, "examples", "horovod", "torch_mnist.py")}},
}


def image_config(image: str) -> dict:
    base = (image or "").split(":")[0]
    extra = os.environ.get("B200MPI_IMAGE_REGISTRY")
    reg = dict(IMAGE_REGISTRY)
    if extra and os.path.exists(extra):
        with open(extra) as f:
            reg.update(json.load(f))
    return reg.get(image or "", reg.get(base, {}))


def _rand(n=5):
    return "".join(random.choice(string.ascii_lowercase + string.digits) for _ in range(n))


class _Proc:
    def __init__(self):
        self.popen: Optional[subprocess.Popen] = None
        self.virtual = False
        self.restarts = 0
        self.started_at = ""
        self.next_restart = 0.0
        self.log_path = ""
        self.pod_dir = ""


class NodeAgent:
    def __init__(self, store: ObjectStore, topology: Optional[Topology] = None, state_dir: Optional[str] = None,
                 clock=None, tick: float = 0.05, extra_env: Optional[Dict[str, str]] = None):
        self.store = store
        self.topology = topology or discover_topology()
        self.alloc = GangAllocator(self.topology)
        self.state_dir = state_dir or os.path.join(os.environ.get("TMPDIR", "/tmp"), f"b200mpi-node-{os.getpid()}")
        os.makedirs(self.state_dir, exist_ok=True)
        self.tick = tick
        self.extra_env = dict(extra_env or {})
        self._procs: Dict[str, _Proc] = {}
        self._lock = threading.RLock()
        self._wake = threading.Event()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._cancels = []
        self._job_backoff: Dict[str, float] = {}
        self._unsched_since: Dict[str, float] = {}
        self._deleted = collections.deque()

    # ------------------------------------------------------------ lifecycle --
    def _adopt_existing(self) -> None:
        """Daemon restart: the persisted store may hold pods that were Running under the
        previous daemon. Idle sshd placeholders are simply re-adopted (their GPU slots re-reserved); a pod whose
        process the old daemon owned cannot be waited on any more: its recorded process group is reaped and the
        pod is marked Failed (reason DaemonRestarted), so the Job controller starts a fresh launcher pod
        (counted against backoffLimit) — level-triggered recovery, like a kubelet restart."""
        for pod in self.store.list("pods"):
            if not pod["spec"].get("nodeName") or pod.get("status", {}).get("phase") != "Running":
                continue
            key = M.key_of(pod)
            gpus = [int(x) for x in (M.meta(pod).get("annotations") or {}).get(GPU_ANNOTATION, "").split(",") if x != ""]
            c0 = pod["spec"]["containers"][0]
            argv = list(c0.get("command") or []) + list(c0.get("args") or [])
            if argv and os.path.basename(argv[0]) == "sshd":
                pr = _Proc()
                pr.virtual, pr.pod_dir = True, self.pod_dir(pod)
                pr.started_at = pod["status"].get("startTime", M.now_rfc3339())
                self._procs[key] = pr
                self.alloc.adopt(key, gpus)
                continue
            pgid = (M.meta(pod).get("annotations") or {}).get(PGID_ANNOTATION)
            if pgid:
                try:
                    os.killpg(int(pgid), signal.SIGKILL)
                except (ProcessLookupError, PermissionError, ValueError):
                    pass
            pr = _Proc()
            pr.pod_dir, pr.log_path = self.pod_dir(pod), os.path.join(self.pod_dir(pod), "logs", "0.log")
            self._set_terminal(pod, pr, "Failed", 137, "DaemonRestarted", "the operator daemon restarted while this pod was running")

    def start(self) -> None:
        self._adopt_existing()
        for res in ("pods", "jobs", "configmaps", "volcano-podgroups", "sched-podgroups"):
            self._cancels.append(self.store.watch(res, self._on_event, replay=True))
        self._thread = threading.Thread(target=self._loop, name="node-agent", daemon=True)
        self._thread.start()

    def stop(self) -> None:
        self._stop.set()
        self._wake.set()
        if self._thread:
            self._thread.join(timeout=5)
        for c in self._cancels:
            c()
        with self._lock:
            for key in list(self._procs):
                self._kill(key, grace=0.5)

    def wake(self) -> None:
        """Run a sync round now (capacity changed: a GPU was cordoned / uncordoned)."""
        self._wake.set()

    def _on_event(self, etype, obj, old) -> None:
        # Never take the agent lock here: watch handlers run under the store's dispatch lock,
        # while the sync loop calls into the store holding the agent lock (lock-order inversion).
        if obj.get("kind") == "Pod" and etype == DELETED:
            self._deleted.append(M.key_of(obj))
        if obj.get("kind") == "ConfigMap" and etype == MODIFIED:
            self._cm_dirty = True
        self._wake.set()

    def _next_wait(self) -> float:
        """Store events wake the loop at once; the timed wake-up only exists to poll child processes and restart back-offs
        (every `tick`) and the second-granularity timers — deadlines, TTLs, schedule time-outs — for which an idle daemon
        does not need to re-list every pod 20 times a second."""
        for pr in list(self._procs.values()):
            if (pr.popen is not None and not pr.virtual) or pr.next_restart:
                return self.tick
        return max(self.tick, 0.5)

    def _loop(self) -> None:
        while not self._stop.is_set():
            self._wake.wait(self._next_wait())
            self._wake.clear()
            try:
                self.sync_once()
                if time.monotonic() - getattr(self, "_last_rotate", 0.0) > 10.0:
                    self._last_rotate = time.monotonic()
                    self.rotate_logs()
            except Exception:  # noqa: BLE001
                log.exception("node agent sync failed")

    # ---------------------------------------------------------------- sync --
    def sync_once(self) -> None:
        with self._lock:
            reaped = False
            while self._deleted:
                key = self._deleted.popleft()
                self._kill(key)
                self.alloc.release(key)
                reaped = True
            if reaped:
                self._write_all_slots()
            if getattr(self, "_cm_dirty", False):
                self._cm_dirty = False
                self.refresh_config_volumes()
            self._sync_jobs()
            self._pods_by_owner = None   # snapshot is only valid inside _sync_jobs
            self._schedule()
            self._flush_slots()          # launchers started below read the slot map of their job
            self._sync_pods()
            self._flush_slots()
            metrics.gpu_slots_free.set(self.alloc.free_gpus)
            metrics.ranks_active.set(sum(1 for p in self._procs.values() if p.popen is not None and p.popen.poll() is None))

    def _complain(self, key: str, msg: str) -> None:
        """One warning per object and message, not one per tick."""
        seen = self.__dict__.setdefault("_complaints", {})
        if seen.get(key) != msg:
            seen[key] = msg
            if len(seen) > 4096:
                seen.clear()
            log.warning("%s", msg)

    # ------------------------------------------------------- Job controller --
    def _sync_jobs(self) -> None:
        now = time.time()
        jobs = self.store.list("jobs")
        # one pod snapshot per tick, indexed by controller uid: listing (and deep-copying) every pod once per Job made the
        # loop quadratic (benchmarks/controller_bench.py: 50 jobs x 8 workers)
        self._pods_by_owner = {}
        if jobs:
            for p in self.store.list("pods"):
                for ref in M.meta(p).get("ownerReferences", []) or []:
                    if ref.get("controller"):
                        self._pods_by_owner.setdefault(ref.get("uid"), []).append(p)
        for job in jobs:
            try:
                self._sync_job(job, now)
            except errors.ApiError as e:
                if not (errors.is_not_found(e) or errors.is_conflict(e)):
                    raise
            except Exception as e:  # noqa: BLE001  (a Job object this loop cannot digest must not stop the others)
                self._complain(M.key_of(job), f"cannot sync job {M.key_of(job)}: {type(e).__name__}: {e}")

    def _job_pods(self, job: dict) -> List[dict]:
        snap = getattr(self, "_pods_by_owner", None)
        if snap is not None:
            return [p for p in snap.get(M.meta(job).get("uid"), []) if M.is_controlled_by(p, job)]
        return [p for p in self.store.list("pods", M.namespace_of(job)) if M.is_controlled_by(p, job)]

    @staticmethod
    def _cond(job: dict, ctype: str) -> Optional[dict]:
        for c in job.get("status", {}).get("conditions", []) or []:
            if c["type"] == ctype:
                return c
        return None

    def _finished(self, job: dict) -> bool:
        return any((self._cond(job, t) or {}).get("status") == "True" for t in ("Complete", "Failed"))

    def _set_job_cond(self, st: dict, ctype: str, status: str, reason: str = "", message: str = "") -> None:
        conds = [c for c in st.get("conditions", []) if c["type"] != ctype]
        now = M.now_rfc3339()
        conds.append({"type": ctype, "status": status, "reason": reason, "message": message,
                      "lastProbeTime": now, "lastTransitionTime": now})
        st["conditions"] = conds

    def _sync_job(self, job: dict, now: float) -> None:
        ns, name = M.namespace_of(job), M.name_of(job)
        spec = job.get("spec", {})
        uid = M.meta(job)["uid"]
        # admission defaults: selector + pod labels (what the apiserver/job controller add)
        if "selector" not in spec:
            job = copy.deepcopy(job)
            job["spec"]["selector"] = {"matchLabels": {JOB_UID_LABEL: uid}}
            tl = job["spec"]["template"].setdefault("metadata", {}).setdefault("labels", {})
            tl.update({JOB_UID_LABEL: uid, JOB_NAME_LABEL: name, "controller-uid": uid, "job-name": name})
            job = self.store.update("jobs", job)
            spec = job["spec"]
        st = copy.deepcopy(job.get("status", {}))
        pods = self._job_pods(job)
        if self._finished(job):
            ttl = spec.get("ttlSecondsAfterFinished")
            if ttl is not None:
                done_at = st.get("completionTime") or (self._cond(job, "Failed") or {}).get("lastTransitionTime")
                if done_at and now - M.parse_rfc3339(done_at) >= ttl:
                    log.info("TTL expired for job %s/%s: deleting", ns, name)
                    self.store.delete("jobs", ns, name)
            return
        active = [p for p in pods if p.get("status", {}).get("phase") in (None, "", "Pending", "Running")]
        succeeded = [p for p in pods if p.get("status", {}).get("phase") == "Succeeded"]
        failed = [p for p in pods if p.get("status", {}).get("phase") == "Failed"]
        restarts = sum(int(cs.get("restartCount", 0)) for p in active for cs in p.get("status", {}).get("containerStatuses", []) or [])
        if spec.get("suspend"):
            for p in active:
                self.store.delete("pods", ns, M.name_of(p))
            st["active"] = 0
            st.pop("startTime", None)
            if (self._cond({"status": st}, "Suspended") or {}).get("status") != "True":
                self._set_job_cond(st, "Suspended", "True", "JobSuspended", "Job suspended")
            self._put_job_status(job, st)
            return
        if (self._cond({"status": st}, "Suspended") or {}).get("status") == "True":
            self._set_job_cond(st, "Suspended", "False", "JobResumed", "Job resumed")
        if not st.get("startTime"):
            st["startTime"] = M.now_rfc3339(now)
        st["succeeded"] = len(succeeded)
        st["failed"] = len(failed)
        backoff_limit = spec.get("backoffLimit", DEFAULT_BACKOFF_LIMIT)
        deadline = spec.get("activeDeadlineSeconds")
        if succeeded:
            for p in active:
                self.store.delete("pods", ns, M.name_of(p))
            st["active"] = 0
            st["completionTime"] = M.now_rfc3339(now)
            self._set_job_cond(st, "Complete", "True", "", "")
        elif deadline is not None and now - M.parse_rfc3339(st["startTime"]) >= deadline:
            for p in active:
                self.store.delete("pods", ns, M.name_of(p))
            st["active"] = 0
            self._set_job_cond(st, "Failed", "True", "DeadlineExceeded", "Job was active longer than specified deadline")
        elif len(failed) + restarts > backoff_limit:
            for p in active:
                self.store.delete("pods", ns, M.name_of(p))
            st["active"] = 0
            self._set_job_cond(st, "Failed", "True", "BackoffLimitExceeded", "Job has reached the specified backoff limit")
        else:
            if not active:
                key = f"{ns}/{name}"
                nfail = len(failed)
                ready_at = self._job_backoff.get(key, 0.0)
                if nfail and ready_at == 0.0:
                    ready_at = self._job_backoff[key] = now + min(0.25 * (2 ** (nfail - 1)), 10.0)
                if now >= ready_at:
                    self._job_backoff.pop(key, None)
                    self._create_job_pod(job)
                    st["active"] = 1
                else:
                    st["active"] = 0
            else:
                st["active"] = len(active)
        self._put_job_status(job, st)

    def _put_job_status(self, job: dict, st: dict) -> None:
        if st != job.get("status", {}):
            j = copy.deepcopy(job)
            j["status"] = st
            M.meta(j).pop("resourceVersion", None)
            self.store.update_status("jobs", j)

    def _create_job_pod(self, job: dict) -> None:
        tmpl = copy.deepcopy(job["spec"]["template"])
        md = tmpl.get("metadata", {})
        pod = {
            "apiVersion": "v1", "kind": "Pod",
            "metadata": {"name": f"{M.name_of(job)}-{_rand()}", "namespace": M.namespace_of(job),
                         "labels": copy.deepcopy(md.get("labels") or {}),
                         "ownerReferences": [M.new_controller_ref(job, "batch/v1", "Job")]},
            "spec": tmpl["spec"],
        }
        if md.get("annotations"):
            pod["metadata"]["annotations"] = copy.deepcopy(md["annotations"])
        self.store.create("pods", pod)

    # ------------------------------------------------------------ scheduler --
    @staticmethod
    def _gpu_request(pod: dict) -> int:
        """Sum of the containers' nvidia.com/gpu quantities (limits win over requests). Raises ValueError with the field's
        path for a quantity that is not a whole number: extended resources cannot be fractional (the apiserver says the same)."""
        n = 0
        for k, c in enumerate(pod["spec"].get("containers") or []):
            res = c.get("resources") or {}
            v = (res.get("limits") or {}).get(C.GPU_RESOURCE, (res.get("requests") or {}).get(C.GPU_RESOURCE, 0))
            try:
                q = 0 if v is None else (int(str(v).strip() or 0) if not isinstance(v, bool) else None)
            except (TypeError, ValueError):
                q = None
            if q is None or q < 0:
                raise ValueError(f"spec.containers[{k}].resources: {C.GPU_RESOURCE}: Invalid value: {v!r}: must be a non-negative integer")
            n += q
        return n

    @staticmethod
    def _group_of(pod: dict) -> str:
        md = M.meta(pod)
        g = (md.get("annotations") or {}).get(C.VOLCANO_GROUP_NAME_ANNOTATION) or (md.get("labels") or {}).get(C.SCHED_PLUGINS_POD_GROUP_LABEL)
        return f"{md.get('namespace', '')}/{g}" if g else ""

    def _pod_group(self, key: str) -> Optional[dict]:
        ns, name = M.split_key(key)
        for res in ("volcano-podgroups", "sched-podgroups"):
            try:
                return self.store.get(res, ns, name)
            except errors.ApiError:
                continue
        return None

    def _priority(self, pod: dict, pg: Optional[dict]) -> int:
        name = (pg or {}).get("spec", {}).get("priorityClassName") or pod["spec"].get("priorityClassName")
        if not name:
            return 0
        try:
            return int(self.store.get("priorityclasses", "", name).get("value", 0))
        except errors.ApiError:
            return 0

    def _schedule(self) -> None:
        pending = [p for p in self.store.list("pods")
                   if isinstance(p.get("spec"), dict) and not p["spec"].get("nodeName") and (p.get("status") or {}).get("phase") in (None, "", "Pending")
                   and not M.meta(p).get("deletionTimestamp")]
        if not pending:
            return
        groups: Dict[str, List[dict]] = {}
        singles: List[dict] = []
        for p in pending:
            g = self._group_of(p)
            (groups.setdefault(g, []) if g else singles).append(p)
        order = []
        for g, members in groups.items():
            pg = self._pod_group(g)
            order.append((-self._priority(members[0], pg), min(M.meta(m).get("creationTimestamp", "") for m in members), g, members, pg))
        order.sort(key=lambda t: (t[0], t[1], t[2]))
        now = time.time()
        # one malformed pod or gang (a quantity that is not a number, a PodGroup with a bad minMember) must not keep every
        # other pod on the box from being scheduled: it is marked unschedulable with the reason and the loop goes on
        for _, _, g, members, pg in order:
            try:
                self._schedule_gang(g, members, pg, now)
            except Exception as e:  # noqa: BLE001
                self._complain("gang " + g, f"cannot schedule gang {g}: {e}")
                for p in members:
                    self._mark_unschedulable(p, f"invalid scheduling input: {e}")
        for p in sorted(singles, key=lambda p: M.meta(p).get("creationTimestamp", "")):
            try:
                want = self._gpu_request(p)
                got = self.alloc.allocate(SlotRequest(M.key_of(p), want))
                if got is None:
                    self._mark_unschedulable(p, f"needs {want} GPUs, {self.alloc.free_gpus} free")
                else:
                    self._bind(p, got)
            except Exception as e:  # noqa: BLE001
                self._complain(M.key_of(p), f"cannot schedule pod {M.key_of(p)}: {e}")
                self._mark_unschedulable(p, f"invalid scheduling input: {e}")
        self._write_all_slots()

    def _schedule_gang(self, g: str, members: List[dict], pg: Optional[dict], now: float) -> None:
        all_members = [p for p in self.store.list("pods", M.split_key(g)[0]) if self._group_of(p) == g
                       and p.get("status", {}).get("phase") not in ("Succeeded", "Failed")]
        reqs = [SlotRequest(M.key_of(p), self._gpu_request(p), g) for p in all_members]
        spec = (pg or {}).get("spec", {})
        min_member = int(spec.get("minMember", len(reqs)) or 0)
        min_gpus = int((spec.get("minResources") or {}).get(C.GPU_RESOURCE, 0) or 0)
        feasible = min_gpus <= self.topology.gpu_count if self.topology.gpu_count or min_gpus else True
        grant = self.alloc.allocate_gang(reqs, min_member, min_gpus) if feasible and pg is not None else None
        if grant is None:
            since = self._unsched_since.setdefault(g, now)
            timeout = int(spec.get("scheduleTimeoutSeconds", 0) or 0)
            why = "PodGroup not found" if pg is None else (
                f"gang of {len(reqs)}/{min_member} members needs {max(min_gpus, sum(r.gpus for r in reqs))} GPUs, {self.alloc.free_gpus} free")
            if timeout and now - since > timeout:
                why += f" (scheduleTimeoutSeconds={timeout} exceeded)"
            for p in members:
                self._mark_unschedulable(p, why)
            return
        self._unsched_since.pop(g, None)
        for p in members:
            self._bind(p, grant[M.key_of(p)])

    def _mark_unschedulable(self, pod: dict, why: str) -> None:
        st = copy.deepcopy(pod.get("status", {}))
        st["phase"] = "Pending"
        conds = [c for c in st.get("conditions", []) if c["type"] != "PodScheduled"]
        conds.append({"type": "PodScheduled", "status": "False", "reason": "Unschedulable", "message": why})
        st["conditions"] = conds
        if st != pod.get("status", {}):
            p = copy.deepcopy(pod)
            p["status"] = st
            M.meta(p).pop("resourceVersion", None)
            try:
                self.store.update_status("pods", p)
            except errors.ApiError:
                pass

    def _bind(self, pod: dict, gpus: List[int]) -> None:
        p = copy.deepcopy(pod)
        p["spec"]["nodeName"] = NODE_NAME
        M.meta(p).setdefault("annotations", {})[GPU_ANNOTATION] = ",".join(str(g) for g in gpus)
        M.meta(p).pop("resourceVersion", None)
        try:
            p = self.store.update("pods", p)
            st = copy.deepcopy(p.get("status", {}))
            st["phase"] = "Pending"
            st["conditions"] = [c for c in st.get("conditions", []) if c["type"] != "PodScheduled"] + [
                {"type": "PodScheduled", "status": "True"}]
            p["status"] = st
            M.meta(p).pop("resourceVersion", None)
            self.store.update_status("pods", p)
        except errors.ApiError as e:
            if not errors.is_not_found(e):
                raise
            self.alloc.release(M.key_of(pod))

    # -------------------------------------------------------------- kubelet --
    def _sync_pods(self) -> None:
        for pod in self.store.list("pods"):
            key = M.key_of(pod)
            if not isinstance(pod.get("spec"), dict) or not pod["spec"].get("nodeName"):
                continue
            phase = (pod.get("status") or {}).get("phase")
            if phase in ("Succeeded", "Failed"):
                continue
            try:
                if key not in self._procs:
                    self._start_pod(pod)
                else:
                    self._poll_pod(pod)
            except errors.ApiError as e:
                if not (errors.is_not_found(e) or errors.is_conflict(e)):
                    raise
            except Exception as e:  # noqa: BLE001
                # kubelet's answer to a pod spec it cannot turn into a container: the pod fails with the reason, the box goes on
                self._complain(key, f"cannot run pod {key}: {type(e).__name__}: {e}")
                pr = self._procs.get(key)
                if pr is None or pr.popen is None:
                    pr = self._procs.setdefault(key, _Proc())
                    why = f"{type(e).__name__}: {e}"
                    try:
                        self._set_terminal(pod, pr, "Failed", 128, "CreateContainerConfigError", why)
                    except Exception:  # noqa: BLE001  (not even a container list to report on: the bare phase)
                        bare = copy.deepcopy(pod)
                        bare["status"] = {"phase": "Failed", "reason": "CreateContainerConfigError", "message": why}
                        M.meta(bare).pop("resourceVersion", None)
                        try:
                            self.store.update_status("pods", bare)
                        except errors.ApiError:
                            pass
                        self.alloc.release(key)

    def pod_dir(self, pod: dict) -> str:
        return os.path.join(self.state_dir, "pods", M.namespace_of(pod), M.name_of(pod))

    def _materialize_volumes(self, pod: dict, pdir: str) -> Dict[str, str]:
        """Returns mountPath -> real directory for container[0]."""
        vols = {v["name"]: v for v in pod["spec"].get("volumes", []) or []}
        mounts: Dict[str, str] = {}
        c0 = pod["spec"]["containers"][0]
        ns = M.namespace_of(pod)
        for m in c0.get("volumeMounts", []) or []:
            v = vols.get(m["name"])
            if v is None:
                continue
            real = os.path.join(pdir, "rootfs", m["mountPath"].lstrip("/"))
            os.makedirs(real, exist_ok=True)
            try:
                if "configMap" in v:
                    cm = self.store.get("configmaps", ns, v["configMap"]["name"])
                    items = v["configMap"].get("items") or [{"key": k, "path": k} for k in cm.get("data", {})]
                    for it in items:
                        if it["key"] in cm.get("data", {}):
                            self._write_file(os.path.join(real, it["path"]), cm["data"][it["key"]].encode(), it.get("mode", v["configMap"].get("defaultMode", 0o644)))
                elif "secret" in v:
                    sec = self.store.get("secrets", ns, v["secret"]["secretName"])
                    items = v["secret"].get("items") or [{"key": k, "path": k} for k in sec.get("data", {})]
                    for it in items:
                        if it["key"] in sec.get("data", {}):
                            self._write_file(os.path.join(real, it["path"]), base64.b64decode(sec["data"][it["key"]]), it.get("mode", v["secret"].get("defaultMode", 0o644)))
            except errors.ApiError:
                pass  # kubelet would keep the container in ContainerCreating; the next sync retries content
            mounts[m["mountPath"]] = real
        return mounts

    @staticmethod
    def _write_file(path: str, data: bytes, mode: int) -> None:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(data)
        os.chmod(tmp, mode | 0o200)  # keep owner-writable so refreshes can replace it
        os.replace(tmp, path)

    def _job_slots_path(self, ns: str, job_name: str) -> str:
        return os.path.join(self.state_dir, "jobs", ns, job_name, "slots.json")

    def _write_all_slots(self) -> None:
        """Mark the per-job slot maps stale; they are rewritten once per tick (before pods are started and at the end of the
        tick) instead of once per pod event — each rewrite lists every pod."""
        self._slots_dirty = True

    def _flush_slots(self) -> None:
        if getattr(self, "_slots_dirty", False):
            self._slots_dirty = False
            self._write_all_slots_now()

    def _write_all_slots_now(self) -> None:
        """hostname -> GPU list per MPIJob, read by the native mpirun to pin ranks."""
        by_job: Dict[tuple, Dict[str, List[int]]] = {}
        for pod in self.store.list("pods"):
            labels = M.meta(pod).get("labels") or {}
            jn = labels.get(C.JOB_NAME_LABEL)
            if not jn:
                continue
            ann = (M.meta(pod).get("annotations") or {}).get(GPU_ANNOTATION)
            if ann is None or pod.get("status", {}).get("phase") in ("Succeeded", "Failed"):
                continue
            host = pod["spec"].get("hostname") or M.name_of(pod)
            by_job.setdefault((M.namespace_of(pod), jn), {})[host] = [int(x) for x in ann.split(",") if x != ""]
        for (ns, jn), hosts in by_job.items():
            path = self._job_slots_path(ns, jn)
            data = json.dumps({"hosts": hosts}, sort_keys=True).encode()
            try:
                with open(path, "rb") as f:
                    if f.read() == data:
                        continue
            except OSError:
                pass
            self._write_file(path, data, 0o644)

    def _build_env(self, pod: dict, pdir: str, mounts: Dict[str, str]) -> Dict[str, str]:
        c0 = pod["spec"]["containers"][0]
        env = dict(os.environ)
        env["PATH"] = BIN_DIR + os.pathsep + env.get("PATH", "")
        env.update(self.extra_env)
        labels = M.meta(pod).get("labels") or {}
        for e in c0.get("env", []) or []:
            env[e["name"]] = str(e.get("value", "") or "")
        # mount-path rewriting: /etc/mpi/hostfile -> <pod rootfs>/etc/mpi/hostfile
        for k, v in list(env.items()):
            for mp, real in mounts.items():
                if isinstance(v, str) and v.startswith(mp.rstrip("/") + "/"):
                    env[k] = real + v[len(mp.rstrip("/")):]
        env["HOSTNAME"] = pod["spec"].get("hostname") or M.name_of(pod)
        env["B200MPI_POD_NAME"] = M.name_of(pod)
        env["B200MPI_POD_NAMESPACE"] = M.namespace_of(pod)
        env["B200MPI_POD_DIR"] = pdir
        env["B200MPI_POD_ROOTFS"] = os.path.join(pdir, "rootfs")
        env.setdefault("B200MPI_STATS_DIR", os.path.join(pdir, "stats"))  # ranks drop collective counters here on exit
        env["PYTHONPATH"] = os.path.dirname(PKG_DIR) + os.pathsep + env.get("PYTHONPATH", "")
        jn = labels.get(C.JOB_NAME_LABEL)
        if jn:
            env["B200MPI_MPIJOB_NAME"] = jn
            env["B200MPI_SLOTS_FILE"] = self._job_slots_path(M.namespace_of(pod), jn)
            env.setdefault("B200MPI_JOB_ID", f"{M.namespace_of(pod)}.{jn}.{M.meta(pod).get('uid', '')[:8]}")
        # LD-inject the collective runtime into every rank mpirun spawns (north star): unmodified
        # torch.distributed scripts then resolve ncclAllReduce & co. to b200mpi kernels. B200MPI_ALGO=nccl
        # (or B200MPI_INJECT=0) is the baseline mode: same launcher, stock NCCL.
        shim = os.path.join(PKG_DIR, "lib", "libb200mpi_nccl.so")
        # On by default (round 2: unmodified torch DDP under the shim passes at 2, 4 and 8 GPUs, tests/test_multigpu.py);
        # B200MPI_INJECT=0 or B200MPI_ALGO=nccl in the launcher env is the baseline mode (same launcher, stock NCCL).
        # A box without GPUs has nothing to inject into (CPU MPIJobs such as examples/pi): there the default is off.
        default_inject = "1" if self.topology.gpu_count > 0 else "0"
        if (os.path.exists(shim) and env.get("B200MPI_ALGO", "") != "nccl" and env.get("B200MPI_INJECT", default_inject) != "0"
                and "B200MPI_INJECT_LIB" not in env):
            env["B200MPI_INJECT_LIB"] = shim
        gpus = (M.meta(pod).get("annotations") or {}).get(GPU_ANNOTATION, "")
        env["B200MPI_GPUS"] = gpus
        if env.get("NVIDIA_VISIBLE_DEVICES", None) == "" and "NVIDIA_VISIBLE_DEVICES" in {e["name"] for e in c0.get("env", []) or []}:
            env["CUDA_VISIBLE_DEVICES"] = ""  # launcher kept off the GPUs (controller.go:1600-1606)
        elif gpus:
            env["CUDA_VISIBLE_DEVICES"] = gpus
        return env

    def _start_pod(self, pod: dict) -> None:
        key = M.key_of(pod)
        pdir = self.pod_dir(pod)
        pr = _Proc()
        pr.pod_dir = pdir
        pr.log_path = os.path.join(pdir, "logs", "0.log")
        c0 = pod["spec"]["containers"][0]
        img = image_config(c0.get("image", ""))
        argv = list(c0.get("command") or []) + list(c0.get("args") or [])
        if not c0.get("command") and c0.get("args") and img.get("entrypoint") and os.path.basename(argv[0]) != "sshd":
            argv = list(img["entrypoint"]) + argv
        self._procs[key] = pr
        if argv and os.path.basename(argv[0]) == "sshd":
            pr.virtual = True   # idle worker placeholder: nothing runs in it, so nothing to mount either
            pr.started_at = M.now_rfc3339()
            self._set_running(pod, pr)
            return
        os.makedirs(os.path.join(pdir, "logs"), exist_ok=True)
        mounts = self._materialize_volumes(pod, pdir)
        if not argv:
            self._set_terminal(pod, pr, "Failed", 128, "ContainerCannotRun",
                               "container has no command/args and images are not used on a single box")
            return
        env = self._build_env(pod, pdir, mounts)
        argv = [img.get("remap", {}).get(a, a) for a in argv]
        if argv[0] in ("python", "python3"):
            import sys as _sys
            argv[0] = _sys.executable
        env["B200MPI_PYTHON"] = __import__("sys").executable
        self._launch(pod, pr, argv, env, c0.get("workingDir") or img.get("workingDir"))

    def _launch(self, pod: dict, pr: _Proc, argv: List[str], env: Dict[str, str], cwd: Optional[str]) -> None:
        logf = open(pr.log_path, "ab")
        try:
            pr.popen = subprocess.Popen(argv, env=env, cwd=cwd or None, stdout=logf, stderr=subprocess.STDOUT,
                                        stdin=subprocess.DEVNULL, start_new_session=True)
        except OSError as e:
            logf.write(f"exec failed: {e}\n".encode())
            logf.close()
            pr.popen = None
            self._set_terminal(pod, pr, "Failed", 127, "ContainerCannotRun", f"exec {argv[0]!r}: {e}")
            return
        logf.close()
        pr._argv, pr._env, pr._cwd = argv, env, cwd  # for OnFailure restarts
        pr.started_at = M.now_rfc3339()
        try:  # remember the process group so a restarted daemon can reap it
            cur = self.store.get("pods", M.namespace_of(pod), M.name_of(pod))
            M.meta(cur).setdefault("annotations", {})[PGID_ANNOTATION] = str(pr.popen.pid)
            pod = self.store.update("pods", cur)
        except errors.ApiError:
            pass
        self._set_running(pod, pr)

    def _container_status(self, pod: dict, pr: _Proc, state: dict, ready: bool) -> List[dict]:
        return [{"name": pod["spec"]["containers"][0].get("name", "main"), "state": state, "ready": ready,
                 "restartCount": pr.restarts, "started": "running" in state}]

    def _set_running(self, pod: dict, pr: _Proc) -> None:
        p = copy.deepcopy(pod)
        st = p.setdefault("status", {})
        st["phase"] = "Running"
        st.setdefault("startTime", pr.started_at)
        st["hostIP"] = st["podIP"] = "127.0.0.1"
        conds = [c for c in st.get("conditions", []) if c["type"] not in ("Ready", "ContainersReady", "Initialized")]
        conds += [{"type": "Initialized", "status": "True"}, {"type": "ContainersReady", "status": "True"}, {"type": "Ready", "status": "True"}]
        st["conditions"] = conds
        st["containerStatuses"] = self._container_status(pod, pr, {"running": {"startedAt": pr.started_at}}, True)
        M.meta(p).pop("resourceVersion", None)
        self.store.update_status("pods", p)

    def _set_terminal(self, pod: dict, pr: _Proc, phase: str, code: int, reason: str, message: str = "") -> None:
        p = copy.deepcopy(pod)
        st = p.setdefault("status", {})
        st["phase"] = phase
        if phase == "Failed":
            st["reason"], st["message"] = reason, message
        conds = [c for c in st.get("conditions", []) if c["type"] not in ("Ready", "ContainersReady")]
        conds += [{"type": "ContainersReady", "status": "False", "reason": "PodCompleted" if phase == "Succeeded" else "PodFailed"},
                  {"type": "Ready", "status": "False", "reason": "PodCompleted" if phase == "Succeeded" else "PodFailed"}]
        st["conditions"] = conds
        st["containerStatuses"] = self._container_status(pod, pr, {"terminated": {
            "exitCode": code, "reason": reason, "message": message, "startedAt": pr.started_at, "finishedAt": M.now_rfc3339()}}, False)
        M.meta(p).pop("resourceVersion", None)
        self.store.update_status("pods", p)
        self.alloc.release(M.key_of(pod))
        self._write_all_slots()

    @staticmethod
    def _sweep_shm(env) -> None:
        """The container's processes are gone (exited, or killed together with their mpirun, which then could not clean up):
        drop the rendezvous / data segments that carry its job id, like kubelet tearing down a pod's emptyDir."""
        job = (env or {}).get("B200MPI_JOB_ID", "")
        if not job:
            return
        key = "".join(ch if (ch.isalnum() or ch in "-_.") else "_" for ch in job)
        try:
            for n in os.listdir("/dev/shm"):
                # b200mpi-[mpi-]<job id>.<mpirun pid>[-suffix]: match the id up to a separator, not as a substring
                rest = n[len("b200mpi-mpi-"):] if n.startswith("b200mpi-mpi-") else (n[len("b200mpi-"):] if n.startswith("b200mpi-") else None)
                if rest is not None and (rest == key or rest.startswith(key + ".") or rest.startswith(key + "-")):
                    try:
                        os.unlink(os.path.join("/dev/shm", n))
                    except OSError:
                        pass
        except OSError:
            pass

    def _harvest_stats(self, pr: _Proc) -> None:
        """Per-rank collective counters (runtime/comm.py: dump_stats) -> b200mpi_collective_*_total. Files are consumed
        so an OnFailure restart of the same container is not counted twice."""
        d = os.path.join(pr.pod_dir, "stats")
        try:
            names = [n for n in os.listdir(d) if n.startswith("stats-") and n.endswith(".json")]
        except OSError:
            return
        from ..controller import metrics
        for n in names:
            path = os.path.join(d, n)
            try:
                with open(path) as f:
                    metrics.observe_rank_stats(json.load(f))
                os.replace(path, path + ".seen")
            except (OSError, ValueError):
                continue

    def _tail(self, path: str, n: int = 512) -> str:
        try:
            with open(path, "rb") as f:
                f.seek(0, 2)
                size = f.tell()
                f.seek(max(0, size - n))
                return f.read().decode(errors="replace").strip()
        except OSError:
            return ""

    def _poll_pod(self, pod: dict) -> None:
        pr = self._procs[M.key_of(pod)]
        if pr.virtual:
            return
        if pr.popen is None:
            if pr.next_restart and time.time() >= pr.next_restart:
                pr.next_restart = 0.0
                self._launch(pod, pr, pr._argv, pr._env, pr._cwd)
            return
        rc = pr.popen.poll()
        if rc is None:
            return
        pr.popen = None
        self._harvest_stats(pr)
        self._sweep_shm(getattr(pr, "_env", None))
        if rc == 0:
            self._set_terminal(pod, pr, "Succeeded", 0, "Completed")
            return
        policy = pod["spec"].get("restartPolicy", "Never")
        code = rc if rc > 0 else 128 - rc
        msg = self._tail(pr.log_path) or f"exit code {code}"
        if policy == "OnFailure":
            pr.restarts += 1
            pr.next_restart = time.time() + min(0.2 * (2 ** (pr.restarts - 1)), 10.0)
            p = copy.deepcopy(pod)
            st = p.setdefault("status", {})
            st["containerStatuses"] = self._container_status(pod, pr, {"waiting": {"reason": "CrashLoopBackOff", "message": msg}}, False)
            st["containerStatuses"][0]["lastState"] = {"terminated": {"exitCode": code, "reason": "Error", "message": msg}}
            M.meta(p).pop("resourceVersion", None)
            self.store.update_status("pods", p)
        else:
            self._set_terminal(pod, pr, "Failed", code, "Error", msg)

    def _kill(self, key: str, grace: float = 2.0) -> None:
        pr = self._procs.pop(key, None)
        if pr is None or pr.popen is None:
            return
        try:
            pgid = os.getpgid(pr.popen.pid)
            os.killpg(pgid, signal.SIGTERM)
            t0 = time.time()
            while pr.popen.poll() is None and time.time() - t0 < grace:
                time.sleep(0.02)
            if pr.popen.poll() is None:
                os.killpg(pgid, signal.SIGKILL)
                pr.popen.wait(timeout=2)
        except (ProcessLookupError, PermissionError, subprocess.TimeoutExpired):
            pass
        self._sweep_shm(getattr(pr, "_env", None))     # the ranks were killed along with their mpirun: nobody else cleans up

    # ----------------------------------------------------- volume refresh --
    def refresh_config_volumes(self) -> None:
        """ConfigMap volume propagation (discover_hosts.sh refresh for elastic Horovod)."""
        with self._lock:
            for key in list(self._procs):
                ns, name = M.split_key(key)
                try:
                    pod = self.store.get("pods", ns, name)
                except errors.ApiError:
                    continue
                self._materialize_volumes(pod, self._procs[key].pod_dir)

    def logs(self, namespace: str, pod_name: str) -> str:
        path = os.path.join(self.state_dir, "pods", namespace, pod_name, "logs", "0.log")
        try:
            with open(path, "r", errors="replace") as f:
                return f.read()
        except OSError:
            return ""

    def log_slice(self, namespace: str, pod_name: str, offset: int) -> "tuple[bytes, int]":
        """Bytes of the pod's log from ``offset`` on and the offset to ask with next time (what `logs -f` polls with: the cost
        of a poll is the new data, not the whole file). A log that became SHORTER than the offset was restarted or rotated:
        reading starts over at 0."""
        path = os.path.join(self.state_dir, "pods", namespace, pod_name, "logs", "0.log")
        try:
            size = os.path.getsize(path)
            if size < offset:
                offset = 0
            if size == offset:
                return b"", offset
            with open(path, "rb") as f:
                f.seek(offset)
                data = f.read(4 << 20)          # bounded pieces: a follower of a huge log streams it in 4 MiB steps
            return data, offset + len(data)
        except OSError:
            return b"", 0

    def rotate_logs(self) -> int:
        """kubelet-style container log rotation (containerLogMaxSize; here B200MPI_POD_LOG_MAX_BYTES, default 256 MiB, 0 = off):
        a running pod's log that outgrew the limit is copied to `0.log.1` (replacing the previous generation) and truncated in
        place - the container holds it open in append mode, so it keeps writing at the new end. Returns the number of rotations."""
        limit = int(os.environ.get("B200MPI_POD_LOG_MAX_BYTES", 256 << 20))
        if limit <= 0:
            return 0
        n = 0
        with self._lock:
            paths = [pr.log_path for pr in self._procs.values() if pr.log_path and pr.popen is not None and pr.popen.poll() is None]
        for path in paths:
            try:
                if os.path.getsize(path) <= limit:
                    continue
                import shutil
                shutil.copyfile(path, path + ".1")
                with open(path, "r+b") as f:
                    f.truncate(0)
                n += 1
            except OSError:
                continue
        return n
