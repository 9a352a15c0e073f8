"""MPIJob controller: level-triggered reconciliation of MPIJob objects.

Behavioural port (not a translation) of the reference reconciler —
pkg/controller/mpi_job_controller.go: wiring :223-459, run loop :465-562,
``syncHandler`` :567-735, get-or-create :752-1036, cleanup :737-749/:1046-1086,
status :1088-1207, event handlers :1210-1298.  The objects it manipulates live
in the local ``ObjectStore``; the node agent (``mpi_operator_b200.node``) plays
kube-scheduler + kubelet + the batch/v1 Job controller and turns Pods into
processes pinned to GPU slots.
"""
from __future__ import annotations

import copy
import logging
import threading
from typing import Callable, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..api.register import scheme
from ..api.types import MPIJob
from ..api.validation import validate_mpijob
from ..client import errors
from ..client.clientset import Clientset, KubeClient
from ..client.informers import SharedInformerFactory
from . import builders as B
from . import metrics
from . import status as S
from .clock import RealClock
from .events import EVENT_TYPE_NORMAL, EVENT_TYPE_WARNING, EventRecorder
from .podgroup import PodGroupControl, SchedulerPluginsCtrl, VolcanoCtrl
from .workqueue import RateLimitingQueue, default_controller_rate_limiter

log = logging.getLogger("mpi-job-controller")

ERR_RESOURCE_EXISTS = "ErrResourceExists"
MESSAGE_RESOURCE_EXISTS = 'Resource "%s" of Kind "%s" already exists and is not managed by MPIJob'
VALIDATION_ERROR = "ValidationError"
EVENT_MESSAGE_LIMIT = 1024
JOB_REASON_BACKOFF_LIMIT_EXCEEDED = "BackoffLimitExceeded"


class SyncError(Exception):
    """An error returned by syncHandler => the key is re-queued with back-off."""


def truncate_message(message: str) -> str:
    """controller.go:1801-1808."""
    if len(message) <= EVENT_MESSAGE_LIMIT:
        return message
    suffix = "..."
    return message[:EVENT_MESSAGE_LIMIT - len(suffix)] + suffix


def managed_by_external_controller(name: Optional[str]) -> Optional[str]:
    """controller.go:1810-1815."""
    if name is not None and name != C.KUBEFLOW_JOB_CONTROLLER:
        return name
    return None


def is_mpijob_suspended(job: MPIJob) -> bool:
    return bool(job.spec.run_policy.suspend)


def is_job_suspended(job: dict) -> bool:
    return bool(job.get("spec", {}).get("suspend", False))


def get_job_condition(job: dict, ctype: str) -> Optional[dict]:
    for c in job.get("status", {}).get("conditions", []) or []:
        if c.get("type") == ctype:
            return c
    return None


def is_job_failed(job: dict) -> bool:
    c = get_job_condition(job, "Failed")
    return c is not None and c.get("status") == "True"


def is_job_succeeded(job: dict) -> bool:
    c = get_job_condition(job, "Complete")
    return c is not None and c.get("status") == "True"


def is_job_finished(job: dict) -> bool:
    return is_job_succeeded(job) or is_job_failed(job)


def pod_phase(pod: dict) -> str:
    return pod.get("status", {}).get("phase", "")


def is_pod_running(p): return pod_phase(p) == "Running"  # noqa: E704
def is_pod_pending(p): return pod_phase(p) == "Pending"  # noqa: E704
def is_pod_failed(p): return pod_phase(p) == "Failed"  # noqa: E704


def is_clean_up_pods(policy: Optional[str]) -> bool:
    return policy in (C.CLEAN_POD_POLICY_ALL, C.CLEAN_POD_POLICY_RUNNING)


def count_running_pods(pods: List[dict]) -> int:
    return sum(1 for p in pods if is_pod_running(p))


def count_ready_worker_pods(workers: List[dict]) -> int:
    n = 0
    for p in workers:
        for c in p.get("status", {}).get("conditions", []) or []:
            if c.get("type") == "Ready" and c.get("status") == "True":
                n += 1
                break
    return n


class MPIJobController:
    """controller.go:223-265."""

    def __init__(self, kube: KubeClient, kubeflow: Clientset, informers: SharedInformerFactory, *,
                 gang_scheduling: str = "", cluster_domain: str = "", recorder: Optional[EventRecorder] = None,
                 clock=None, queue_rate_limit: float = 10.0, queue_burst: int = 100, namespace: str = ""):
        self.kube, self.kubeflow = kube, kubeflow
        self.informers = informers
        self.cluster_domain = cluster_domain
        self.clock = clock or RealClock()
        self.recorder = recorder or EventRecorder(kube.store, B.CONTROLLER_AGENT_NAME)
        self.namespace = namespace
        f = informers
        self.config_map_lister = f.lister_for("configmaps")
        self.secret_lister = f.lister_for("secrets")
        self.service_lister = f.lister_for("services")
        self.job_lister = f.lister_for("jobs")
        self.pod_lister = f.lister_for("pods")
        self.priority_class_lister = f.lister_for("priorityclasses")
        self.mpijob_lister = f.lister_for("mpijobs")
        self.pod_group_ctrl: Optional[PodGroupControl] = None
        if gang_scheduling == C.GANG_SCHEDULER_VOLCANO:  # controller.go:319-327
            self.pod_group_ctrl = VolcanoCtrl(kube, f, self.priority_class_lister)
        elif gang_scheduling:
            self.pod_group_ctrl = SchedulerPluginsCtrl(kube, f, gang_scheduling, self.priority_class_lister)
        self.queue = RateLimitingQueue(default_controller_rate_limiter(queue_rate_limit, queue_burst), "MPIJob")
        # test seam (controller.go:257-258)
        self.update_status_handler: Callable[[MPIJob], None] = self.do_update_job_status
        self._threads: List[threading.Thread] = []
        self._stop = threading.Event()
        self._register_handlers()

    # ------------------------------------------------------------- wiring --
    def _register_handlers(self) -> None:
        """controller.go:392-457."""
        f = self.informers
        f.informer_for("mpijobs").add_event_handler(
            add=self.add_mpijob, update=lambda old, new: self.enqueue_mpijob(new))
        for res in ("configmaps", "secrets", "services", "jobs", "pods"):
            f.informer_for(res).add_event_handler(add=self.handle_object, update=self.handle_object_update,
                                                  delete=self.handle_object)
        if self.pod_group_ctrl is not None:
            self.pod_group_ctrl.informer.add_event_handler(add=self.handle_object, update=self.handle_object_update,
                                                           delete=self.handle_object)
            f.informer_for("priorityclasses")

    def add_mpijob(self, obj: dict) -> None:
        """controller.go:1210-1216 (defaulting happens again in syncHandler)."""
        self.enqueue_mpijob(obj)

    def enqueue_mpijob(self, obj) -> None:
        key = obj.key if isinstance(obj, MPIJob) else M.key_of(obj)
        self.queue.add_rate_limited(key)

    def handle_object(self, obj: dict) -> None:
        """Owner resolution incl. the Pod -> Job -> MPIJob hop (controller.go:1236-1286)."""
        ref = M.get_controller_of(obj)
        if ref is None:
            return
        group, _, version = ref.get("apiVersion", "").partition("/")
        kind = ref.get("kind", "")
        if group == "batch" and kind == "Job":
            try:
                j = self.job_lister.namespaced(M.namespace_of(obj)).get(ref["name"])
            except errors.ApiError as e:
                log.debug("obtaining owning k8s Job: %s", e)
                return
            ref = M.get_controller_of(j)
            if ref is None:
                return
            group, _, version = ref.get("apiVersion", "").partition("/")
            kind = ref.get("kind", "")
        if kind != C.KIND or group != C.GROUP_NAME or version != C.GROUP_VERSION:
            return
        try:
            job = self.mpijob_lister.mpijobs(M.namespace_of(obj)).get(ref["name"])
        except errors.ApiError:
            log.debug("ignoring orphaned object '%s' of mpi job '%s'", M.key_of(obj), ref.get("name"))
            return
        self.enqueue_mpijob(job)

    def handle_object_update(self, old: dict, new: dict) -> None:
        if M.meta(new).get("resourceVersion") == M.meta(old).get("resourceVersion"):
            return
        self.handle_object(new)

    # ----------------------------------------------------------- run loop --
    def run(self, threadiness: int = 2) -> None:
        """controller.go:465-500 (non-blocking: workers run on daemon threads)."""
        log.info("Starting MPIJob controller")
        self.informers.start()
        if not self.informers.wait_for_cache_sync():
            raise RuntimeError("failed to wait for caches to sync")
        log.info("Starting workers")
        for i in range(threadiness):
            t = threading.Thread(target=self.run_worker, name=f"mpijob-worker-{i}", daemon=True)
            t.start()
            self._threads.append(t)
        log.info("Started workers")

    def stop(self) -> None:
        self._stop.set()
        self.queue.shut_down()
        for t in self._threads:
            t.join(timeout=2)
        log.info("Shutting down workers")

    def run_worker(self) -> None:
        while self.process_next_work_item():
            pass

    def process_next_work_item(self, timeout: Optional[float] = None) -> bool:
        """controller.go:512-562."""
        key, shutdown = self.queue.get(timeout)
        if shutdown:
            return False
        if key is None:
            return True
        try:
            if not isinstance(key, str):
                self.queue.forget(key)
                log.error("expected string in workqueue but got %r", key)
                return True
            try:
                self.sync_handler(key)
            except Exception as e:  # noqa: BLE001 - any error requeues with back-off
                self.queue.add_rate_limited(key)
                log.warning("error syncing '%s': %s", key, e)
                return True
            self.queue.forget(key)
            return True
        finally:
            self.queue.done(key)

    # ------------------------------------------------------------ reconcile --
    def sync_handler(self, key: str) -> None:
        """controller.go:567-735."""
        start = self.clock.now()
        try:
            self._sync(key)
        finally:
            d = self.clock.since(start)
            metrics.reconcile_seconds.observe(max(d, 0.0))
            log.info('Finished syncing job "%s" (%.3fms)', key, d * 1e3)

    def _sync(self, key: str) -> None:
        try:
            namespace, name = M.split_key(key)
        except ValueError:
            log.error("invalid resource key: %s", key)
            return
        try:
            shared = self.mpijob_lister.mpijobs(namespace).get(name)
        except errors.ApiError as e:
            if errors.is_not_found(e):
                log.debug("MPIJob has been deleted: %s", key)
                return
            raise SyncError(f"obtaining job: {e}")
        job: MPIJob = shared.deepcopy()  # never mutate the cache
        scheme.default(job)
        if managed_by_external_controller(job.spec.run_policy.managed_by) is not None:
            log.info("Skipping MPIJob managed by a custom controller managed-by=%s", job.spec.run_policy.managed_by)
            return
        if job.deletion_timestamp is not None:
            return
        errs = validate_mpijob(job)
        if errs:
            msg = truncate_message(f"Found validation errors: {errs.to_aggregate()}")
            self.recorder.event(job, EVENT_TYPE_WARNING, VALIDATION_ERROR, msg)
            return  # do not requeue
        if len(job.status.conditions) == 0:
            msg = f"MPIJob {job.namespace}/{job.name} is created."
            S.update_mpijob_conditions(job, C.JOB_CREATED, C.CONDITION_TRUE, S.MPIJOB_CREATED_REASON, msg, self._now())
            self.recorder.event(job, EVENT_TYPE_NORMAL, "MPIJobCreated", msg)
            metrics.mpi_jobs_created.inc()

        # CompletionTime is only filled when the launcher Job succeeded or stopped retrying.
        if S.is_finished(job.status) and job.status.completion_time is not None:
            if is_clean_up_pods(job.spec.run_policy.clean_pod_policy):
                self.clean_up_worker_pods(job)
                self.update_status_handler(job)
            return

        if job.status.start_time is None and not is_mpijob_suspended(job):
            job.status.start_time = self._now()

        launcher = self.get_launcher_job(job)
        worker: List[dict] = []
        done = launcher is not None and is_job_finished(launcher)
        if not done:
            try:
                self.get_or_create_service(job, B.new_job_service(job))
            except errors.ApiError as e:
                raise SyncError(f"getting or creating Service to front workers: {e}")
            self.get_or_create_config_map(job)
            self.get_or_create_ssh_auth_secret(job)
            if not is_mpijob_suspended(job):
                if self.pod_group_ctrl is not None:
                    self.get_or_create_pod_groups(job)
                worker = self.get_or_create_worker(job)
            if launcher is None:
                if (job.spec.launcher_creation_policy == C.LAUNCHER_CREATION_POLICY_AT_STARTUP
                        or count_ready_worker_pods(worker) == len(worker)):
                    try:
                        launcher = self.kube.jobs(namespace).create(
                            B.new_launcher_job(job, self.pod_group_ctrl, self.recorder))
                    except errors.ApiError as e:
                        self.recorder.eventf(job, EVENT_TYPE_WARNING, S.MPIJOB_FAILED_REASON, "launcher pod created failed: %s", e)
                        raise SyncError(f"creating launcher Pod: {e}")
                else:
                    log.debug("Waiting for workers %s/%s to start.", job.namespace, job.name)

        if launcher is not None:
            if not is_mpijob_suspended(job) and is_job_suspended(launcher):
                # resuming: the Job template is immutable once StartTime is set -> clear it first
                if launcher.get("status", {}).get("startTime") is not None:
                    launcher["status"]["startTime"] = None
                    launcher["status"].pop("startTime")
                    launcher = self.kube.jobs(namespace).update_status(launcher)
                desired = B.new_launcher_pod_template(job, self.pod_group_ctrl, self.recorder)
                B.sync_launcher_scheduling_directives(launcher, desired)
                launcher["spec"]["suspend"] = False
                launcher = self.kube.jobs(namespace).update(launcher)
            elif is_mpijob_suspended(job) and not is_job_suspended(launcher):
                launcher["spec"]["suspend"] = True
                launcher = self.kube.jobs(namespace).update(launcher)

        if is_mpijob_suspended(job):
            self.clean_up_worker_pods(job)

        self.update_mpijob_status(job, launcher, worker)

    def _now(self) -> str:
        return M.now_rfc3339(self.clock.now())

    # ----------------------------------------------------------- cleanup --
    def clean_up_worker_pods(self, job: MPIJob) -> None:
        """controller.go:737-749."""
        self.delete_worker_pods(job)
        S.initialize_replica_statuses(job, C.REPLICA_TYPE_WORKER)
        if self.pod_group_ctrl is not None:
            self.delete_pod_groups(job)
        job.status.replica_statuses[C.REPLICA_TYPE_WORKER].active = 0

    def delete_worker_pods(self, job: MPIJob) -> None:
        """controller.go:1046-1086 (CleanPodPolicy=Running keeps finished pods, deletes Running AND Pending)."""
        wspec = job.spec.replica(C.REPLICA_TYPE_WORKER)
        if wspec is None:
            return
        for i in range(int(wspec.replicas or 0)):
            name = B.worker_name(job, i)
            try:
                pod = self.pod_lister.namespaced(job.namespace).get(name)
            except errors.ApiError as e:
                if errors.is_not_found(e):
                    continue
                raise
            self._must_own(job, pod, "Pod")
            if (job.spec.run_policy.clean_pod_policy == C.CLEAN_POD_POLICY_RUNNING
                    and not is_pod_running(pod) and not is_pod_pending(pod)):
                continue
            try:
                self.kube.pods(job.namespace).delete(name)
            except errors.ApiError as e:
                if not errors.is_not_found(e):
                    log.error("Failed to delete pod[%s/%s]: %s", job.namespace, name, e)
                    raise

    def _must_own(self, job: MPIJob, obj: dict, kind: str) -> None:
        if not M.is_controlled_by(obj, job.to_dict()):
            msg = MESSAGE_RESOURCE_EXISTS % (M.name_of(obj), obj.get("kind", kind))
            self.recorder.event(job, EVENT_TYPE_WARNING, ERR_RESOURCE_EXISTS, msg)
            raise SyncError(msg)

    # ------------------------------------------------------ get-or-create --
    def get_launcher_job(self, job: MPIJob) -> Optional[dict]:
        """controller.go:752-773."""
        try:
            launcher = self.job_lister.namespaced(job.namespace).get(B.launcher_name(job))
        except errors.ApiError as e:
            if errors.is_not_found(e):
                return None
            raise
        self._must_own(job, launcher, "Job")
        return launcher

    def get_or_create_pod_groups(self, job: MPIJob) -> dict:
        """controller.go:776-801."""
        new_pg = self.pod_group_ctrl.new_pod_group(job)
        try:
            pg = self.pod_group_ctrl.get_pod_group(M.namespace_of(new_pg), M.name_of(new_pg))
        except errors.ApiError as e:
            if errors.is_not_found(e):
                return self.pod_group_ctrl.create_pod_group(new_pg)
            raise
        self._must_own(job, pg, "PodGroup")
        if not self.pod_group_ctrl.pg_specs_are_equal(pg, new_pg):
            return self.pod_group_ctrl.update_pod_group(pg, new_pg)
        return pg

    def delete_pod_groups(self, job: MPIJob) -> None:
        """controller.go:804-831."""
        try:
            pg = self.pod_group_ctrl.get_pod_group(job.namespace, job.name)
        except errors.ApiError as e:
            if errors.is_not_found(e):
                return
            raise
        self._must_own(job, pg, "PodGroup")
        self.pod_group_ctrl.delete_pod_group(job.namespace, job.name)

    def get_running_worker_pods(self, job: MPIJob) -> List[dict]:
        """controller.go:834-852: only Running pods enter discover_hosts.sh."""
        pods = self.pod_lister.namespaced(job.namespace).list(B.worker_selector(job.name))
        return [p for p in pods if is_pod_running(p)]

    def get_or_create_config_map(self, job: MPIJob) -> dict:
        """controller.go:869-905."""
        new_cm = B.new_config_map(job, job.worker_replicas(), self.cluster_domain)
        B.update_discover_hosts_in_config_map(new_cm, job, self.get_running_worker_pods(job), self.cluster_domain)
        try:
            cm = self.config_map_lister.namespaced(job.namespace).get(job.name + B.CONFIG_SUFFIX)
        except errors.ApiError as e:
            if errors.is_not_found(e):
                try:
                    return self.kube.config_maps(job.namespace).create(new_cm)
                except errors.ApiError as e2:
                    raise SyncError(f"getting or creating ConfigMap: {e2}")
            raise
        self._must_own(job, cm, "ConfigMap")
        if cm.get("data") != new_cm["data"]:
            cm = copy.deepcopy(cm)
            cm["data"] = new_cm["data"]
            cm = self.kube.config_maps(job.namespace).update(cm)
        return cm

    def get_or_create_service(self, job: MPIJob, new_svc: dict) -> dict:
        """controller.go:907-930."""
        try:
            svc = self.service_lister.namespaced(job.namespace).get(M.name_of(new_svc))
        except errors.ApiError as e:
            if errors.is_not_found(e):
                return self.kube.services(job.namespace).create(new_svc)
            raise
        self._must_own(job, svc, "Service")
        if (svc["spec"].get("selector") != new_svc["spec"]["selector"]
                or bool(svc["spec"].get("publishNotReadyAddresses")) != bool(new_svc["spec"]["publishNotReadyAddresses"])):
            svc = copy.deepcopy(svc)
            svc["spec"]["selector"] = new_svc["spec"]["selector"]
            svc["spec"]["publishNotReadyAddresses"] = new_svc["spec"]["publishNotReadyAddresses"]
            return self.kube.services(job.namespace).update(svc)
        return svc

    def get_or_create_ssh_auth_secret(self, job: MPIJob) -> dict:
        """controller.go:934-963."""
        try:
            secret = self.secret_lister.namespaced(job.namespace).get(job.name + B.SSH_AUTH_SECRET_SUFFIX)
        except errors.ApiError as e:
            if errors.is_not_found(e):
                try:
                    return self.kube.secrets(job.namespace).create(B.new_ssh_auth_secret(job))
                except errors.ApiError as e2:
                    raise SyncError(f"creating SSH auth secret: {e2}")
            raise
        self._must_own(job, secret, "Secret")
        want_keys = sorted([B.SSH_PRIVATE_KEY, B.SSH_PUBLIC_KEY])
        if sorted((secret.get("data") or {}).keys()) != want_keys:
            secret = copy.deepcopy(secret)
            secret["data"] = B.new_ssh_auth_secret(job)["data"]
            return self.kube.secrets(job.namespace).update(secret)
        return secret

    def get_or_create_worker(self, job: MPIJob) -> List[dict]:
        """controller.go:976-1036 incl. scale-down of index >= replicas (elastic)."""
        pods: List[dict] = []
        wspec = job.spec.replica(C.REPLICA_TYPE_WORKER)
        if wspec is None:
            return pods
        replicas = int(wspec.replicas or 0)
        full = self.pod_lister.namespaced(job.namespace).list(B.worker_selector(job.name))
        if len(full) > replicas:
            for pod in full:
                idx = (M.meta(pod).get("labels") or {}).get(C.REPLICA_INDEX_LABEL)
                if idx is None:
                    return []  # the reference bails out with (nil, nil) when a worker lacks its index label
                try:
                    index = int(idx)
                except ValueError:
                    continue
                if index >= replicas:
                    self.kube.pods(M.namespace_of(pod)).delete(M.name_of(pod))
        for i in range(replicas):
            pod = None
            try:
                pod = self.pod_lister.namespaced(job.namespace).get(B.worker_name(job, i))
            except errors.ApiError as e:
                if not errors.is_not_found(e):
                    self.recorder.eventf(job, EVENT_TYPE_WARNING, S.MPIJOB_FAILED_REASON, "worker pod created failed: %s", e)
                    raise
                try:
                    pod = self.kube.pods(job.namespace).create(B.new_worker(job, i, self.pod_group_ctrl))
                except errors.ApiError as e2:
                    self.recorder.eventf(job, EVENT_TYPE_WARNING, S.MPIJOB_FAILED_REASON, "worker pod created failed: %s", e2)
                    raise
            if pod is not None:
                self._must_own(job, pod, "Pod")
            pods.append(pod)
        return pods

    # --------------------------------------------------------------- status --
    def job_pods(self, launcher: dict) -> List[dict]:
        """controller.go:1665-1681: pods selected by the Job and controlled by it."""
        sel = (launcher.get("spec", {}).get("selector") or {}).get("matchLabels")
        if sel is None:
            sel = {"job-name": M.name_of(launcher)}
        pods = self.pod_lister.namespaced(M.namespace_of(launcher)).list(sel)
        return [p for p in pods if M.is_controlled_by(p, launcher)]

    def update_mpijob_status(self, job: MPIJob, launcher: Optional[dict], worker: List[dict]) -> None:
        """controller.go:1088-1174."""
        old_status = copy.deepcopy(job.status)
        if is_mpijob_suspended(job):
            if S.update_mpijob_conditions(job, C.JOB_SUSPENDED, C.CONDITION_TRUE, S.MPIJOB_SUSPENDED_REASON, "MPIJob suspended", self._now()):
                self.recorder.event(job, EVENT_TYPE_NORMAL, "MPIJobSuspended", "MPIJob suspended")
        elif S.get_condition(job.status, C.JOB_SUSPENDED) is not None:
            if S.update_mpijob_conditions(job, C.JOB_SUSPENDED, C.CONDITION_FALSE, S.MPIJOB_RESUMED_REASON, "MPIJob resumed", self._now()):
                self.recorder.event(job, EVENT_TYPE_NORMAL, "MPIJobResumed", "MPIJob resumed")
                job.status.start_time = self._now()
        launcher_pods_cnt = 0
        if launcher is not None:
            launcher_pods = self.job_pods(launcher)
            # Job.status.active counts Pending too: count Running pods from the lister instead
            launcher_pods_cnt = count_running_pods(launcher_pods)
            S.initialize_replica_statuses(job, C.REPLICA_TYPE_LAUNCHER)
            lstat = job.status.replica_statuses[C.REPLICA_TYPE_LAUNCHER]
            lstat.failed = int(launcher.get("status", {}).get("failed", 0) or 0)
            if is_job_succeeded(launcher):
                lstat.succeeded = 1
                msg = f"MPIJob {job.namespace}/{job.name} successfully completed."
                self.recorder.event(job, EVENT_TYPE_NORMAL, S.MPIJOB_SUCCEEDED_REASON, msg)
                if job.status.completion_time is None:
                    job.status.completion_time = launcher.get("status", {}).get("completionTime")
                S.update_mpijob_conditions(job, C.JOB_SUCCEEDED, C.CONDITION_TRUE, S.MPIJOB_SUCCEEDED_REASON, msg, self._now())
                metrics.mpi_jobs_successful.inc()
                metrics.observe_job_duration(job, "Succeeded")
            elif is_job_failed(launcher):
                self.update_mpijob_failed_status(job, launcher, launcher_pods)
            else:
                lstat.active = launcher_pods_cnt
            metrics.mpi_job_info.labels(M.name_of(launcher), job.namespace).set(1)

        running = evict = 0
        S.initialize_replica_statuses(job, C.REPLICA_TYPE_WORKER)
        wstat = job.status.replica_statuses[C.REPLICA_TYPE_WORKER]
        for p in worker or []:
            ph = pod_phase(p)
            if ph == "Failed":
                wstat.failed += 1
                if p.get("status", {}).get("reason") == "Evicted":
                    evict += 1
            elif ph == "Succeeded":
                wstat.succeeded += 1
            elif ph == "Running":
                running += 1
                wstat.active += 1
        if evict > 0:
            msg = f"{evict}/{len(worker)} workers are evicted"
            log.info("MPIJob <%s/%s>: %s", job.namespace, job.name, msg)
            S.update_mpijob_conditions(job, C.JOB_FAILED, C.CONDITION_TRUE, S.MPIJOB_EVICT, msg, self._now())
            self.recorder.event(job, EVENT_TYPE_WARNING, S.MPIJOB_EVICT, msg)

        if is_mpijob_suspended(job):
            msg = f"MPIJob {job.namespace}/{job.name} is suspended."
            S.update_mpijob_conditions(job, C.JOB_RUNNING, C.CONDITION_FALSE, S.MPIJOB_SUSPENDED_REASON, msg, self._now())
        elif launcher is not None and launcher_pods_cnt >= 1 and running == len(worker or []):
            msg = f"MPIJob {job.namespace}/{job.name} is running."
            S.update_mpijob_conditions(job, C.JOB_RUNNING, C.CONDITION_TRUE, S.MPIJOB_RUNNING_REASON, msg, self._now())
            self.recorder.eventf(job, EVENT_TYPE_NORMAL, "MPIJobRunning", "MPIJob %s/%s is running", job.namespace, job.name)

        if old_status != job.status:
            self.update_status_handler(job)

    def update_mpijob_failed_status(self, job: MPIJob, launcher: dict, launcher_pods: List[dict]) -> None:
        """controller.go:1176-1207."""
        cond = get_job_condition(launcher, "Failed") or {}
        reason = cond.get("reason") or S.MPIJOB_FAILED_REASON
        msg = cond.get("message") or f"MPIJob {job.namespace}/{job.name} has failed"
        if reason == JOB_REASON_BACKOFF_LIMIT_EXCEEDED:
            last = None
            for p in launcher_pods:
                if is_pod_failed(p) and (last is None or M.meta(last).get("creationTimestamp", "") < M.meta(p).get("creationTimestamp", "")):
                    last = p
            if last is not None:
                reason += "/" + last.get("status", {}).get("reason", "")
                msg += ": " + last.get("status", {}).get("message", "")
                msg = truncate_message(msg)
        self.recorder.event(job, EVENT_TYPE_WARNING, reason, msg)
        if job.status.completion_time is None:
            job.status.completion_time = self._now()
        S.update_mpijob_conditions(job, C.JOB_FAILED, C.CONDITION_TRUE, reason, msg, self._now())
        metrics.mpi_jobs_failed.inc()
        metrics.observe_job_duration(job, "Failed")

    def do_update_job_status(self, job: MPIJob) -> None:
        """controller.go:1301-1304."""
        self.kubeflow.kubeflow_v2beta1().mpijobs(job.namespace).update_status(job)
