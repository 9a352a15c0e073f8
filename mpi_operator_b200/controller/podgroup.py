"""Gang scheduling: PodGroup control for Volcano- and scheduler-plugins-style
groups (reference: pkg/controller/podgroup.go:42-475).

On the single box the PodGroup is consumed by the local slot allocator
(``mpi_operator_b200.node.allocator``) instead of an external scheduler: a job's
pods only start when ``minMember`` slots (and ``minResources["nvidia.com/gpu"]``
GPUs) can be granted all-or-nothing; ``queue`` / ``priorityClassName`` order the
pending list and ``scheduleTimeoutSeconds`` bounds the wait.
"""
from __future__ import annotations

import copy
import logging
from typing import Dict, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..api.quantity import Quantity, add_to, to_strings
from ..api.types import MPIJob, ReplicaSpec
from ..client import errors
from ..client.clientset import KubeClient
from ..client.informers import Lister, SharedInformerFactory

log = logging.getLogger("mpi-job-controller")


def calculate_min_available(job: MPIJob) -> int:
    """podgroup.go:392-397: schedulingPolicy.minAvailable or workers + 1."""
    sp = job.spec.run_policy.scheduling_policy
    if sp is not None and sp.min_available is not None:
        return sp.min_available
    return job.worker_replicas() + 1


def calculate_priority_class_name(replicas: Dict[str, Optional[ReplicaSpec]], sp) -> str:
    """podgroup.go:403-416: policy > launcher template > worker template."""
    if sp is not None and sp.priority_class:
        return sp.priority_class
    for rt in (C.REPLICA_TYPE_LAUNCHER, C.REPLICA_TYPE_WORKER):
        r = (replicas or {}).get(rt)
        if r is not None and r.pod_spec.get("priorityClassName"):
            return r.pod_spec["priorityClassName"]
    return ""


def add_resources(min_resources: Dict[str, Quantity], resources: Optional[dict], replicas: int) -> None:
    """podgroup.go:420-443: requests win, limits fill the gaps."""
    if min_resources is None or not resources:
        return
    merged: Dict[str, Quantity] = {}
    for name, req in (resources.get("requests") or {}).items():
        merged[name] = Quantity.parse(req)
    for name, lim in (resources.get("limits") or {}).items():
        if name not in merged:
            merged[name] = Quantity.parse(lim)
    for name, q in merged.items():
        add_to(min_resources, name, q * replicas)


def cal_pg_min_resource(min_member: Optional[int], job: MPIJob, pc_lister: Optional[Lister]) -> Optional[Dict[str, str]]:
    """podgroup.go:337-388: sum over the highest-priority ``minMember`` replicas."""
    order = []
    for rt, replica in (job.spec.mpi_replica_specs or {}).items():
        if replica is None:
            continue
        prio = 0
        pc_name = replica.pod_spec.get("priorityClassName", "")
        if pc_name and pc_lister is not None:
            try:
                prio = int(pc_lister.get(pc_name).get("value", 0))
            except errors.ApiError as e:
                log.warning("Ignore replica %r priority class %r: %s", rt, pc_name, e)
        order.append({"priority": prio, "type": rt, "replicas": replica.replicas, "spec": replica})
    # sort.Sort(sort.Reverse(order)) on priority; Launcher first on ties for determinism
    order.sort(key=lambda r: (-r["priority"], 0 if r["type"] == C.REPLICA_TYPE_LAUNCHER else 1))
    if not order:
        return None
    replicas = order[0]["replicas"] or 0
    if len(order) > 1:
        replicas += order[1]["replicas"] or 0
    if min_member is not None and replicas > min_member:
        if len(order) > 1 and order[0]["priority"] == order[1]["priority"]:
            widx = next((i for i, r in enumerate(order) if r["type"] == C.REPLICA_TYPE_WORKER), -1)
            if widx == -1:
                log.warning("Couldn't find the worker replicas")
                return None
            order[widx]["replicas"] = min_member - 1
        elif len(order) > 1:
            order[1]["replicas"] = min_member - 1
    total: Dict[str, Quantity] = {}
    for r in order:
        if r["replicas"] is None:
            continue
        for c in r["spec"].containers:
            add_resources(total, c.get("resources"), int(r["replicas"]))
    return to_strings(total)


class PodGroupControl:
    """podgroup.go:42-65 interface."""
    resource = ""
    scheduler_name = ""

    def __init__(self, kube: KubeClient, informers: SharedInformerFactory, pc_lister: Optional[Lister] = None):
        self.kube = kube
        self.informer = informers.informer_for(self.resource)
        self.lister = informers.lister_for(self.resource)
        self.pc_lister = pc_lister

    # interface -----------------------------------------------------------------
    def new_pod_group(self, job: MPIJob) -> dict:
        raise NotImplementedError

    def decorate_pod_template_spec(self, template: dict, job_name: str) -> None:
        raise NotImplementedError

    def get_pod_group(self, namespace: str, name: str) -> dict:
        return self.lister.namespaced(namespace).get(name)

    def _client(self, ns):
        return self.kube._rc(self.resource, ns)

    def create_pod_group(self, pg: dict) -> dict:
        return self._client(M.namespace_of(pg)).create(pg)

    def update_pod_group(self, old: dict, new: dict) -> dict:
        old = copy.deepcopy(old)
        old["spec"] = copy.deepcopy(new["spec"])
        return self._client(M.namespace_of(old)).update(old)

    def delete_pod_group(self, namespace: str, name: str) -> None:
        self._client(namespace).delete(name)

    def calculate_pg_min_resources(self, min_member: Optional[int], job: MPIJob) -> Optional[Dict[str, str]]:
        sp = job.spec.run_policy.scheduling_policy
        if sp is not None and sp.min_resources is not None:
            return sp.min_resources
        if min_member is not None and min_member == 0:
            return None
        return cal_pg_min_resource(min_member, job, self.pc_lister)

    def pg_specs_are_equal(self, a: dict, b: dict) -> bool:
        return a.get("spec") == b.get("spec")

    def _meta(self, job: MPIJob) -> dict:
        return {"name": job.name, "namespace": job.namespace, "ownerReferences": [M.new_controller_ref(job.to_dict())]}


class VolcanoCtrl(PodGroupControl):
    """podgroup.go:68-194."""
    resource = "volcano-podgroups"
    scheduler_name = C.GANG_SCHEDULER_VOLCANO

    def new_pod_group(self, job: MPIJob) -> dict:
        min_member = calculate_min_available(job)
        queue = job.annotations.get(C.VOLCANO_QUEUE_NAME_ANNOTATION, "")
        sp = job.spec.run_policy.scheduling_policy
        if sp is not None and sp.queue:
            queue = sp.queue
        spec = {"minMember": min_member}
        if queue:
            spec["queue"] = queue
        pc = calculate_priority_class_name(job.spec.mpi_replica_specs, sp)
        if pc:
            spec["priorityClassName"] = pc
        mr = self.calculate_pg_min_resources(min_member, job)
        if mr is not None:
            spec["minResources"] = mr
        return {"apiVersion": "scheduling.volcano.sh/v1beta1", "kind": "PodGroup", "metadata": self._meta(job), "spec": spec}

    def decorate_pod_template_spec(self, template: dict, job_name: str) -> None:
        spec = template.setdefault("spec", {})
        if spec.get("schedulerName", "") != self.scheduler_name:
            log.warning("%s scheduler is specified when gang-scheduling is enabled and it will be overwritten", spec.get("schedulerName", ""))
        spec["schedulerName"] = self.scheduler_name
        template.setdefault("metadata", {}).setdefault("annotations", {})[C.VOLCANO_GROUP_NAME_ANNOTATION] = job_name


class SchedulerPluginsCtrl(PodGroupControl):
    """podgroup.go:197-334."""
    resource = "sched-podgroups"

    def __init__(self, kube, informers, scheduler_name: str, pc_lister=None):
        super().__init__(kube, informers, pc_lister)
        self.scheduler_name = scheduler_name

    def new_pod_group(self, job: MPIJob) -> dict:
        sp = job.spec.run_policy.scheduling_policy
        timeout = 0
        if sp is not None and sp.schedule_timeout_seconds is not None:
            timeout = sp.schedule_timeout_seconds
        min_member = calculate_min_available(job)
        spec = {"minMember": min_member, "scheduleTimeoutSeconds": timeout}
        mr = self.calculate_pg_min_resources(min_member, job)
        if mr:
            spec["minResources"] = mr
        return {"apiVersion": "scheduling.x-k8s.io/v1alpha1", "kind": "PodGroup", "metadata": self._meta(job), "spec": spec}

    def decorate_pod_template_spec(self, template: dict, job_name: str) -> None:
        spec = template.setdefault("spec", {})
        if spec.get("schedulerName", "") != self.scheduler_name:
            log.warning("%s scheduler is specified when gang-scheduling is enabled and it will be overwritten", spec.get("schedulerName", ""))
        spec["schedulerName"] = self.scheduler_name
        template.setdefault("metadata", {}).setdefault("labels", {})[C.SCHED_PLUGINS_POD_GROUP_LABEL] = job_name
