"""Builders for every object the controller owns.

Reference: pkg/controller/mpi_job_controller.go:1309-1645,1764-1787 —
``newConfigMap`` (hostfile), ``updateDiscoverHostsInConfigMap``,
``newJobService``, ``newSSHAuthSecret``, ``newWorker``, ``newLauncherJob``,
``newLauncherPodTemplate``, ``setupSSHOnPod``.  Outputs are k8s-JSON dicts and
are byte-compatible where the reference has goldens (hostfile,
discover_hosts.sh, env names, labels; mpi_job_controller_test.go:1424-2402).
On the single box the ssh Secret is inert (no sshd) but still produced, so a
user's ``sshAuthMountPath`` keeps its meaning and the object graph matches.
"""
from __future__ import annotations

import base64
import copy
from typing import Dict, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..api.types import MPIJob, ReplicaSpec

CONTROLLER_AGENT_NAME = "mpi-job-controller"
CONFIG_SUFFIX = "-config"
CONFIG_VOLUME_NAME = "mpi-job-config"
CONFIG_MOUNT_PATH = "/etc/mpi"
HOSTFILE_NAME = "hostfile"
DISCOVER_HOSTS_SCRIPT_NAME = "discover_hosts.sh"
SSH_AUTH_SECRET_SUFFIX = "-ssh"
SSH_AUTH_VOLUME = "ssh-auth"
ROOT_SSH_PATH = "/root/.ssh"
LAUNCHER = "launcher"
WORKER = "worker"
LAUNCHER_SUFFIX = "-launcher"
WORKER_SUFFIX = "-worker"
SSH_PUBLIC_KEY = "ssh-publickey"
SSH_PRIVATE_KEY = "ssh-privatekey"  # corev1.SSHAuthPrivateKey
SSH_PRIVATE_KEY_FILE = "id_rsa"
SSH_PUBLIC_KEY_FILE = SSH_PRIVATE_KEY_FILE + ".pub"
SSH_AUTHORIZED_KEYS_FILE = "authorized_keys"
OPENMPI_SLOTS_ENV = "OMPI_MCA_orte_set_default_slots"
INTEL_MPI_SLOTS_ENV = "I_MPI_PERHOST"

SSH_VOLUME_ITEMS = [
    {"key": SSH_PRIVATE_KEY, "path": SSH_PRIVATE_KEY_FILE},
    {"key": SSH_PUBLIC_KEY, "path": SSH_PUBLIC_KEY_FILE},
    {"key": SSH_PUBLIC_KEY, "path": SSH_AUTHORIZED_KEYS_FILE},
]
CONFIG_VOLUME_ITEMS = [
    {"key": HOSTFILE_NAME, "path": HOSTFILE_NAME, "mode": 0o444},
    {"key": DISCOVER_HOSTS_SCRIPT_NAME, "path": DISCOVER_HOSTS_SCRIPT_NAME, "mode": 0o555},
]
LAUNCHER_ENV_VARS = [{"name": "K_MPI_JOB_ROLE", "value": LAUNCHER}]
WORKER_ENV_VARS = [{"name": "K_MPI_JOB_ROLE", "value": WORKER}]
OMPI_ENV_VARS = [
    {"name": "OMPI_MCA_orte_keep_fqdn_hostnames", "value": "true"},
    {"name": "OMPI_MCA_orte_default_hostfile", "value": f"{CONFIG_MOUNT_PATH}/{HOSTFILE_NAME}"},
    {"name": "OMPI_MCA_plm_rsh_args", "value": "-o ConnectionAttempts=10"},
]
INTEL_ENV_VARS = [
    {"name": "I_MPI_HYDRA_HOST_FILE", "value": f"{CONFIG_MOUNT_PATH}/{HOSTFILE_NAME}"},
    {"name": "I_MPI_HYDRA_BOOTSTRAP_EXEC_EXTRA_ARGS", "value": "-o ConnectionAttempts=10"},
]
MPICH_ENV_VARS = [
    {"name": "HYDRA_HOST_FILE", "value": f"{CONFIG_MOUNT_PATH}/{HOSTFILE_NAME}"},
    {"name": "HYDRA_LAUNCH_EXTRA_ARGS", "value": "-o ConnectionAttempts=10"},
]
NVIDIA_DISABLE_ENV_VARS = [{"name": "NVIDIA_VISIBLE_DEVICES"}, {"name": "NVIDIA_DRIVER_CAPABILITIES"}]


def _owner(job: MPIJob) -> List[dict]:
    return [M.new_controller_ref(job.to_dict())]


def run_launcher_as_worker(job: MPIJob) -> bool:
    return bool(job.spec.run_launcher_as_worker)


def worker_name(job: MPIJob, index: int) -> str:
    return f"{job.name}{WORKER_SUFFIX}-{index}"


def launcher_name(job: MPIJob) -> str:
    return job.name + LAUNCHER_SUFFIX


def worker_replica_index_label(job: MPIJob, index: int) -> str:
    """controller.go:1461-1468: padded by one when the launcher is also a worker."""
    return str(index + 1 if run_launcher_as_worker(job) else index)


def default_labels(job_name: str, role: str) -> Dict[str, str]:
    return {C.OPERATOR_NAME_LABEL: C.OPERATOR_NAME, C.JOB_NAME_LABEL: job_name, C.JOB_ROLE_LABEL: role}


def worker_selector(job_name: str) -> Dict[str, str]:
    return default_labels(job_name, WORKER)


def _domain_format(cluster_domain: str):
    def fmt(pod: str, job: str, ns: str) -> str:
        s = f"{pod}.{job}.{ns}.svc"
        return s + f".{cluster_domain}" if cluster_domain else s
    return fmt


def hostfile_text(job: MPIJob, worker_replicas: int, cluster_domain: str = "") -> str:
    """controller.go:1309-1337 (goldens: mpi_job_controller_test.go:1920-2145)."""
    slots = job.spec.slots_per_worker if job.spec.slots_per_worker is not None else 1
    fmt = _domain_format(cluster_domain)
    impl = job.spec.mpi_implementation
    lines = []

    def line(host):
        if impl == C.MPI_IMPLEMENTATION_OPENMPI:
            lines.append(f"{host} slots={slots}\n")
        elif impl in (C.MPI_IMPLEMENTATION_INTEL, C.MPI_IMPLEMENTATION_MPICH):
            lines.append(f"{host}:{slots}\n")
    if run_launcher_as_worker(job):
        line(fmt(launcher_name(job), job.name, job.namespace))
    for i in range(worker_replicas):
        line(fmt(worker_name(job, i), job.name, job.namespace))
    return "".join(lines)


def discover_hosts_text(job: MPIJob, running_pods: List[dict], cluster_domain: str = "") -> str:
    """controller.go:1357-1381 (goldens: mpi_job_controller_test.go:2188-2382)."""
    fmt = _domain_format(cluster_domain)
    out = ["#!/bin/sh\n"]
    if run_launcher_as_worker(job):
        out.append(f"echo {fmt(launcher_name(job), job.name, job.namespace)}\n")
    for p in sorted(running_pods, key=M.name_of):
        out.append(f"echo {fmt(M.name_of(p), job.name, M.namespace_of(p))}\n")
    return "".join(out)


def new_config_map(job: MPIJob, worker_replicas: int, cluster_domain: str = "") -> dict:
    return {
        "apiVersion": "v1", "kind": "ConfigMap",
        "metadata": {"name": job.name + CONFIG_SUFFIX, "namespace": job.namespace, "labels": {"app": job.name},
                     "ownerReferences": _owner(job)},
        "data": {HOSTFILE_NAME: hostfile_text(job, worker_replicas, cluster_domain)},
    }


def update_discover_hosts_in_config_map(cm: dict, job: MPIJob, running_pods: List[dict], cluster_domain: str = "") -> None:
    cm["data"][DISCOVER_HOSTS_SCRIPT_NAME] = discover_hosts_text(job, running_pods, cluster_domain)


def new_job_service(job: MPIJob) -> dict:
    """Headless Service fronting launcher + workers (controller.go:1384-1412)."""
    selector = {C.OPERATOR_NAME_LABEL: C.OPERATOR_NAME, C.JOB_NAME_LABEL: job.name}
    return {
        "apiVersion": "v1", "kind": "Service",
        "metadata": {"name": job.name, "namespace": job.namespace, "labels": {"app": job.name}, "ownerReferences": _owner(job)},
        "spec": {"clusterIP": "None", "selector": selector,
                 # must be true only with runLauncherAsWorker, to avoid waiting on launcher readiness
                 "publishNotReadyAddresses": run_launcher_as_worker(job)},
    }


def new_ssh_auth_secret(job: MPIJob) -> dict:
    """ECDSA P-521 key pair (controller.go:1416-1451)."""
    from cryptography.hazmat.primitives import serialization
    from cryptography.hazmat.primitives.asymmetric import ec
    key = ec.generate_private_key(ec.SECP521R1())
    private_pem = key.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.TraditionalOpenSSL,
                                    serialization.NoEncryption())
    public = key.public_key().public_bytes(serialization.Encoding.OpenSSH, serialization.PublicFormat.OpenSSH) + b"\n"
    return {
        "apiVersion": "v1", "kind": "Secret", "type": "kubernetes.io/ssh-auth",
        "metadata": {"name": job.name + SSH_AUTH_SECRET_SUFFIX, "namespace": job.namespace, "labels": {"app": job.name},
                     "ownerReferences": _owner(job)},
        "data": {SSH_PRIVATE_KEY: base64.b64encode(private_pem).decode(), SSH_PUBLIC_KEY: base64.b64encode(public).decode()},
    }


def set_restart_policy(template: dict, spec: ReplicaSpec) -> None:
    """controller.go:1693-1699: ExitCode degrades to Never."""
    pol = C.RESTART_POLICY_NEVER if spec.restart_policy == C.RESTART_POLICY_EXIT_CODE else spec.restart_policy
    template.setdefault("spec", {})["restartPolicy"] = pol


def setup_ssh_on_pod(pod_spec: dict, job: MPIJob) -> None:
    """controller.go:1764-1787 (0600 only for the default /root/.ssh path)."""
    secret = {"secretName": job.name + SSH_AUTH_SECRET_SUFFIX, "items": copy.deepcopy(SSH_VOLUME_ITEMS)}
    if job.spec.ssh_auth_mount_path == ROOT_SSH_PATH:
        secret["defaultMode"] = 0o600
    pod_spec.setdefault("volumes", []).append({"name": SSH_AUTH_VOLUME, "secret": secret})
    main = pod_spec["containers"][0]
    main.setdefault("volumeMounts", []).append({"name": SSH_AUTH_VOLUME, "mountPath": job.spec.ssh_auth_mount_path})


def new_worker(job: MPIJob, index: int, pod_group_ctrl=None) -> dict:
    """controller.go:1473-1526."""
    name = worker_name(job, index)
    wspec = job.spec.replica(C.REPLICA_TYPE_WORKER)
    tmpl = copy.deepcopy(wspec.template)
    md = tmpl.setdefault("metadata", {})
    labels = md.get("labels") or {}
    labels.update(default_labels(job.name, WORKER))
    labels[C.REPLICA_INDEX_LABEL] = worker_replica_index_label(job, index)
    md["labels"] = labels
    spec = tmpl.setdefault("spec", {})
    spec["hostname"] = name
    spec["subdomain"] = job.name  # matches the job's Service name
    if spec.get("hostNetwork"):
        spec["dnsPolicy"] = "ClusterFirstWithHostNet"
    search = f"{job.name}.{job.namespace}.svc.cluster.local"
    if spec.get("dnsConfig") is None:
        spec["dnsConfig"] = {"searches": [search]}
    else:
        spec["dnsConfig"].setdefault("searches", []).append(search)
    set_restart_policy(tmpl, wspec)
    container = spec["containers"][0]
    if not container.get("command") and not container.get("args"):
        container["command"] = ["/usr/sbin/sshd", "-De"]
    container.setdefault("env", []).extend(copy.deepcopy(WORKER_ENV_VARS))
    setup_ssh_on_pod(spec, job)
    if pod_group_ctrl is not None:
        pod_group_ctrl.decorate_pod_template_spec(tmpl, job.name)
    pod_md = {"name": name, "namespace": job.namespace, "labels": tmpl["metadata"].get("labels"), "ownerReferences": _owner(job)}
    if tmpl["metadata"].get("annotations"):
        pod_md["annotations"] = tmpl["metadata"]["annotations"]
    return {"apiVersion": "v1", "kind": "Pod", "metadata": pod_md, "spec": tmpl["spec"]}


def new_launcher_pod_template(job: MPIJob, pod_group_ctrl=None, recorder=None) -> dict:
    """controller.go:1556-1645."""
    lname = launcher_name(job)
    lspec = job.spec.replica(C.REPLICA_TYPE_LAUNCHER)
    tmpl = copy.deepcopy(lspec.template)
    md = tmpl.setdefault("metadata", {})
    labels = md.get("labels") or {}
    labels.update(default_labels(job.name, LAUNCHER))
    md["labels"] = labels
    if pod_group_ctrl is not None:
        pod_group_ctrl.decorate_pod_template_spec(tmpl, job.name)
    if run_launcher_as_worker(job):
        md["labels"][C.REPLICA_INDEX_LABEL] = "0"
    spec = tmpl.setdefault("spec", {})
    spec["hostname"] = lname
    spec["subdomain"] = job.name
    if spec.get("hostNetwork"):
        spec["dnsPolicy"] = "ClusterFirstWithHostNet"
    container = spec["containers"][0]
    env = container.setdefault("env", [])
    env.extend(copy.deepcopy(LAUNCHER_ENV_VARS))
    slots = str(int(job.spec.slots_per_worker))
    impl = job.spec.mpi_implementation
    if impl == C.MPI_IMPLEMENTATION_OPENMPI:
        env.extend(copy.deepcopy(OMPI_ENV_VARS))
        env.append({"name": OPENMPI_SLOTS_ENV, "value": slots})
    elif impl == C.MPI_IMPLEMENTATION_INTEL:
        env.extend(copy.deepcopy(INTEL_ENV_VARS))
        env.append({"name": INTEL_MPI_SLOTS_ENV, "value": slots})
    elif impl == C.MPI_IMPLEMENTATION_MPICH:
        env.extend(copy.deepcopy(MPICH_ENV_VARS))
    if not run_launcher_as_worker(job):
        # keep the launcher off the GPUs (controller.go:1600-1606)
        env.extend(copy.deepcopy(NVIDIA_DISABLE_ENV_VARS))
    setup_ssh_on_pod(spec, job)
    if spec.get("restartPolicy"):
        msg = "Restart policy in pod template overridden by restart policy in replica spec"
        if recorder is not None:
            recorder.event(job, "Warning", "SetPodTemplateRestartPolicy", msg)
    set_restart_policy(tmpl, lspec)
    spec.setdefault("volumes", []).append({
        "name": CONFIG_VOLUME_NAME,
        "configMap": {"name": job.name + CONFIG_SUFFIX, "items": copy.deepcopy(CONFIG_VOLUME_ITEMS)}})
    container.setdefault("volumeMounts", []).append({"name": CONFIG_VOLUME_NAME, "mountPath": CONFIG_MOUNT_PATH})
    out_md = {"labels": tmpl["metadata"].get("labels"), "ownerReferences": _owner(job)}
    if tmpl["metadata"].get("annotations"):
        out_md["annotations"] = tmpl["metadata"]["annotations"]
    return {"metadata": out_md, "spec": tmpl["spec"]}


def new_launcher_job(job: MPIJob, pod_group_ctrl=None, recorder=None) -> dict:
    """batch/v1 Job wrapping the launcher pod template (controller.go:1528-1551)."""
    spec = {"template": new_launcher_pod_template(job, pod_group_ctrl, recorder)}
    rp = job.spec.run_policy
    if rp.ttl_seconds_after_finished is not None:
        spec["ttlSecondsAfterFinished"] = rp.ttl_seconds_after_finished
    if rp.active_deadline_seconds is not None:
        spec["activeDeadlineSeconds"] = rp.active_deadline_seconds
    if rp.backoff_limit is not None:
        spec["backoffLimit"] = rp.backoff_limit
    if bool(rp.suspend):
        spec["suspend"] = True
    return {
        "apiVersion": "batch/v1", "kind": "Job",
        "metadata": {"name": launcher_name(job), "namespace": job.namespace, "labels": {"app": job.name}, "ownerReferences": _owner(job)},
        "spec": spec,
    }


def merge_maps(a: Optional[dict], b: Optional[dict]) -> dict:
    out = dict(a or {})
    out.update(b or {})
    return out


def sync_launcher_scheduling_directives(launcher: dict, desired_template: dict) -> None:
    """KEP-2926 mutable directives on resume (controller.go:1654-1663)."""
    t = launcher["spec"]["template"]
    t.setdefault("metadata", {})
    t["metadata"]["labels"] = merge_maps(t["metadata"].get("labels"), desired_template["metadata"].get("labels"))
    ann = merge_maps(t["metadata"].get("annotations"), desired_template["metadata"].get("annotations"))
    if ann:
        t["metadata"]["annotations"] = ann
    for k in ("nodeSelector", "tolerations", "schedulingGates"):
        v = desired_template["spec"].get(k)
        if v:
            t["spec"][k] = copy.deepcopy(v)
        else:
            t["spec"].pop(k, None)
