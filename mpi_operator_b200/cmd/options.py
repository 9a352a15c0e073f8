"""Daemon flags (reference: cmd/mpi-operator/app/options/options.go:25-96).

Every reference flag is accepted with the same name and default; the
Kubernetes-connection flags (--master, --kubeConfig, --kube-api-qps/burst) are
parsed and recorded but inert on a single box.  New flags configure the local
API endpoint, state directory and topology.
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass, field
from typing import List, Optional

from ..api import constants as C


@dataclass
class ServerOption:
    kubeconfig: str = ""
    master_url: str = ""
    threadiness: int = 2
    monitoring_port: int = 0
    print_version: bool = False
    gang_scheduling_name: str = ""
    namespace: str = ""
    lock_namespace: str = "mpi-operator"
    qps: int = 5
    burst: int = 10
    controller_rate_limit: int = 10
    controller_burst: int = 100
    cluster_domain: str = ""
    # single-box additions
    listen: str = "127.0.0.1:8087"
    healthz_port: int = 8080
    state_dir: str = ""
    fake_gpus: Optional[int] = None
    verbosity: int = 0
    leader_elect: bool = True
    auth_token_file: str = ""   # non-empty: the object API wants `Authorization: Bearer <token>` (created 0600 if missing)


def add_flags(p: argparse.ArgumentParser) -> None:
    d = ServerOption()
    p.add_argument("--master", dest="master_url", default="", help="(inert) address of the Kubernetes API server")
    p.add_argument("--kubeConfig", "--kubeconfig", dest="kubeconfig", default="", help="(inert) path to a kubeconfig")
    p.add_argument("--namespace", default=os.environ.get(C.ENV_KUBEFLOW_NAMESPACE, ""),
                   help="namespace to monitor mpijobs; empty = all namespaces")
    p.add_argument("--threadiness", type=int, default=d.threadiness, help="reconcile worker threads")
    p.add_argument("--version", dest="print_version", action="store_true", help="show version and quit")
    p.add_argument("--monitoring-port", type=int, default=d.monitoring_port, help="Prometheus /metrics port (0 = off)")
    p.add_argument("--gang-scheduling", dest="gang_scheduling_name", default="",
                   help='gang scheduler name: "volcano" or a scheduler-plugins scheduler name')
    p.add_argument("--lock-namespace", default=d.lock_namespace, help="namespace of the leader-election Lease")
    p.add_argument("--kube-api-qps", dest="qps", type=int, default=d.qps, help="(inert) QPS to the API server")
    p.add_argument("--kube-api-burst", dest="burst", type=int, default=d.burst, help="(inert) burst to the API server")
    p.add_argument("--controller-queue-rate-limit", dest="controller_rate_limit", type=int, default=d.controller_rate_limit)
    p.add_argument("--controller-queue-burst", dest="controller_burst", type=int, default=d.controller_burst)
    p.add_argument("--cluster-domain", default="", help="cluster domain appended to hostfile FQDNs")
    p.add_argument("--listen", default=os.environ.get("MPIJOB_SERVER", d.listen), help="host:port of the local REST API")
    p.add_argument("--healthz-port", type=int, default=d.healthz_port, help="/healthz port (0 = off)")
    p.add_argument("--state-dir", default=os.environ.get("B200MPI_STATE_DIR", ""), help="job store + pod sandboxes")
    p.add_argument("--fake-gpus", type=int, default=None, help="pretend the box has N GPUs (tests)")
    p.add_argument("-v", "--v", dest="verbosity", type=int, default=0, help="log verbosity (klog -v)")
    p.add_argument("--alsologtostderr", action="store_true", help="accepted for manifest compatibility")
    # the rest of klog's flag set (options.go:93-95 registers all of them): accepted, inert — logging goes to stderr
    p.add_argument("--logtostderr", nargs="?", const="true", default="true", help=argparse.SUPPRESS)
    p.add_argument("--stderrthreshold", default="", help=argparse.SUPPRESS)
    p.add_argument("--log_dir", "--log-dir", dest="log_dir", default="", help=argparse.SUPPRESS)
    p.add_argument("--log_file", "--log-file", dest="log_file", default="", help=argparse.SUPPRESS)
    p.add_argument("--vmodule", default="", help=argparse.SUPPRESS)
    p.add_argument("--skip_headers", "--skip-headers", dest="skip_headers", nargs="?", const="true", default="", help=argparse.SUPPRESS)
    p.add_argument("--no-leader-elect", dest="leader_elect", action="store_false")
    p.add_argument("--auth-token-file", default=os.environ.get("MPIJOB_TOKEN_FILE", ""),
                   help="require `Authorization: Bearer <token>` on the object API; the file is created (mode 0600, random token) "
                        "if it does not exist. The reference relies on the cluster's RBAC (manifests/base/cluster-role.yaml); on one "
                        "box this is what keeps other local users from creating pods (= running commands as the daemon's user)")


def parse(argv: Optional[List[str]] = None) -> ServerOption:
    p = argparse.ArgumentParser(prog="mpi-operator", description="single-box MPIJob operator daemon")
    add_flags(p)
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    # Go's flag package (the reference binary) takes -flag and --flag alike; its manifests use the single dash
    # (manifests/base/deployment.yaml: `-alsologtostderr`), so do the same here
    argv = ["-" + a if (a.startswith("-") and not a.startswith("--") and len(a.split("=", 1)[0]) > 2) else a for a in argv]
    ns = p.parse_args(argv)
    opt = ServerOption()
    for k in opt.__dataclass_fields__:
        if hasattr(ns, k):
            setattr(opt, k, getattr(ns, k))
    return opt
