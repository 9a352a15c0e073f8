"""Operator process: leader election, health, metrics, REST API, controller +
node agent wiring (reference: cmd/mpi-operator/app/server.go:79-314,
cmd/mpi-operator/main.go:29-53)."""
from __future__ import annotations

import fcntl
import json
import logging
import os
import signal
import socket
import sys
import threading
import time
import urllib.parse
import uuid
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Optional

from .. import version
from ..api import constants as C
from ..api import meta as M
from ..api.register import scheme
from ..api.types import MPIJob
from ..api.validation import validate_mpijob
from ..client import errors
from ..client.clientset import Clientset, KubeClient
from ..client.informers import SharedInformerFactory
from ..client.store import RESOURCES, ObjectStore
from ..controller import metrics
from ..controller.controller import MPIJobController
from ..node.agent import NodeAgent
from ..node.topology import discover_topology
from .options import ServerOption

log = logging.getLogger("mpi-operator")

# leader election timings (server.go:61-63)
LEASE_DURATION = 15.0
RENEW_DURATION = 5.0
RETRY_DURATION = 3.0
LEADER_LOCK_NAME = "mpi-operator"

# REST prefixes -> store resource
API_GROUPS = {
    ("api/v1", "pods"): "pods", ("api/v1", "services"): "services", ("api/v1", "configmaps"): "configmaps",
    ("api/v1", "secrets"): "secrets", ("api/v1", "events"): "events",
    ("apis/batch/v1", "jobs"): "jobs", ("apis/kubeflow.org/v2beta1", "mpijobs"): "mpijobs",
    ("apis/coordination.k8s.io/v1", "leases"): "leases",
    ("apis/scheduling.k8s.io/v1", "priorityclasses"): "priorityclasses",
    ("apis/scheduling.volcano.sh/v1beta1", "podgroups"): "volcano-podgroups",
    ("apis/scheduling.x-k8s.io/v1alpha1", "podgroups"): "sched-podgroups",
}


class LeaderElector:
    """Lease lock: an flock()ed file is the mutual exclusion, a Lease object in the
    store is the observable record (holderIdentity / renewTime), renewed every
    RENEW_DURATION; candidates retry every RETRY_DURATION."""

    def __init__(self, store: ObjectStore, state_dir: str, lock_namespace: str, identity: Optional[str] = None):
        self.store, self.ns = store, lock_namespace
        self.identity = identity or f"{socket.gethostname()}_{uuid.uuid4()}"
        os.makedirs(os.path.join(state_dir, "leases"), exist_ok=True)
        self.path = os.path.join(state_dir, "leases", f"{lock_namespace}.{LEADER_LOCK_NAME}.lock")
        self._fd: Optional[int] = None
        self.last_renew = 0.0
        self._stop = threading.Event()

    def try_acquire(self) -> bool:
        fd = os.open(self.path, os.O_CREAT | os.O_RDWR, 0o600)
        try:
            fcntl.flock(fd, fcntl.LOCK_EX | fcntl.LOCK_NB)
        except OSError:
            os.close(fd)
            return False
        self._fd = fd
        self._renew()
        return True

    def _renew(self) -> None:
        now = M.now_rfc3339()
        lease = {"apiVersion": "coordination.k8s.io/v1", "kind": "Lease",
                 "metadata": {"name": LEADER_LOCK_NAME, "namespace": self.ns},
                 "spec": {"holderIdentity": self.identity, "leaseDurationSeconds": int(LEASE_DURATION), "renewTime": now}}
        try:
            cur = self.store.get("leases", self.ns, LEADER_LOCK_NAME)
            if cur["spec"].get("holderIdentity") != self.identity:
                lease["spec"]["acquireTime"] = now
                lease["spec"]["leaseTransitions"] = int(cur["spec"].get("leaseTransitions", 0)) + 1
            else:
                lease["spec"]["acquireTime"] = cur["spec"].get("acquireTime", now)
                lease["spec"]["leaseTransitions"] = cur["spec"].get("leaseTransitions", 0)
            lease["metadata"]["resourceVersion"] = cur["metadata"]["resourceVersion"]
            self.store.update("leases", lease)
        except errors.ApiError:
            lease["spec"]["acquireTime"] = now
            self.store.create("leases", lease)
        self.last_renew = time.time()

    def run(self, on_started, on_stopped) -> None:
        while not self._stop.is_set() and not self.try_acquire():
            log.info("failed to acquire lease %s/%s; retrying in %.0fs", self.ns, LEADER_LOCK_NAME, RETRY_DURATION)
            self._stop.wait(RETRY_DURATION)
        if self._stop.is_set():
            return
        log.info("successfully acquired lease %s/%s", self.ns, LEADER_LOCK_NAME)
        metrics.is_leader.set(1)
        on_started()
        while not self._stop.wait(RENEW_DURATION):
            try:
                self._renew()
            except Exception:  # noqa: BLE001
                log.exception("failed to renew lease")
                if time.time() - self.last_renew > LEASE_DURATION:
                    break
        metrics.is_leader.set(0)
        on_stopped()

    def healthy(self, slack: float = 20.0) -> bool:
        """LeaderHealthzAdaptor: unhealthy if we hold the lease but failed to renew for lease+slack."""
        return self._fd is None or time.time() - self.last_renew < LEASE_DURATION + slack

    def release(self) -> None:
        self._stop.set()
        if self._fd is not None:
            try:
                fcntl.flock(self._fd, fcntl.LOCK_UN)
                os.close(self._fd)
            except OSError:
                pass
            self._fd = None


class Operator:
    """Everything app.Run wires together, usable in-process (tests, `mpijobctl run`)."""

    def __init__(self, opt: Optional[ServerOption] = None, store: Optional[ObjectStore] = None, clock=None):
        self.opt = opt or ServerOption()
        self._ephemeral_state = not self.opt.state_dir      # no --state-dir: a throw-away directory, removed again by stop()
        self.state_dir = self.opt.state_dir or os.path.join(os.environ.get("TMPDIR", "/tmp"), f"b200mpi-operator-{os.getpid()}")
        os.makedirs(self.state_dir, exist_ok=True)
        self.auth_token = _load_or_create_token(self.opt.auth_token_file) if self.opt.auth_token_file else ""
        if self.opt.fake_gpus is not None:
            os.environ["B200MPI_FAKE_GPUS"] = str(self.opt.fake_gpus)
        self.store = store or ObjectStore(os.path.join(self.state_dir, "store.json") if self.opt.state_dir else None)
        self.kube = KubeClient(self.store)
        self.clientset = Clientset(self.store)
        self.informers = SharedInformerFactory(self.store, self.opt.namespace)
        self.controller = MPIJobController(
            self.kube, self.clientset, self.informers, gang_scheduling=self.opt.gang_scheduling_name,
            cluster_domain=self.opt.cluster_domain, clock=clock, queue_rate_limit=self.opt.controller_rate_limit,
            queue_burst=self.opt.controller_burst, namespace=self.opt.namespace)
        self.agent = NodeAgent(self.store, discover_topology(), os.path.join(self.state_dir, "node"))
        self.elector = LeaderElector(self.store, self.state_dir, self.opt.lock_namespace)
        from ..controller.events import EventRecorder
        from ..node.health import GpuHealthMonitor
        self.gpu_health = GpuHealthMonitor(self.agent, recorder=EventRecorder(self.store, "node-agent"))
        self._http: list = []
        self._started = False

    def check_crd_exists(self) -> bool:
        """server.go:302-314: the daemon refuses to run without the MPIJob kind registered."""
        return self.clientset.discovery_has_mpijob_crd() and scheme.recognizes(C.API_VERSION, C.KIND)

    def start(self, leader_elect: Optional[bool] = None) -> None:
        if not self.check_crd_exists():
            log.error("CRD doesn't exist. Exiting")
            raise SystemExit(1)
        if self._started:
            return
        self._started = True
        if leader_elect is None:
            leader_elect = self.opt.leader_elect
        if leader_elect:
            t = threading.Thread(target=self.elector.run, args=(self._run_leading, self._stopped_leading), daemon=True)
            t.start()
        else:
            self._run_leading()

    def _run_leading(self) -> None:
        self.controller.run(self.opt.threadiness)
        self.agent.start()
        if self.agent.topology.source not in ("fake", "none"):   # real GPUs: NVML health probe -> cordon (node/health.py)
            self.gpu_health.start()

    def _stopped_leading(self) -> None:
        log.critical("leader election lost")
        os._exit(1)

    def stop(self) -> None:
        self.gpu_health.stop()
        for s in self._http:
            s.shutdown()
        self.controller.stop()
        self.agent.stop()
        self.informers.stop()
        self.elector.release()
        self.store.close()
        if self._ephemeral_state and os.environ.get("B200MPI_KEEP_STATE") != "1":
            import shutil
            shutil.rmtree(self.state_dir, ignore_errors=True)

    # ------------------------------------------------------------- serving --
    def serve(self, listen: str, block: bool = False, restricted: bool = False):
        """``restricted``: GET /metrics and /healthz only (what the reference serves on --monitoring-port, main.go:29-40,
        and on :8080, server.go:188-204). The object API (create pods = run commands, read Secrets) is served only on
        ``--listen``, which defaults to loopback."""
        host, _, port = listen.rpartition(":")
        srv = ThreadingHTTPServer((host or "127.0.0.1", int(port)), _make_handler(self, restricted=restricted))
        srv.daemon_threads = True
        self._http.append(srv)
        t = threading.Thread(target=srv.serve_forever, daemon=True)
        t.start()
        if block:
            t.join()
        return srv


def _load_or_create_token(path: str) -> str:
    """--auth-token-file: read the bearer token, or create the file (0600) with a random one on first start."""
    import secrets
    try:
        with open(path) as f:
            tok = f.read().strip()
        if tok:
            return tok
    except FileNotFoundError:
        pass
    tok = secrets.token_urlsafe(32)
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
    with os.fdopen(fd, "w") as f:
        f.write(tok + "\n")
    return tok


MAX_BODY_BYTES = 16 << 20     # larger than any object the apiserver would take (etcd's limit is 1.5 MiB; generous for ConfigMaps)


def _make_handler(op: Operator, restricted: bool = False):
    store = op.store

    class H(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def _refuse(self):
            """Restricted listeners (monitoring / healthz ports): nothing but GET /metrics and GET /healthz. The object API
            with --auth-token-file: everything but GET /healthz and /version needs the bearer token (401 otherwise)."""
            if not restricted:
                token = getattr(op, "auth_token", "")
                if not token or (self.command == "GET" and self._route()[0] in (["healthz"], ["version"])):
                    return False
                import hmac
                got = self.headers.get("Authorization", "")
                if got.startswith("Bearer ") and hmac.compare_digest(got[7:].strip().encode(), token.encode()):
                    return False
                n = int(self.headers.get("Content-Length", 0) or 0)
                if n:
                    self.rfile.read(n)   # keep the connection in step (HTTP/1.1 keep-alive)
                self._send(401, {"kind": "Status", "apiVersion": "v1", "status": "Failure", "reason": "Unauthorized", "code": 401,
                                 "message": "Unauthorized: this daemon was started with --auth-token-file; send "
                                            "'Authorization: Bearer <token>' (clients read MPIJOB_TOKEN or MPIJOB_TOKEN_FILE)"})
                return True
            if self.command == "GET" and self._route()[0] in (["metrics"], ["healthz"]):
                return False
            self._send(404 if self.command == "GET" else 405, {"kind": "Status", "status": "Failure", "code": 404 if self.command == "GET" else 405,
                                                                "message": "this port serves GET /metrics and GET /healthz only"})
            return True

        def log_message(self, fmt, *args):  # quiet
            log.debug("http: " + fmt, *args)

        def _send(self, code: int, body, ctype="application/json"):
            data = body if isinstance(body, bytes) else (json.dumps(body).encode() if ctype == "application/json" else str(body).encode())
            self.send_response(code)
            self.send_header("Content-Type", ctype)
            self.send_header("Content-Length", str(len(data)))
            self.end_headers()
            self.wfile.write(data)

        def _body(self):
            try:
                n = int(self.headers.get("Content-Length", 0) or 0)
            except ValueError:
                raise errors.bad_request("Content-Length is not a number")
            if n > MAX_BODY_BYTES:
                self.close_connection = True      # the unread body would be parsed as the next request
                raise errors.ApiError("RequestEntityTooLarge", f"request body of {n} bytes exceeds the limit of {MAX_BODY_BYTES}", 413)
            raw = self.rfile.read(n) if n > 0 else b""
            try:
                obj = json.loads(raw) if raw.strip() else {}
            except (ValueError, RecursionError) as e:     # RecursionError: absurdly nested input
                raise errors.bad_request(f"the request body is not valid JSON: {e}")
            if not isinstance(obj, dict):
                raise errors.bad_request(f"the request body must be a JSON object, got {type(obj).__name__}")
            return obj

        def _route(self):
            u = urllib.parse.urlparse(self.path)
            parts = [p for p in u.path.split("/") if p]
            q = urllib.parse.parse_qs(u.query)
            return parts, q

        def _resolve(self, parts):
            """-> (resource, namespace, name, subresource) or None."""
            for (prefix, plural), res in API_GROUPS.items():
                pp = prefix.split("/")
                if parts[:len(pp)] != pp:
                    continue
                rest = parts[len(pp):]
                ns = ""
                if len(rest) >= 2 and rest[0] == "namespaces":
                    ns, rest = rest[1], rest[2:]
                if not rest or rest[0] != plural:
                    continue
                name = rest[1] if len(rest) > 1 else ""
                sub = rest[2] if len(rest) > 2 else ""
                if (name and M.name_problem(name)) or (ns and M.name_problem(ns)):
                    return None       # not an object name (and never a path component of the state dir): 404
                return res, ns, name, sub
            return None

        def do_GET(self):  # noqa: N802
            if self._refuse():
                return None
            parts, q = self._route()
            try:
                if parts == ["healthz"]:
                    ok = op.elector.healthy()
                    return self._send(200 if ok else 500, b"ok" if ok else b"leader election lease expired", "text/plain")
                if parts == ["metrics"]:
                    return self._send(200, metrics.render(), "text/plain; version=0.0.4")
                if parts == ["version"]:
                    return self._send(200, version.info())
                if parts == ["topology"]:
                    return self._send(200, self._topology())
                if parts[:3] == ["api", "v1", "nodes"] and len(parts) <= 4:
                    node = self._node()
                    if len(parts) == 4:
                        if parts[3] != node["metadata"]["name"]:
                            return self._send(404, errors.not_found("nodes", parts[3]).to_status())
                        return self._send(200, node)
                    return self._send(200, {"apiVersion": "v1", "kind": "NodeList", "metadata": {}, "items": [node]})
                disc = self._discovery(parts)
                if disc is not None:
                    return self._send(200, disc)
                r = self._resolve(parts)
                if r is None:
                    return self._send(404, errors.not_found("path", self.path).to_status())
                res, ns, name, sub = r
                if res == "pods" and sub == "log":
                    if q.get("follow", ["false"])[0].lower() in ("true", "1"):
                        return self._follow_log(ns, name, q)
                    text = op.agent.logs(ns, name)
                    if "tailLines" in q:       # kubectl logs --tail=N
                        n_tail = max(0, int(q["tailLines"][0]))
                        text = "".join(text.splitlines(keepends=True)[-n_tail:]) if n_tail else ""
                    return self._send(200, text.encode(), "text/plain")
                if name:
                    return self._send(200, store.get(res, ns, name))
                sel = None
                if "labelSelector" in q:
                    sel = dict(kv.split("=", 1) for kv in q["labelSelector"][0].split(",") if "=" in kv)
                if q.get("watch", ["false"])[0].lower() in ("true", "1"):
                    return self._watch(res, ns, sel, q)
                items = store.list(res, ns or None, sel)
                api_version, kind, _ = RESOURCES[res]
                return self._send(200, {"apiVersion": api_version, "kind": kind + "List", "metadata": {}, "items": items})
            except errors.ApiError as e:
                return self._send(e.code, e.to_status())

        def _watch(self, res, ns, sel, q):
            """`?watch=true` (what client-go informers and `kubectl get -w` use): one JSON object per line,
            {"type": ADDED|MODIFIED|DELETED, "object": ...}, chunked, existing objects first, until `timeoutSeconds` (default
            300) or the client goes away. Filters: namespace of the path, `labelSelector`, `fieldSelector=metadata.name=<n>`."""
            import queue
            events: "queue.Queue" = queue.Queue()
            name_filter = None
            for term in (q.get("fieldSelector", [""])[0].split(",") if "fieldSelector" in q else []):
                k, _, v = term.partition("=")
                if k.strip() == "metadata.name":
                    name_filter = v.strip().lstrip("=")

            def wanted(o) -> bool:
                md = o.get("metadata", {})
                if ns and md.get("namespace", "") != ns:
                    return False
                if name_filter and md.get("name") != name_filter:
                    return False
                return not sel or all((md.get("labels") or {}).get(k) == v for k, v in sel.items())
            cancel = store.watch(res, lambda t, o, old: events.put((t, o)) if wanted(o) else None, replay=True)
            try:
                deadline = time.time() + float(q.get("timeoutSeconds", ["300"])[0])
                self.send_response(200)
                self.send_header("Content-Type", "application/json")
                self.send_header("Transfer-Encoding", "chunked")
                self.end_headers()
                while time.time() < deadline:
                    try:
                        t, o = events.get(timeout=min(1.0, max(0.05, deadline - time.time())))
                    except queue.Empty:
                        continue
                    line = (json.dumps({"type": t, "object": o}) + "\n").encode()
                    self.wfile.write(b"%x\r\n" % len(line) + line + b"\r\n")
                    self.wfile.flush()
                self.wfile.write(b"0\r\n\r\n")
            except (BrokenPipeError, ConnectionResetError, OSError):
                self.close_connection = True
            finally:
                cancel()

        def _discovery(self, parts):
            """API discovery documents (`/api`, `/apis`, `/api/v1`, `/apis/<group>/<version>`, `/openapi/v2`): what kubectl and
            client libraries read before they address a resource, so generic Kubernetes tooling can be pointed at the daemon."""
            verbs = ["create", "delete", "get", "list", "patch", "update", "watch"]
            groups = {}
            for (prefix, plural), res in API_GROUPS.items():
                groups.setdefault(prefix, []).append((plural, res))
            if parts == ["api"]:
                return {"kind": "APIVersions", "versions": ["v1"], "serverAddressByClientCIDRs": []}
            if parts == ["apis"]:
                out = []
                for prefix in sorted(groups):
                    if prefix.startswith("apis/"):
                        _, g, v = prefix.split("/")
                        gv = {"groupVersion": f"{g}/{v}", "version": v}
                        out.append({"name": g, "versions": [gv], "preferredVersion": gv})
                return {"kind": "APIGroupList", "apiVersion": "v1", "groups": out}
            if parts == ["openapi", "v2"]:
                from ..api.openapi import swagger
                return swagger()
            key = "/".join(parts)
            if key in groups:
                rl = []
                for plural, res in sorted(groups[key]):
                    _, kind, namespaced = RESOURCES[res]
                    short = {"mpijobs": ["mpijob", "mj"], "pods": ["po"], "services": ["svc"], "configmaps": ["cm"], "events": ["ev"]}.get(plural, [])
                    rl.append({"name": plural, "singularName": kind.lower(), "namespaced": namespaced, "kind": kind, "verbs": verbs, "shortNames": short})
                    if res in ("mpijobs", "pods", "jobs"):
                        rl.append({"name": plural + "/status", "singularName": "", "namespaced": namespaced, "kind": kind, "verbs": ["get", "patch", "update"]})
                    if res == "pods":
                        rl.append({"name": "pods/log", "singularName": "", "namespaced": True, "kind": "Pod", "verbs": ["get"]})
                        rl.append({"name": "nodes", "singularName": "node", "namespaced": False, "kind": "Node", "verbs": ["get", "list"],
                                   "shortNames": ["no"]})
                gv = key[len("apis/"):] if key.startswith("apis/") else "v1"
                return {"kind": "APIResourceList", "apiVersion": "v1", "groupVersion": gv, "resources": rl}
            return None

        def _follow_log(self, ns, name, q):
            """`pods/<name>/log?follow=true` (kubectl logs -f): chunks as the container writes them, until the pod has finished
            (Succeeded / Failed / deleted) and its log is drained, or `timeoutSeconds` (default 3600)."""
            deadline = time.time() + float(q.get("timeoutSeconds", ["3600"])[0])
            sent = 0
            try:
                self.send_response(200)
                self.send_header("Content-Type", "text/plain")
                self.send_header("Transfer-Encoding", "chunked")
                self.end_headers()
                while True:
                    piece, sent = op.agent.log_slice(ns, name, sent)   # only what is new (restart / rotation: starts over)
                    if piece:
                        self.wfile.write(b"%x\r\n" % len(piece) + piece + b"\r\n")
                        self.wfile.flush()
                        continue
                    try:
                        phase = (store.get("pods", ns, name).get("status") or {}).get("phase", "")
                    except errors.ApiError:
                        phase = "Deleted"
                    if phase in ("Succeeded", "Failed", "Deleted") or time.time() > deadline:
                        break
                    time.sleep(0.1)
                self.wfile.write(b"0\r\n\r\n")
            except (BrokenPipeError, ConnectionResetError, OSError):
                self.close_connection = True

        def _admit(self, res, obj):
            """API-server side admission for MPIJobs: CRD schema defaults + structural validation."""
            if res == "mpijobs":
                from ..api.schema import structural_errors
                # the reference's own SDK example passes `command="mpirun"` (sdk/python/v2beta1/tensorflow-mnist.py:32,48): a bare
                # string where core/v1 wants a list. Accepted the way a shell user means it - one word - instead of refused.
                specs = (obj.get("spec") or {}).get("mpiReplicaSpecs") if isinstance(obj.get("spec"), dict) else None
                for rs in (specs.values() if isinstance(specs, dict) else ()):
                    pod = ((rs.get("template") or {}).get("spec") if isinstance(rs, dict) and isinstance(rs.get("template"), dict) else None)
                    for c in (pod.get("containers") if isinstance(pod, dict) and isinstance(pod.get("containers"), list) else ()):
                        for k in ("command", "args"):
                            if isinstance(c, dict) and isinstance(c.get(k), str):
                                c[k] = [c[k]]
                errs = structural_errors(obj)
                if errs:
                    md = obj.get("metadata")
                    raise errors.invalid("MPIJob.kubeflow.org", str(md.get("name") or "") if isinstance(md, dict) else "",
                                         errs[0] if len(errs) == 1 else "[" + ", ".join(errs[:8]) + "]")
                if obj.get("kind", C.KIND) != C.KIND or obj.get("apiVersion", C.API_VERSION) != C.API_VERSION:
                    raise errors.invalid("mpijobs", M.name_of(obj), f"expected {C.API_VERSION}/{C.KIND}")
                if not (obj.get("spec") or {}).get("mpiReplicaSpecs"):
                    raise errors.invalid("mpijobs", M.name_of(obj), "spec.mpiReplicaSpecs: Required value")
            elif res == "pods":
                from ..api.schema import core_structural_errors
                errs = core_structural_errors(res, obj, required=True)
                if errs:
                    raise errors.invalid("pods", M.name_of(obj), errs[0] if len(errs) == 1 else "[" + ", ".join(errs[:8]) + "]")

        def do_POST(self):  # noqa: N802
            if self._refuse():
                return None
            parts, _ = self._route()
            try:
                r = self._resolve(parts)
                if r is None:
                    return self._send(404, errors.not_found("path", self.path).to_status())
                res, ns, _, _ = r
                obj = self._body()
                if ns:
                    M.meta(obj).setdefault("namespace", ns)
                self._admit(res, obj)
                return self._send(201, store.create(res, obj))
            except errors.ApiError as e:
                return self._send(e.code, e.to_status())

        def do_PUT(self):  # noqa: N802
            if self._refuse():
                return None
            parts, _ = self._route()
            try:
                r = self._resolve(parts)
                if r is None or not r[2]:
                    return self._send(404, errors.not_found("path", self.path).to_status())
                res, ns, name, sub = r
                obj = self._body()
                M.meta(obj)["name"] = name
                if ns:
                    M.meta(obj)["namespace"] = ns
                out = store.update_status(res, obj) if sub == "status" else store.update(res, obj)
                return self._send(200, out)
            except errors.ApiError as e:
                return self._send(e.code, e.to_status())

        def _node(self):
            """The box as a v1.Node (`kubectl get nodes` / `describe node`): capacity = discovered GPUs, allocatable = those not
            cordoned, the NVML facts as labels, cordons as taints. Read-only and synthesised on every request."""
            import socket
            topo, alloc = op.agent.topology, op.agent.alloc
            cordoned = alloc.cordoned
            n = topo.gpu_count
            name = socket.gethostname()
            gpu0 = topo.gpus[0] if topo.gpus else None
            labels = {"kubernetes.io/hostname": name, "b200mpi.kubeflow.org/topology-source": topo.source}
            if gpu0 is not None:
                labels["nvidia.com/gpu.product"] = (gpu0.name or "").replace(" ", "-")
                labels["nvidia.com/gpu.count"] = str(n)
                labels["nvidia.com/gpu.memory"] = str(gpu0.memory_bytes >> 20)
            return {"apiVersion": "v1", "kind": "Node",
                    "metadata": {"name": name, "uid": "node", "labels": labels,
                                 "annotations": {"b200mpi.kubeflow.org/cordoned-gpus": json.dumps({str(g): w for g, w in sorted(cordoned.items())}),
                                                 "b200mpi.kubeflow.org/free-gpus": str(alloc.free_gpus)}},
                    "spec": {"unschedulable": n > 0 and len(cordoned) == n,
                             "taints": [{"key": f"b200mpi.kubeflow.org/gpu-{g}", "value": why, "effect": "NoSchedule"} for g, why in sorted(cordoned.items())]},
                    "status": {"capacity": {"nvidia.com/gpu": str(n)}, "allocatable": {"nvidia.com/gpu": str(n - len(cordoned))},
                               "conditions": [{"type": "Ready", "status": "True", "reason": "NodeAgentRunning"}],
                               "nodeInfo": {"kubeletVersion": version.info().get("gitVersion", "") if isinstance(version.info(), dict) else ""}}}

        def _topology(self):
            t = op.agent.topology.to_dict()
            t["free_gpus"] = op.agent.alloc.free_gpus
            t["cordoned"] = {str(g): why for g, why in sorted(op.agent.alloc.cordoned.items())}
            return t

        def do_PATCH(self):  # noqa: N802
            if self._refuse():
                return None
            parts, _ = self._route()
            if parts == ["topology"]:
                # `kubectl cordon / uncordon` for GPUs: {"cordon": [3], "uncordon": [5], "reason": "..."}
                try:
                    body = self._body()
                    for g in body.get("cordon", []) or []:
                        op.agent.alloc.cordon(int(g), str(body.get("reason") or "cordoned by the operator"))
                    for g in body.get("uncordon", []) or []:
                        op.agent.alloc.uncordon(int(g))
                except (ValueError, TypeError) as e:
                    return self._send(422, {"kind": "Status", "status": "Failure", "reason": "Invalid", "code": 422, "message": str(e)})
                metrics.gpu_slots_free.set(op.agent.alloc.free_gpus)
                op.agent.wake()
                return self._send(200, self._topology())
            try:
                r = self._resolve(parts)
                if r is None or not r[2]:
                    return self._send(404, errors.not_found("path", self.path).to_status())
                res, ns, name, sub = r
                return self._send(200, store.patch(res, ns, name, self._body(), status=(sub == "status")))
            except errors.ApiError as e:
                return self._send(e.code, e.to_status())

        def do_DELETE(self):  # noqa: N802
            if self._refuse():
                return None
            parts, _ = self._route()
            try:
                r = self._resolve(parts)
                if r is None:
                    return self._send(404, errors.not_found("path", self.path).to_status())
                res, ns, name, _ = r
                if name:
                    return self._send(200, store.delete(res, ns, name))
                n = store.delete_collection(res, ns or None)
                return self._send(200, {"kind": "Status", "status": "Success", "details": {"deleted": n}})
            except errors.ApiError as e:
                return self._send(e.code, e.to_status())

    def guarded(fn):
        """No request may end in a dropped connection: API errors become their Status, a body that does not have the shape of
        the object it claims to be (a string where a map belongs ...) is a 400 like the apiserver's decoder gives, anything
        else a 500 with the exception's text; the traceback goes to the log."""
        def wrapper(self):
            try:
                return fn(self)
            except errors.ApiError as e:
                return self._send(e.code, e.to_status())
            except (BrokenPipeError, ConnectionResetError):
                self.close_connection = True
            except (AttributeError, TypeError, KeyError, IndexError, ValueError) as e:
                log.debug("http: malformed request %s %s", self.command, self.path, exc_info=True)
                return self._send(400, errors.bad_request(f"the request could not be decoded: {type(e).__name__}: {e}").to_status())
            except Exception as e:  # noqa: BLE001
                log.exception("http: %s %s failed", self.command, self.path)
                return self._send(500, errors.ApiError("InternalError", f"{type(e).__name__}: {e}", 500).to_status())
        wrapper.__name__ = fn.__name__
        return wrapper

    for verb in ("do_GET", "do_POST", "do_PUT", "do_PATCH", "do_DELETE"):
        setattr(H, verb, guarded(getattr(H, verb)))
    return H


def run(opt: ServerOption) -> int:
    """app.Run (server.go:79-256)."""
    if opt.print_version:
        version.print_version_and_exit()
    logging.basicConfig(level=logging.DEBUG if opt.verbosity >= 4 else logging.INFO,
                        format="%(levelname).1s%(asctime)s %(name)s] %(message)s", datefmt="%m%d %H:%M:%S")
    log.info("%s", version.info())
    log.info("Server options: %s", opt)
    if opt.namespace == "":
        log.info("Using cluster scoped operator")
    else:
        log.info("Scoping operator to namespace %s", opt.namespace)
    stop = threading.Event()
    signal.signal(signal.SIGTERM, lambda *_: stop.set())
    signal.signal(signal.SIGINT, lambda *_: stop.set())
    op = Operator(opt)
    op.serve(opt.listen)
    log.info("REST API on http://%s", opt.listen)
    if opt.healthz_port:
        try:
            op.serve(f"127.0.0.1:{opt.healthz_port}", restricted=True)
        except OSError as e:
            log.warning("healthz port %d unavailable: %s", opt.healthz_port, e)
    if opt.monitoring_port:
        op.serve(f"0.0.0.0:{opt.monitoring_port}", restricted=True)   # Prometheus scrape port: /metrics + /healthz only
    op.start()
    stop.wait()
    op.stop()
    return 0
