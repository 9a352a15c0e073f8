"""mpijobctl — kubectl-shaped CLI for the single-box MPIJob operator.

    mpijobctl apply -f examples/pi/pi.yaml        # create or update
    mpijobctl get mpijobs | pods | jobs | events  # list
    mpijobctl describe mpijob pi                  # spec summary, status, conditions, events
    mpijobctl logs pi [--worker N]                # launcher (or a pod's) output
    mpijobctl scale pi --replicas 8               # elastic rescale
    mpijobctl suspend|resume pi                   # runPolicy.suspend
    mpijobctl wait pi --for Succeeded --timeout 300
    mpijobctl delete mpijob pi
    mpijobctl topology
    mpijobctl run -f job.yaml [--gpus N]          # no daemon: in-process operator, wait, print logs

Talks to the daemon's REST API (``--server host:port`` or $MPIJOB_SERVER), the
local replacement for ``kubectl`` + kube-apiserver in the reference workflow
(README.md:63-170).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import Dict, List, Optional

import yaml

from ..api import constants as C
from ..sdk.client import MPIJobClient, ApiException

KIND_ALIASES = {"mpijob": "mpijobs", "mpijobs": "mpijobs", "mj": "mpijobs", "pod": "pods", "pods": "pods", "po": "pods",
                "job": "jobs", "jobs": "jobs", "svc": "services", "service": "services", "services": "services",
                "cm": "configmaps", "configmap": "configmaps", "configmaps": "configmaps", "secret": "secrets",
                "secrets": "secrets", "event": "events", "events": "events", "ev": "events", "podgroup": "podgroups",
                "podgroups": "podgroups", "pg": "podgroups", "lease": "leases", "leases": "leases", "node": "nodes", "nodes": "nodes",
                "no": "nodes"}


def _age(ts: Optional[str]) -> str:
    if not ts:
        return "<none>"
    from ..api.meta import parse_rfc3339
    d = int(time.time() - parse_rfc3339(ts))
    return f"{d}s" if d < 120 else (f"{d // 60}m" if d < 7200 else f"{d // 3600}h")


def _job_state(j: dict) -> str:
    conds = j.get("status", {}).get("conditions", []) or []
    for t in ("Failed", "Succeeded", "Suspended", "Running", "Created"):
        if any(c["type"] == t and c["status"] == "True" for c in conds):
            return t
    return "Pending"


def cmd_apply(cli: MPIJobClient, a) -> int:
    rc = 0
    for path in a.filename:
        text = sys.stdin.read() if path == "-" else open(path).read()
        for doc in yaml.safe_load_all(text):
            if not doc:
                continue
            doc.setdefault("metadata", {}).setdefault("namespace", a.namespace)
            if doc.get("kind") != C.KIND:
                print(f"skipping {doc.get('kind')}: only MPIJob documents are applied", file=sys.stderr)
                continue
            try:
                out, verb = cli.apply(doc)
                print(f"mpijob.kubeflow.org/{out['metadata']['name']} {verb}")
            except ApiException as e:
                print(f"error: {e}", file=sys.stderr)
                rc = 1
    return rc


def cmd_get(cli: MPIJobClient, a) -> int:
    res = KIND_ALIASES.get(a.kind.lower())
    if res is None:
        print(f"error: unknown resource type {a.kind!r}", file=sys.stderr)
        return 1
    try:
        items = [cli.get_resource(res, a.namespace, a.name)] if a.name else cli.list_resource(res, None if a.all_namespaces else a.namespace)
    except ApiException as e:
        print(f"Error from server ({e.reason}): {e}", file=sys.stderr)
        return 1
    if a.output == "json":
        print(json.dumps(items[0] if a.name else {"items": items}, indent=2))
        return 0
    if a.output == "yaml":
        print(yaml.safe_dump(items[0] if a.name else {"items": items}, sort_keys=False))
        return 0
    if res == "mpijobs":
        print(f"{'NAME':24} {'STATE':10} {'WORKERS':8} {'AGE':6}")
        for j in items:
            w = ((j.get("spec", {}).get("mpiReplicaSpecs") or {}).get("Worker") or {}).get("replicas", 0)
            print(f"{j['metadata']['name']:24} {_job_state(j):10} {str(w):8} {_age(j['metadata'].get('creationTimestamp')):6}")
    elif res == "pods":
        print(f"{'NAME':36} {'STATUS':10} {'RESTARTS':8} {'GPUS':10} {'AGE':6}")
        for p in items:
            cs = (p.get("status", {}).get("containerStatuses") or [{}])[0]
            gp = (p["metadata"].get("annotations") or {}).get("b200mpi.kubeflow.org/gpus", "")
            print(f"{p['metadata']['name']:36} {p.get('status', {}).get('phase', 'Pending'):10} {str(cs.get('restartCount', 0)):8} {gp or '-':10} {_age(p['metadata'].get('creationTimestamp')):6}")
    elif res == "nodes":
        print(f"{'NAME':24} {'STATUS':26} {'GPUS':6} {'ALLOCATABLE':12} {'FREE':6} CORDONED")
        for nd in items:
            st = "Ready" + (",SchedulingDisabled" if nd.get("spec", {}).get("unschedulable") else "")
            ann = nd["metadata"].get("annotations", {})
            cord = json.loads(ann.get("b200mpi.kubeflow.org/cordoned-gpus", "{}"))
            print(f"{nd['metadata']['name']:24} {st:26} {nd['status']['capacity'].get('nvidia.com/gpu', '0'):6} "
                  f"{nd['status']['allocatable'].get('nvidia.com/gpu', '0'):12} {ann.get('b200mpi.kubeflow.org/free-gpus', '?'):6} "
                  f"{','.join(sorted(cord, key=int)) or '-'}")
    elif res == "events":
        print(f"{'LAST SEEN':10} {'TYPE':8} {'REASON':28} {'OBJECT':28} MESSAGE")
        for e in sorted(items, key=lambda e: e.get("lastTimestamp", "")):
            io = e.get("involvedObject", {})
            print(f"{_age(e.get('lastTimestamp')):10} {e.get('type', ''):8} {e.get('reason', ''):28} {(io.get('kind', '') + '/' + io.get('name', '')).lower():28} {e.get('message', '')}")
    else:
        print(f"{'NAME':40} {'AGE':6}")
        for o in items:
            print(f"{o['metadata']['name']:40} {_age(o['metadata'].get('creationTimestamp')):6}")
    return 0


def cmd_get_watch(cli: MPIJobClient, a) -> int:
    """`get <kind> -w`: the table once, then one line per event of the server's `?watch=true` stream — kubectl get -w."""
    rc = cmd_get(cli, a)
    if rc:
        return rc
    res = KIND_ALIASES.get(a.kind.lower())
    ns = None if a.all_namespaces else a.namespace
    try:
        listed = [cli.get_resource(res, a.namespace, a.name)] if a.name else cli.list_resource(res, ns)
        shown = {(o["metadata"].get("namespace", ""), o["metadata"]["name"]): o["metadata"].get("resourceVersion") for o in listed}
        for ev in cli.watch(res, ns, timeout=a.watch_timeout or 3600.0, name=a.name):
            o = ev["object"]
            key = (o["metadata"].get("namespace", ""), o["metadata"]["name"])
            if ev["type"] == "ADDED" and shown.get(key) == o["metadata"].get("resourceVersion"):
                continue          # the stream first replays what the table already showed
            state = "Deleted" if ev["type"] == "DELETED" else (_job_state(o) if res == "mpijobs" else (o.get("status", {}).get("phase", "") if res == "pods" else ""))
            print(f"{o['metadata']['name']:36} {state:10} {_age(o['metadata'].get('creationTimestamp')):6}", flush=True)
    except KeyboardInterrupt:
        pass
    return 0


def cmd_describe(cli: MPIJobClient, a) -> int:
    try:
        j = cli.get(a.name, a.namespace)
    except ApiException as e:
        print(f"Error from server ({e.reason}): {e}", file=sys.stderr)
        return 1
    md, spec, st = j["metadata"], j.get("spec", {}), j.get("status", {})
    print(f"Name:         {md['name']}\nNamespace:    {md.get('namespace', '')}\nAPI Version:  {j.get('apiVersion')}\nKind:         {j.get('kind')}")
    print(f"UID:          {md.get('uid', '')}\nCreated:      {md.get('creationTimestamp', '')}")
    print("Spec:")
    for k in ("slotsPerWorker", "mpiImplementation", "launcherCreationPolicy", "runLauncherAsWorker", "sshAuthMountPath"):
        if k in spec:
            print(f"  {k}: {spec[k]}")
    print(f"  runPolicy: {json.dumps(spec.get('runPolicy', {}))}")
    for rt, rs in (spec.get("mpiReplicaSpecs") or {}).items():
        c0 = ((rs.get("template", {}).get("spec", {}).get("containers")) or [{}])[0]
        print(f"  {rt}: replicas={rs.get('replicas')} restartPolicy={rs.get('restartPolicy', '')} "
              f"command={' '.join((c0.get('command') or []) + (c0.get('args') or []))!r}")
    print("Status:")
    for k in ("startTime", "completionTime"):
        if st.get(k):
            print(f"  {k}: {st[k]}")
    print(f"  replicaStatuses: {json.dumps(st.get('replicaStatuses', {}))}")
    print("  Conditions:")
    print(f"    {'TYPE':10} {'STATUS':7} {'REASON':32} MESSAGE")
    for c in st.get("conditions", []) or []:
        print(f"    {c['type']:10} {c['status']:7} {c.get('reason', ''):32} {c.get('message', '')}")
    print("Events:")
    for e in sorted(cli.list_resource("events", a.namespace), key=lambda e: e.get("lastTimestamp", "")):
        if e.get("involvedObject", {}).get("name") == a.name:
            print(f"  {e.get('type', ''):8} {e.get('reason', ''):28} x{e.get('count', 1):<3} {e.get('message', '')}")
    return 0


def cmd_delete(cli: MPIJobClient, a) -> int:
    res = KIND_ALIASES.get(a.kind.lower(), a.kind)
    try:
        cli.delete_resource(res, a.namespace, a.name)
        print(f"{a.kind}/{a.name} deleted")
        return 0
    except ApiException as e:
        print(f"Error from server ({e.reason}): {e}", file=sys.stderr)
        return 1


def cmd_logs(cli: MPIJobClient, a) -> int:
    """`logs <job>` prints the launcher's log (where mpirun multiplexes every rank's output); `-f` keeps printing what is
    appended until the job finishes (kubectl logs -f)."""
    import time
    printed = 0
    deadline = time.time() + a.timeout if getattr(a, "timeout", None) else None
    while True:
        try:
            text = cli.logs(a.name, a.namespace, worker=a.worker, pod=a.pod, tail=None if getattr(a, "follow", False) else getattr(a, "tail", None))
        except ApiException as e:
            if not getattr(a, "follow", False) or printed:
                print(f"error: {e}", file=sys.stderr)
                return 1
            text = ""          # -f before the launcher pod exists: keep waiting
        if len(text) < printed:   # the container restarted: its log starts over
            printed = 0
        sys.stdout.write(text[printed:])
        sys.stdout.flush()
        printed = len(text)
        if not getattr(a, "follow", False):
            return 0
        try:
            done = _job_state(cli.get(a.name, a.namespace)) in ("Succeeded", "Failed")
        except ApiException:
            done = True
        if done:
            tail = cli.logs(a.name, a.namespace, worker=a.worker, pod=a.pod) if printed else ""
            sys.stdout.write(tail[printed:])
            return 0
        if deadline and time.time() > deadline:
            return 1
        time.sleep(0.25)


def cmd_scale(cli: MPIJobClient, a) -> int:
    cli.patch(a.name, {"spec": {"mpiReplicaSpecs": {"Worker": {"replicas": a.replicas}}}}, a.namespace)
    print(f"mpijob.kubeflow.org/{a.name} scaled")
    return 0


def cmd_suspend(cli: MPIJobClient, a, value: bool) -> int:
    cli.patch(a.name, {"spec": {"runPolicy": {"suspend": value}}}, a.namespace)
    print(f"mpijob.kubeflow.org/{a.name} {'suspended' if value else 'resumed'}")
    return 0


def cmd_patch(cli: MPIJobClient, a) -> int:
    """`kubectl patch mpijob NAME --type merge -p '{"spec": {"runPolicy": {"suspend": true}}}'` (how Kueue-style controllers and
    scripts flip fields of a running job); only merge patches exist here."""
    if a.kind.lower() not in ("mpijob", "mpijobs", "mpijob.kubeflow.org"):
        print("error: only mpijobs can be patched through the CLI", file=sys.stderr)
        return 2
    if a.type not in ("merge", "strategic"):
        print(f"error: patch type {a.type!r} is not supported (merge only)", file=sys.stderr)
        return 2
    try:
        body = yaml.safe_load(a.patch) if a.patch is not None else yaml.safe_load(open(a.patch_file).read())
    except Exception as e:  # noqa: BLE001
        print(f"error: cannot parse the patch: {e}", file=sys.stderr)
        return 2
    if not isinstance(body, dict):
        print("error: the patch must be a JSON / YAML object", file=sys.stderr)
        return 2
    cli.patch(a.name, body, a.namespace)
    print(f"mpijob.kubeflow.org/{a.name} patched")
    return 0


def cmd_meta(cli: MPIJobClient, a, field: str) -> int:
    """`kubectl label / annotate mpijob NAME key=value ... key-` (a trailing dash removes the key)."""
    changes = {}
    for kv in a.pairs:
        if kv.endswith("-") and "=" not in kv:
            changes[kv[:-1]] = None
        elif "=" in kv:
            k, v = kv.split("=", 1)
            changes[k] = v
        else:
            print(f"error: expected key=value or key-, got {kv!r}", file=sys.stderr)
            return 2
    if not a.overwrite:
        cur = (cli.get(a.name, a.namespace).get("metadata", {}).get(field) or {})
        clash = [k for k, v in changes.items() if v is not None and k in cur and cur[k] != v]
        if clash:
            print(f"error: {field[:-1]} {clash[0]!r} already has a value ({cur[clash[0]]}), and --overwrite is false", file=sys.stderr)
            return 1
    cli.patch(a.name, {"metadata": {field: changes}}, a.namespace)
    print(f"mpijob.kubeflow.org/{a.name} {'labeled' if field == 'labels' else 'annotated'}")
    return 0


def cmd_wait(cli: MPIJobClient, a) -> int:
    try:
        j = cli.wait_for_condition(a.name, a.condition, a.namespace, timeout=a.timeout)
        print(f"mpijob.kubeflow.org/{a.name} condition met: {_job_state(j)}")
        return 0
    except TimeoutError as e:
        print(f"error: {e}", file=sys.stderr)
        return 1


def cmd_run(a) -> int:
    """Standalone: in-process operator, apply, wait for completion, print launcher logs."""
    import logging
    from ..cmd.options import ServerOption
    from ..cmd.server import Operator
    from ..api import yaml_io
    logging.basicConfig(level=logging.INFO if a.verbose else logging.WARNING)
    op = Operator(ServerOption(fake_gpus=a.fake_gpus, leader_elect=False, gang_scheduling_name=a.gang_scheduling or ""))
    op.start()
    rc = 0
    try:
        for path in a.filename:
            for job in yaml_io.load_file(path):
                job.metadata.setdefault("namespace", a.namespace)
                if a.replicas is not None and job.spec.replica("Worker") is not None:
                    job.spec.replica("Worker").replicas = a.replicas
                if a.np is not None:  # rewrite the launcher's -np/-n value (scaling the same YAML)
                    c0 = job.spec.replica("Launcher").main_container()
                    for key in ("command", "args"):
                        argv = c0.get(key) or []
                        for i, tok in enumerate(argv[:-1]):
                            if tok in ("-np", "-n", "--np", "--n"):
                                argv[i + 1] = str(a.np)
                c = op.clientset.kubeflow_v2beta1().mpijobs(job.namespace)
                c.create(job)
                t0 = time.time()
                state = "Pending"
                printed: Dict[str, int] = {}  # launcher pod -> bytes of its log already written

                def drain() -> None:
                    # The launcher pod is deleted when its Job hits the backoff limit (batch/v1 semantics with
                    # restartPolicy OnFailure), so the log is copied out while the pod still exists.
                    for p in op.store.list("pods", job.namespace):
                        md = p["metadata"]
                        if md.get("labels", {}).get(C.JOB_ROLE_LABEL) != "launcher" or md.get("labels", {}).get(C.JOB_NAME_LABEL, job.name) != job.name:
                            continue
                        while True:   # only the new bytes (a restarted / rotated log starts over at 0)
                            piece, printed[md["name"]] = op.agent.log_slice(job.namespace, md["name"], printed.get(md["name"], 0))
                            if not piece:
                                break
                            sys.stdout.write(piece.decode(errors="replace"))
                            sys.stdout.flush()

                last_drain = 0.0
                while time.time() - t0 < a.timeout:
                    j = c.get(job.name).to_dict()
                    state = _job_state(j)
                    if time.time() - last_drain > 0.5:
                        drain()
                        last_drain = time.time()
                    if state in ("Succeeded", "Failed"):
                        break
                    time.sleep(0.05)
                drain()
                print(f"mpijob.kubeflow.org/{job.name}: {state} after {time.time() - t0:.2f}s")
                rc = rc or (0 if state == "Succeeded" else 1)
    finally:
        op.stop()
    return rc


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="mpijobctl")
    ap.add_argument("--server", default=os.environ.get("MPIJOB_SERVER", "127.0.0.1:8087"))
    ap.add_argument("--token-file", default=os.environ.get("MPIJOB_TOKEN_FILE", ""),
                    help="bearer token of a daemon started with --auth-token-file (also: $MPIJOB_TOKEN)")
    ap.add_argument("-n", "--namespace", default="default")
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name in ("apply", "create"):
        p = sub.add_parser(name)
        p.add_argument("-f", "--filename", action="append", required=True)
    p = sub.add_parser("get")
    p.add_argument("kind")
    p.add_argument("name", nargs="?")
    p.add_argument("-o", "--output", default="")
    p.add_argument("-A", "--all-namespaces", action="store_true")
    p.add_argument("-w", "--watch", action="store_true", help="after listing, print a line whenever an object changes")
    p.add_argument("--watch-timeout", type=float, default=None, help="with -w: stop after this many seconds")
    p = sub.add_parser("describe")
    p.add_argument("kind", nargs="?", default="mpijob")
    p.add_argument("name")
    p = sub.add_parser("delete")
    p.add_argument("kind")
    p.add_argument("name")
    p = sub.add_parser("logs")
    p.add_argument("name")
    p.add_argument("--worker", type=int, default=None)
    p.add_argument("--pod", default=None)
    p.add_argument("-f", "--follow", action="store_true", help="stream the log until the job finishes")
    p.add_argument("--tail", type=int, default=None, help="only the last N lines")
    p.add_argument("--timeout", type=float, default=None, help="with -f: give up after this many seconds")
    p = sub.add_parser("scale")
    p.add_argument("name")
    p.add_argument("--replicas", type=int, required=True)
    for name in ("suspend", "resume"):
        p = sub.add_parser(name)
        p.add_argument("name")
    p = sub.add_parser("patch")
    p.add_argument("kind")
    p.add_argument("name")
    p.add_argument("--type", default="merge")
    g = p.add_mutually_exclusive_group(required=True)
    g.add_argument("-p", "--patch", default=None)
    g.add_argument("--patch-file", default=None)
    for name in ("label", "annotate"):
        p = sub.add_parser(name)
        p.add_argument("kind")
        p.add_argument("name")
        p.add_argument("pairs", nargs="+", help="key=value ... (key- removes)")
        p.add_argument("--overwrite", action="store_true")
    p = sub.add_parser("wait")
    p.add_argument("name")
    p.add_argument("--for", dest="condition", default="Succeeded")
    p.add_argument("--timeout", type=float, default=300)
    sub.add_parser("topology")
    for name in ("cordon", "uncordon"):
        p = sub.add_parser(name, help="take a GPU out of / back into scheduling (running jobs keep it)")
        p.add_argument("gpu", type=int, nargs="+")
        if name == "cordon":
            p.add_argument("--reason", default="")
    sub.add_parser("version")
    p = sub.add_parser("run")
    p.add_argument("-f", "--filename", action="append", required=True)
    p.add_argument("--timeout", type=float, default=600)
    p.add_argument("--fake-gpus", type=int, default=None)
    p.add_argument("--replicas", type=int, default=None)
    p.add_argument("--np", type=int, default=None, help="rewrite the launcher mpirun -np value")
    p.add_argument("--gang-scheduling", default="")
    p.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args(argv)
    if a.cmd == "run":
        return cmd_run(a)
    if a.cmd == "version":
        from .. import version
        print(json.dumps(version.info()))
        return 0
    if a.token_file:
        os.environ["MPIJOB_TOKEN_FILE"] = a.token_file   # picked up by the SDK's Configuration
    cli = MPIJobClient(a.server)
    if a.cmd in ("apply", "create"):
        return cmd_apply(cli, a)
    if a.cmd == "get":
        return cmd_get_watch(cli, a) if a.watch else cmd_get(cli, a)
    if a.cmd == "describe":
        return cmd_describe(cli, a)
    if a.cmd == "delete":
        return cmd_delete(cli, a)
    if a.cmd == "logs":
        return cmd_logs(cli, a)
    if a.cmd == "scale":
        return cmd_scale(cli, a)
    if a.cmd == "suspend":
        return cmd_suspend(cli, a, True)
    if a.cmd == "resume":
        return cmd_suspend(cli, a, False)
    if a.cmd == "patch":
        return cmd_patch(cli, a)
    if a.cmd in ("label", "annotate"):
        return cmd_meta(cli, a, "labels" if a.cmd == "label" else "annotations")
    if a.cmd == "wait":
        return cmd_wait(cli, a)
    if a.cmd == "topology":
        print(json.dumps(cli.raw_get("/topology"), indent=2))
        return 0
    if a.cmd in ("cordon", "uncordon"):
        body = {a.cmd: a.gpu}
        if getattr(a, "reason", ""):
            body["reason"] = a.reason
        try:
            t = cli.api.call_api("/topology", "PATCH", body=body)
        except ApiException as e:
            try:
                msg = json.loads(e.body).get("message", str(e))
            except Exception:  # noqa: BLE001
                msg = str(e)
            print(f"error: {msg}", file=sys.stderr)
            return 1
        for g in a.gpu:
            print(f"gpu/{g} {'cordoned' if a.cmd == 'cordon' else 'uncordoned'}")
        print(f"free GPUs: {t['free_gpus']}; cordoned: {t['cordoned'] or 'none'}")
        return 0
    return 2


if __name__ == "__main__":
    sys.exit(main())
