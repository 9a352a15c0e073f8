"""Mutation fuzzer for the daemon's REST surface (run from the repo root: `python tools/fuzz_rest_api.py`, SEED=n for other streams).

Starts an in-process operator with fake GPUs, POSTs / PUTs / PATCHes mutated MPIJobs and pods (fields replaced by junk of the
wrong type, fields removed, malformed transport-level bodies, odd query strings) and reports: responses >= 500 or dropped
connections (must be 0), the phases the accepted pods reached (all terminal: nothing may wedge the node agent), and every
warning the controller or the node agent logged. The findings of its first runs are the regression tests
`test_sdk_server.py::test_malformed_requests_get_a_status_never_a_dropped_connection` and
`test_integration.py::test_a_pod_the_agent_cannot_digest_fails_alone`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import logging; logging.basicConfig(level=logging.WARNING)
import time
import json, random, socket, sys, copy, tempfile, urllib.request, urllib.error, logging

from mpi_operator_b200.cmd.options import ServerOption
from mpi_operator_b200.cmd.server import Operator
s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
op = Operator(ServerOption(fake_gpus=2, leader_elect=False, state_dir=tempfile.mkdtemp()))
op.serve(f"127.0.0.1:{port}"); op.start()
base = f"http://127.0.0.1:{port}"
good = {"apiVersion": "kubeflow.org/v2beta1", "kind": "MPIJob", "metadata": {"name": "f"},
        "spec": {"slotsPerWorker": 1, "runPolicy": {"cleanPodPolicy": "Running", "backoffLimit": 1},
                 "mpiReplicaSpecs": {"Launcher": {"replicas": 1, "template": {"spec": {"containers": [{"name": "l", "command": ["true"]}]}}},
                                     "Worker": {"replicas": 1, "template": {"spec": {"containers": [{"name": "w", "command": ["true"], "resources": {"limits": {"nvidia.com/gpu": 1}}}]}}}}}}
junk = [None, 0, -1, 1.5, "", "x", [], {}, [1], {"a": 1}, True, "9" * 300, 2**70, float("1e308"), [[]], {"": None}]
import os; rnd = random.Random(int(os.environ.get("SEED", 7)))
def paths(o, p=()):
    yield p
    if isinstance(o, dict):
        for k, v in o.items(): yield from paths(v, p + (k,))
    elif isinstance(o, list):
        for i, v in enumerate(o): yield from paths(v, p + (i,))
def setp(o, p, v):
    for k in p[:-1]: o = o[k]
    o[p[-1]] = v
def delp(o, p):
    for k in p[:-1]: o = o[k]
    del o[p[-1]]
def call(method, path, body=None, raw=None, ctype="application/json"):
    data = raw if raw is not None else (json.dumps(body).encode() if body is not None else None)
    req = urllib.request.Request(base + path, data=data, method=method, headers={"Content-Type": ctype})
    try:
        with urllib.request.urlopen(req, timeout=10) as r: return r.status, r.read()
    except urllib.error.HTTPError as e: return e.code, e.read()
    except Exception as e: return 599, repr(e).encode()
bad = []
allp = [p for p in paths(good) if p]
coll = "/apis/kubeflow.org/v2beta1/namespaces/default/mpijobs"
n = 0
for it in range(400):
    b = copy.deepcopy(good)
    b["metadata"]["name"] = f"f{it}"
    for _ in range(rnd.choice([1, 1, 2, 3])):
        p = rnd.choice(allp)
        try:
            if rnd.random() < 0.25: delp(b, p)
            else: setp(b, p, copy.deepcopy(rnd.choice(junk)))
        except (KeyError, IndexError, TypeError): pass
    try:
        body = json.dumps(b)
    except (TypeError, ValueError):
        continue
    st, out = call("POST", coll, raw=body.encode())
    n += 1
    if st >= 500: bad.append(("POST", st, body[:300], out[:300]))
    if st < 300:
        name = b["metadata"]["name"] if isinstance(b.get("metadata"), dict) else None
        if isinstance(name, str) and name:
            # mutate via PUT / PATCH
            for m in ("PUT", "PATCH"):
                c = copy.deepcopy(b)
                p = rnd.choice(allp)
                try: setp(c, p, copy.deepcopy(rnd.choice(junk))); json.dumps(c)
                except (KeyError, IndexError, TypeError, ValueError): continue
                st2, out2 = call(m, f"{coll}/{name}", raw=json.dumps(c).encode(), ctype="application/merge-patch+json" if m == "PATCH" else "application/json")
                if st2 >= 500: bad.append((m, st2, json.dumps(c)[:300], out2[:300]))
            time.sleep(0.15)
            call("DELETE", f"{coll}/{name}")
# malformed transport-level bodies and odd paths
for raw in [b"", b"{", b"[]", b"null", b"\xff\xfe", b'"str"', b"1", b'{"metadata": 5}', b"{" * 10000]:
    for m, path in (("POST", coll), ("PUT", coll + "/x"), ("PATCH", coll + "/x"), ("PATCH", "/topology"), ("POST", "/api/v1/namespaces/default/pods")):
        st, out = call(m, path, raw=raw)
        if st >= 500: bad.append((m, path, st, raw[:40], out[:300]))
for path in ["/apis/kubeflow.org/v2beta1/namespaces/default/mpijobs?limit=abc", "/apis/kubeflow.org/v2beta1/namespaces/default/mpijobs?labelSelector=%3D%3D%3D",
             "/apis/kubeflow.org/v2beta1/namespaces/default/mpijobs?watch=true&timeoutSeconds=x", "/api/v1/namespaces/default/pods/nope/log?tailLines=zz",
             "/apis/kubeflow.org/v2beta1/namespaces//mpijobs", "/apis/kubeflow.org/v9/namespaces/default/mpijobs", "/api/v1/nodes/%00", "/" + "a" * 5000,
             "/apis/kubeflow.org/v2beta1/namespaces/default/mpijobs?fieldSelector=metadata.name", "/api/v1/namespaces/default/events?resourceVersion=-5"]:
    st, out = call("GET", path)
    if st >= 500: bad.append(("GET", path[:80], st, out[:300]))
podcoll = "/api/v1/namespaces/default/pods"
goodpod = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p"}, "spec": {"containers": [{"name": "c", "command": ["true"], "resources": {"limits": {"nvidia.com/gpu": 1}}, "env": [{"name": "A", "value": "b"}], "volumeMounts": []}], "volumes": [], "restartPolicy": "Never"}}
podpaths = [p for p in paths(goodpod) if p]
for it in range(150):
    b = copy.deepcopy(goodpod); b["metadata"]["name"] = f"p{it}"
    p = rnd.choice(podpaths)
    try:
        if rnd.random() < 0.3: delp(b, p)
        else: setp(b, p, copy.deepcopy(rnd.choice(junk)))
        body = json.dumps(b)
    except (KeyError, IndexError, TypeError, ValueError): continue
    st, out = call("POST", podcoll, raw=body.encode())
    if st >= 500: bad.append(("POST pod", st, body[:300], out[:300]))
time.sleep(3)
st, out = call("GET", podcoll)
pods = json.loads(out)["items"]
from collections import Counter
print("pod phases", Counter((p.get("status") or {}).get("phase") for p in pods))
print("requests", n, "server errors", len(bad))
for b in bad[:15]: print(b)
op.stop()
