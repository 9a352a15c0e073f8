#!/usr/bin/env python
"""Control-plane benchmark (no GPU): how fast the operator turns MPIJobs into their dependent objects and how long a
CPU job takes from `create` to `Succeeded`.

The reference publishes no control-plane numbers; its knobs are 2 reconcile threads, workqueue 10 qps / burst 100 and
API 5 qps / burst 10 (cmd/mpi-operator/app/options/options.go:73,87-91), its sample job takes 75 s from creation to
completion on a cluster (README.md:145-169) and the e2e suite allows 200 s per wait (test/e2e/e2e_suite_test.go:72-73).

  part 1  N suspended-free MPIJobs x W workers with virtual (sshd) workers and a trivial launcher: time until every job has
          its Service / ConfigMap / Secret / W worker pods / launcher Job and reports Running or Succeeded
  part 2  the pi example end to end (2 ranks through the native mpirun + libmpi shim): create -> Succeeded latency
"""
import argparse
import json
import os
import statistics
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import conds, new_mpijob  # noqa: E402
from mpi_operator_b200.api import yaml_io  # noqa: E402
from mpi_operator_b200.cmd.options import ServerOption  # noqa: E402
from mpi_operator_b200.cmd.server import Operator  # noqa: E402


def wait_all(fn, names, timeout):
    t0 = time.time()
    pending = set(names)
    done_at = {}
    while pending and time.time() - t0 < timeout:
        for n in list(pending):
            if fn(n):
                pending.discard(n)
                done_at[n] = time.time() - t0
        time.sleep(0.01)
    if pending:
        raise SystemExit(f"timed out; {len(pending)} jobs not ready: {sorted(pending)[:5]}")
    return done_at


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=50)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--threadiness", type=int, default=2)
    ap.add_argument("--pi-runs", type=int, default=5)
    ap.add_argument("--queue-qps", type=int, default=10, help="--controller-queue-rate-limit (reference default 10)")
    ap.add_argument("--queue-burst", type=int, default=100, help="--controller-queue-burst (reference default 100)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    out = {"jobs": a.jobs, "workers_per_job": a.workers, "threadiness": a.threadiness, "queue_qps": a.queue_qps, "queue_burst": a.queue_burst}

    with tempfile.TemporaryDirectory() as d:
        op = Operator(ServerOption(fake_gpus=0, leader_elect=False, state_dir=d, threadiness=a.threadiness,
                                   controller_rate_limit=a.queue_qps, controller_burst=a.queue_burst))
        op.start()
        try:
            cs = op.clientset.kubeflow_v2beta1().mpijobs("default")
            names = [f"bench-{i}" for i in range(a.jobs)]
            t0 = time.time()
            for n in names:
                cs.create(new_mpijob(n, workers=a.workers, launcher_cmd=("true",), launcher_args=None, worker_cmd=("/usr/sbin/sshd", "-De")))
            t_submit = time.time() - t0

            def materialised(n):
                c = conds(cs.get(n))
                return c.get("Running") == "True" or c.get("Succeeded") == "True"
            done_at = wait_all(materialised, names, 300)
            t_all = max(done_at.values()) + t_submit
            objs = sum(len(op.store.list(r, "default")) for r in ("pods", "services", "configmaps", "secrets", "jobs"))
            out["part1"] = {"submit_s": round(t_submit, 3), "all_running_s": round(t_all, 3),
                            "jobs_per_s": round(a.jobs / t_all, 1), "dependent_objects": objs,
                            "objects_per_s": round(objs / t_all, 1),
                            "per_job_latency_ms_p50": round(1000 * statistics.median(done_at.values()), 1),
                            "per_job_latency_ms_max": round(1000 * max(done_at.values()), 1)}
            wait_all(lambda n: conds(cs.get(n)).get("Succeeded") == "True", names, 300)
            for n in names:
                cs.delete(n)
        finally:
            op.stop()

    lat = []
    with tempfile.TemporaryDirectory() as d:
        op = Operator(ServerOption(fake_gpus=0, leader_elect=False, state_dir=d))
        op.start()
        try:
            cs = op.clientset.kubeflow_v2beta1().mpijobs("default")
            for i in range(a.pi_runs):
                job = yaml_io.load_file(os.path.join(ROOT, "examples/pi/pi.yaml"))[0]
                job.metadata["namespace"] = "default"
                job.metadata["name"] = f"pi-{i}"
                t0 = time.time()
                cs.create(job)
                wait_all(lambda n: conds(cs.get(n)).get("Succeeded") == "True", [job.name], 120)
                lat.append(time.time() - t0)
        finally:
            op.stop()
    out["part2_pi_create_to_succeeded_s"] = {"runs": a.pi_runs, "median": round(statistics.median(lat), 3), "min": round(min(lat), 3),
                                             "max": round(max(lat), 3), "reference_sample_job_s": 75, "reference_e2e_wait_budget_s": 200}
    line = json.dumps(out)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
