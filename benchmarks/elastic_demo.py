#!/usr/bin/env python
"""Elastic rescale demo + timing (BASELINE.json config #4: 4 -> 8 -> 4 workers mid-run).

Runs examples/horovod/elastic_mnist.py as an MPIJob through the in-process operator, scales Worker.replicas while it runs
and reports, per transition, how long it took from the `scale` request until the new world size reported its first step
(controller: new worker pods + regenerated discover_hosts.sh; ranks: notice at the next commit, leave with the rescale
code; launcher: restart mpirun on the new world; state: restored from rank 0's checkpoint and broadcast).
On a host without CUDA (or with --cpu) the ranks use the libmpi-shim backend; on a GPU box each worker takes one GPU.
The reference describes the mechanism (proposals/elastic-horovod.md) but publishes no timing."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import conds, new_mpijob  # noqa: E402
from mpi_operator_b200.cmd.options import ServerOption  # noqa: E402
from mpi_operator_b200.cmd.server import Operator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4,8,4", help="world sizes to step through")
    ap.add_argument("--cpu", action="store_true", help="force the CPU backend (B200MPI_HVD_DEVICE=cpu)")
    ap.add_argument("--total-steps", type=int, default=400)
    ap.add_argument("--step-sleep", type=float, default=0.02)
    ap.add_argument("--out", default="")
    ap.add_argument("--inplace", action="store_true",
                    help="B200MPI_ELASTIC_INPLACE=1: mpirun spawns only the additional ranks and survivors re-form the communicator in place")
    a = ap.parse_args()
    sizes = [int(v) for v in a.sizes.split(",")]
    import torch
    cpu = a.cpu or not torch.cuda.is_available()
    with tempfile.TemporaryDirectory() as d:
        op = Operator(ServerOption(leader_elect=False, state_dir=os.path.join(d, "state"), fake_gpus=max(sizes) if cpu else None))
        op.start()
        try:
            job = new_mpijob("elastic", workers=sizes[0], launcher_cmd=("mpirun",), worker_cmd=("/usr/sbin/sshd", "-De"),
                             launcher_args=("python", os.path.join(ROOT, "examples/horovod/elastic_mnist.py"), "--total-steps", str(a.total_steps),
                                            "--commit-every", "5", "--step-sleep", str(a.step_sleep), "--checkpoint", os.path.join(d, "ckpt.pt")))
            c0 = job.spec.replica("Launcher").template["spec"]["containers"][0]
            c0["env"] = [{"name": "B200MPI_ELASTIC_INPLACE", "value": "1"}] if a.inplace else []
            if cpu:
                c0["env"].append({"name": "B200MPI_HVD_DEVICE", "value": "cpu"})
            else:
                job.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = {"limits": {"nvidia.com/gpu": 1}}
            c = op.clientset.kubeflow_v2beta1().mpijobs("default")
            t_create = time.time()
            c.create(job)

            def logs():
                return "".join(op.agent.logs("default", p["metadata"]["name"]) for p in op.store.list("pods", "default")
                               if "launcher" in p["metadata"]["name"])

            def wait(pred, what, timeout=300):
                t0 = time.time()
                while time.time() - t0 < timeout:
                    if pred():
                        return time.time()
                    time.sleep(0.02)
                raise SystemExit(f"timeout: {what}\n{logs()[-3000:]}")

            seen = {}
            t_first = wait(lambda: f"with world size {sizes[0]}" in logs(), "first incarnation")
            seen[sizes[0]] = 1
            phases = [{"phase": f"start at {sizes[0]}", "seconds": round(t_first - t_create, 2)}]
            for prev, nxt in zip(sizes, sizes[1:]):
                time.sleep(1.0)  # let the current world make some progress first
                want = seen.get(nxt, 0) + 1
                j = c.get("elastic")
                j.spec.replica("Worker").replicas = nxt
                t0 = time.time()
                c.update(j)
                t1 = wait(lambda: logs().count(f"with world size {nxt};") >= want, f"world size {nxt}")
                seen[nxt] = want
                phases.append({"phase": f"{prev} -> {nxt}", "seconds": round(t1 - t0, 2)})
            t_done = wait(lambda: conds(c.get("elastic")).get("Succeeded") == "True", "job success", timeout=600)
            text = logs()
            restarts = [ln for ln in text.splitlines() if "(re)started at step" in ln]
            reformed = [ln.split("re-formed in place: ")[1] for ln in text.splitlines() if "re-formed in place" in ln]
            out = {"backend": "cpu (libmpi shim)" if cpu else "gpu (b200mpi)", "mode": "in place" if a.inplace else "restart", "sizes": sizes,
                   "phases": phases, "survivor_reform": reformed,
                   "total_seconds": round(t_done - t_create, 2), "resumed_from_steps": [int(ln.split("step ")[1].split()[0]) for ln in restarts],
                   "worlds_seen": text.split("world sizes seen: ")[-1].strip().splitlines()[0] if "world sizes seen" in text else None}
            print(json.dumps(out))
            if a.out:
                with open(a.out, "w") as f:
                    f.write(json.dumps(out, indent=1) + "\n")
        finally:
            op.stop()


if __name__ == "__main__":
    main()
