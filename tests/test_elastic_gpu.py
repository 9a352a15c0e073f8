"""Elastic rescale on real GPUs through the operator: 1 -> 2 -> 1 workers (2 GPUs) or 2 -> 4 -> 2 (>= 4 GPUs)
mid-run; the job must resume from the committed step on each new world and succeed (BASELINE.json config #4)."""
import os
import sys
import time

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_elastic_scale_up_and_down(tmp_path):
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from helpers import conds, new_mpijob
    from mpi_operator_b200.cmd.options import ServerOption
    from mpi_operator_b200.cmd.server import Operator
    ngpu = torch.cuda.device_count()
    small, big = (2, 4) if ngpu >= 4 else (1, 2)
    ckpt = str(tmp_path / "ckpt.pt")
    op = Operator(ServerOption(leader_elect=False, state_dir=str(tmp_path / "state")))
    op.start()
    try:
        job = new_mpijob("elastic", workers=small, launcher_cmd=("mpirun",), worker_cmd=("/usr/sbin/sshd", "-De"),
                         launcher_args=("python", os.path.join(REPO, "examples/horovod/elastic_mnist.py"), "--total-steps", "1500",
                                        "--commit-every", "10", "--step-sleep", "0.005", "--checkpoint", ckpt))
        job.spec.replica("Worker").template["spec"]["containers"][0]["resources"] = {"limits": {"nvidia.com/gpu": 1}}
        c = op.clientset.kubeflow_v2beta1().mpijobs("default")
        c.create(job)

        def logs():
            out = ""
            for p in op.store.list("pods", "default"):
                if "launcher" in p["metadata"]["name"]:
                    out += op.agent.logs("default", p["metadata"]["name"])
            return out

        def wait(pred, what, timeout=120):
            t0 = time.time()
            while time.time() - t0 < timeout:
                if pred():
                    return
                time.sleep(0.1)
            raise AssertionError(f"timeout: {what}\n{logs()}")

        def scale(n):
            j = c.get("elastic")
            j.spec.replica("Worker").replicas = n
            c.update(j)

        wait(lambda: f"world size {small}" in logs(), "first incarnation")
        scale(big)
        wait(lambda: f"world size {big}" in logs(), "scaled up incarnation")
        assert "(re)started at step 0 " not in logs().split(f"world size {big}")[0].split("\n")[-1]
        scale(small)
        wait(lambda: logs().count(f"world size {small}") >= 2, "scaled down incarnation")
        wait(lambda: conds(c.get("elastic")).get("Succeeded") == "True", "job success", timeout=240)
        text = logs()
        assert f"world sizes seen: [{small}, {big}, {small}]" in text, text
        # each restart resumed from a committed step > 0
        restarts = [l for l in text.splitlines() if "(re)started at step" in l]
        assert len(restarts) == 3 and all(int(l.split("step ")[1].split()[0]) > 0 for l in restarts[1:]), restarts
    finally:
        op.stop()
