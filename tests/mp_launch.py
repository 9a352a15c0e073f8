"""Tiny local gang launcher for tests: python mp_launch.py -n N script.py [args]."""
import argparse
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpi_operator_b200.launch.env import build_rank_env  # noqa: E402


def launch(n, argv, timeout=300, extra_env=None, log_dir=None):
    """Gang-launch `argv` as n ranks. With log_dir (or $MP_LAUNCH_LOG_DIR) every rank's stdout+stderr goes to
    <log_dir>/<script>.rank<r>.log so a failing GPU run can be diagnosed from one gpurun call."""
    job = f"test-{os.getpid()}-{int(time.time() * 1000) % 100000}"
    log_dir = log_dir or os.environ.get("MP_LAUNCH_LOG_DIR")
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
    procs = []
    for r in range(n):
        env = dict(os.environ)
        env.update(build_rank_env(rank=r, world_size=n, local_rank=r, local_size=n, job_id=job,
                                  master_port=29500 + os.getpid() % 2000))
        env.update(extra_env or {})
        out = None
        if log_dir:
            out = open(os.path.join(log_dir, f"{os.path.basename(argv[0])}.rank{r}.log"), "w")
        procs.append(subprocess.Popen([sys.executable] + argv, env=env, stdout=out, stderr=subprocess.STDOUT if out else None))
    deadline = time.time() + timeout
    rcs = []
    for p in procs:
        try:
            rcs.append(p.wait(timeout=max(1, deadline - time.time())))
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            rcs.append(-9)
    return rcs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", type=int, default=2)
    ap.add_argument("--timeout", type=int, default=300)
    ap.add_argument("--log-dir", default=None)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rcs = launch(a.n, a.rest, a.timeout, log_dir=a.log_dir)
    print("exit codes:", rcs)
    sys.exit(max(abs(r) for r in rcs))
