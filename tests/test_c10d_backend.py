"""torch.distributed backend "b200mpi" (mpi_operator_b200/parallel/c10d_backend.py): the c10d front-end next to the LD_PRELOAD
shim. CPU tensors over the libmpi shim here; the same process group drives the NVSwitch runtime
for CUDA tensors (tests/test_multigpu.py::test_c10d_backend_on_gpus)."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MPIRUN = os.path.join(REPO, "mpi_operator_b200/bin/mpirun")
WORKER = os.path.join(REPO, "tests/c10d_worker.py")


@pytest.mark.skipif(not os.path.exists(MPIRUN), reason="native launcher not built (run make)")
@pytest.mark.parametrize("np_", [1, 3])
def test_collectives_and_ddp_under_the_native_mpirun(np_):
    r = subprocess.run([MPIRUN, "-np", str(np_), sys.executable, WORKER], capture_output=True, text=True, timeout=240, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert r.stdout.count("failures=0") == np_


def test_collectives_and_ddp_under_torchrun():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29741", WORKER], capture_output=True, text=True, timeout=240, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert r.stdout.count("failures=0") == 2
    # no launcher of ours was around to clean up: the backend finalises MPI at exit and rank 0 unlinks the rendezvous segment
    assert not [f for f in os.listdir("/dev/shm") if "29741" in f], os.listdir("/dev/shm")
