"""Rank script for the fault-injection tests: a Horovod-shaped loop that calls the injector once per step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpi_operator_b200.launch.env import rank_info_from_env  # noqa: E402
from mpi_operator_b200.utils import fault  # noqa: E402

info = rank_info_from_env()
inj = fault.injector(info.rank)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for step in range(1, steps + 1):
    inj.on_step()
    print(f"rank {info.rank}/{info.world_size} step {step}", flush=True)
    time.sleep(float(os.environ.get("STEP_SLEEP", "0.01")))
print(f"rank {info.rank} done", flush=True)
