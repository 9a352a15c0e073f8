"""``horovodrun``: Horovod's launcher command line on top of the native ``mpirun`` (images of the reference's Horovod
examples ship both; examples/v2beta1/horovod/tensorflow-mnist.yaml uses ``mpirun``, Horovod's own docs ``horovodrun``).

    horovodrun -np 4 python train.py
    horovodrun -np 8 --timeline-filename /tmp/tl.json --fusion-threshold-mb 32 --cycle-time-ms 2 python train.py

The tuning flags become the ``HOROVOD_*`` variables the background engine reads (``csrc/hvd_core``); ``-H`` / ``--hostfile``
are passed to mpirun, which maps every host onto this box's slots. ``--check-build`` prints what this build provides."""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path
from typing import List, Optional

MPIRUN = Path(__file__).resolve().parent.parent / "bin" / "mpirun"


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="horovodrun", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("-np", "--num-proc", type=int, dest="np", help="total number of training processes")
    p.add_argument("-H", "--hosts", dest="hosts", help="host1:slots,host2:slots (all hosts are this box)")
    p.add_argument("-hostfile", "--hostfile", dest="hostfile")
    p.add_argument("-cb", "--check-build", action="store_true")
    p.add_argument("--fusion-threshold-mb", type=float)
    p.add_argument("--cycle-time-ms", type=float)
    p.add_argument("--cache-capacity", type=int)
    p.add_argument("--timeline-filename")
    p.add_argument("--timeline-mark-cycles", action="store_true")
    p.add_argument("--no-stall-check", action="store_true")
    p.add_argument("--stall-check-warning-time-seconds", type=float)
    p.add_argument("--stall-check-shutdown-time-seconds", type=float)
    p.add_argument("--autotune", action="store_true", help="search cycle time x fusion threshold for the best allreduce throughput")
    p.add_argument("--autotune-log-file")
    p.add_argument("--autotune-warmup-samples", type=int)
    p.add_argument("--autotune-steps-per-sample", type=int)
    p.add_argument("--start-timeout", type=int, help="seconds to wait for all ranks to join the rendezvous")
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--mpi", action="store_true", help="accepted: the MPI-style launcher is the only one")
    p.add_argument("--gloo", action="store_true", help="accepted and ignored")
    p.add_argument("--mpi-args", default="", help="extra arguments handed to mpirun verbatim")
    p.add_argument("-x", dest="exports", action="append", default=[], metavar="VAR[=value]")
    p.add_argument("command", nargs=argparse.REMAINDER)
    return p


def engine_env(a) -> dict:
    """Horovod's tuning flags -> the variables the engine reads (same names as Horovod's)."""
    e = {}
    if a.fusion_threshold_mb is not None:
        e["HOROVOD_FUSION_THRESHOLD"] = str(int(a.fusion_threshold_mb * 1024 * 1024))
    if a.cycle_time_ms is not None:
        e["HOROVOD_CYCLE_TIME"] = repr(a.cycle_time_ms)
    if a.cache_capacity is not None:
        e["HOROVOD_CACHE_CAPACITY"] = str(a.cache_capacity)
    if a.timeline_filename:
        e["HOROVOD_TIMELINE"] = a.timeline_filename
    if a.timeline_mark_cycles:
        e["HOROVOD_TIMELINE_MARK_CYCLES"] = "1"
    if a.no_stall_check:
        e["HOROVOD_STALL_CHECK_DISABLE"] = "1"
    if a.stall_check_warning_time_seconds is not None:
        e["HOROVOD_STALL_CHECK_TIME_SECONDS"] = repr(a.stall_check_warning_time_seconds)
    if a.stall_check_shutdown_time_seconds is not None:
        e["HOROVOD_STALL_SHUTDOWN_TIME_SECONDS"] = repr(a.stall_check_shutdown_time_seconds)
    if a.autotune:
        e["HOROVOD_AUTOTUNE"] = "1"
    if a.autotune_log_file:
        e["HOROVOD_AUTOTUNE_LOG"] = a.autotune_log_file
    if a.autotune_warmup_samples is not None:
        e["HOROVOD_AUTOTUNE_WARMUP_SAMPLES"] = str(a.autotune_warmup_samples)
    if a.autotune_steps_per_sample is not None:
        e["HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE"] = str(a.autotune_steps_per_sample)
    if a.start_timeout is not None:
        e["B200MPI_INIT_TIMEOUT_MS"] = str(a.start_timeout * 1000)
    if a.verbose:
        e["B200MPI_DEBUG"] = "1"
    return e


def mpirun_argv(a) -> List[str]:
    argv = [str(MPIRUN)]
    if a.np:
        argv += ["-np", str(a.np)]
    if a.hosts:
        argv += ["-H", a.hosts]
    if a.hostfile:
        argv += ["-hostfile", a.hostfile]
    for k, v in engine_env(a).items():
        argv += ["-x", f"{k}={v}"]
    for x in a.exports:
        argv += ["-x", x]
    argv += a.mpi_args.split()
    cmd = a.command[1:] if a.command[:1] == ["--"] else a.command
    return argv + cmd


def check_build() -> str:
    from .. import hvd
    from ..version import __version__ as VERSION
    rows = [("PyTorch", True), ("TensorFlow", False), ("MXNet", False)]
    ctl = [("MPI (native mpirun + libmpi shim)", True), ("Gloo", False)]
    ops = [("b200mpi NVLink / NVLS kernels (in place of NCCL)", bool(hvd.nccl_built())), ("MPI (shared-memory host path)", True),
           ("DDL", False), ("CCL", False), ("Gloo", False)]
    def fmt(items):
        return "\n".join(f"    [{'X' if ok else ' '}] {name}" for name, ok in items)
    return (f"mpi-operator-b200 v{VERSION} (horovod.torch-compatible front-end):\n\nAvailable Frameworks:\n{fmt(rows)}\n\n"
            f"Available Controllers:\n{fmt(ctl)}\n\nAvailable Tensor Operations:\n{fmt(ops)}\n")


def main(argv: Optional[List[str]] = None) -> int:
    a = build_parser().parse_args(argv)
    if a.check_build:
        print(check_build())
        return 0
    if not a.command or a.command == ["--"]:
        print("horovodrun: no command given", file=sys.stderr)
        return 2
    if not MPIRUN.exists():
        print(f"horovodrun: {MPIRUN} is missing; run `make`", file=sys.stderr)
        return 2
    args = mpirun_argv(a)
    os.execv(args[0], args)
    return 127


if __name__ == "__main__":
    sys.exit(main())
