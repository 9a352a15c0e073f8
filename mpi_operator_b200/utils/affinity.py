"""CPU / NUMA affinity of a rank = the cores next to its GPU.

Open MPI does this for the reference's jobs through ``-bind-to`` / ``-map-by`` (the example YAMLs pass ``-bind-to none``,
examples/v2beta1/tensorflow-benchmarks/tensorflow-benchmarks.yaml:21-24, and leave placement to the kernel). On an
8-GPU box every rank copies its batch from pinned host memory each step; pinned pages are first-touched on the NUMA node
the process runs on, so a rank that wanders to the far socket pays the inter-socket link on every H2D copy. ``bind_to_gpu``
pins the calling process to the CPU set NVML reports for the GPU (intersected with what the cgroup allows) *before* the
pinned buffers are allocated. Best effort: any failure leaves the affinity untouched. ``B200MPI_NO_AFFINITY=1`` disables it."""
from __future__ import annotations

import os
from typing import Optional, Set


def gpu_cpu_set(device_index: int) -> Optional[Set[int]]:
    """Logical CPUs NVML considers local to the GPU (None when NVML or the query is unavailable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in visible.split(",") if v.strip() != ""]
            phys = int(ids[device_index]) if ids and device_index < len(ids) and ids[device_index].strip().isdigit() else device_index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
            return cpus or None
        finally:
            pynvml.nvmlShutdown()
    except Exception:  # noqa: BLE001 - no NVML, no permission, exotic topology: stay where we are
        return None


def bind_to_gpu(device_index: int) -> Optional[Set[int]]:
    """Restrict this process to the GPU-local CPUs. Returns the new CPU set, or None if nothing was changed."""
    if os.environ.get("B200MPI_NO_AFFINITY") == "1" or not hasattr(os, "sched_setaffinity"):
        return None
    local = gpu_cpu_set(device_index)
    if not local:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        target = allowed & local
        if not target or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return target
    except OSError:
        return None
