"""GPU health monitoring: the single-box stand-in for what a cluster gets from the NVIDIA device plugin's health checks plus
`kubectl cordon` (the reference delegates both to Kubernetes; SURVEY.md section 5.3 "failure detection").

A probe reports, per GPU index, ``None`` (healthy) or a reason string. The monitor cordons GPUs whose probe fails in the gang
allocator - no NEW ranks are placed on them; reservations that already hold the GPU keep it, their ranks fail on their own
through the collective watchdog / exit codes -, lifts only the cordons it set itself once the probe passes again, records a
``v1.Event`` on a Node-shaped object and exports ``b200mpi_gpu_healthy{gpu}``. Manual cordons
(``mpijobctl cordon <gpu>`` / ``PATCH /topology``) are never lifted by the monitor.

The default probe asks NVML (nvidia-ml-py): a device that cannot be queried (fell off the bus, ``NVML_ERROR_GPU_IS_LOST``),
a FAILED row remap, and - only with ``B200MPI_GPU_HEALTH_STRICT=1`` - pending remaps / page retirements and uncorrected volatile ECC errors. Every query is optional: what a
driver or GPU generation does not support is skipped.
"""
from __future__ import annotations

import logging
import os
import threading
from typing import Callable, Dict, Optional

from ..controller import metrics

log = logging.getLogger("node-agent")
HEALTH_PREFIX = "health: "

Probe = Callable[[], Dict[int, Optional[str]]]


def nvml_probe() -> Dict[int, Optional[str]]:
    """{gpu index: None | reason}; {} when NVML is not available (CPU-only box, fake GPUs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
    except Exception:  # noqa: BLE001
        return {}
    out: Dict[int, Optional[str]] = {}
    try:
        n = pynvml.nvmlDeviceGetCount()
        for i in range(n):
            reason = None
            try:
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                pynvml.nvmlDeviceGetMemoryInfo(h)                      # a lost GPU fails here
            except Exception as e:  # noqa: BLE001
                out[i] = f"not reachable through NVML ({type(e).__name__}: {e})"
                continue
            # Default reasons are the unambiguous ones: the GPU cannot be reached, or a row remap FAILED (the board needs service).
            # A pending row remap / page retirement or a non-zero volatile uncorrected-ECC counter means the error was contained
            # (the faulting process was killed) and the GPU stays usable until its next reset: these only count with
            # B200MPI_GPU_HEALTH_STRICT=1, so a box does not lose capacity over a condition NVIDIA documents as "reset when convenient".
            strict = os.environ.get("B200MPI_GPU_HEALTH_STRICT", "0") == "1"
            checks = (
                ("row remapping failed", lambda: bool(pynvml.nvmlDeviceGetRemappedRows(h)[3])),
                ("row remapping pending (GPU reset needed)", lambda: strict and bool(pynvml.nvmlDeviceGetRemappedRows(h)[2])),
                ("page retirement pending (reboot / GPU reset needed)",
                 lambda: strict and pynvml.nvmlDeviceGetRetiredPagesPendingStatus(h) == pynvml.NVML_FEATURE_ENABLED),
                ("uncorrected ECC errors since the last reset",
                 lambda: strict and pynvml.nvmlDeviceGetTotalEccErrors(h, pynvml.NVML_MEMORY_ERROR_TYPE_UNCORRECTED, pynvml.NVML_VOLATILE_ECC) > 0),
            )
            for what, bad in checks:
                try:
                    if bad():
                        reason = what
                        break
                except Exception:  # noqa: BLE001  - not supported on this GPU / driver
                    continue
            out[i] = reason
    finally:
        try:
            pynvml.nvmlShutdown()
        except Exception:  # noqa: BLE001
            pass
    return out


class GpuHealthMonitor:
    def __init__(self, agent, probe: Optional[Probe] = None, interval: Optional[float] = None, recorder=None):
        self.agent = agent
        self.probe = probe or nvml_probe
        self.interval = float(os.environ.get("B200MPI_GPU_HEALTH_INTERVAL_S", 30)) if interval is None else float(interval)
        self.recorder = recorder
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.last: Dict[int, Optional[str]] = {}

    def node_object(self) -> dict:
        import socket
        return {"apiVersion": "v1", "kind": "Node", "metadata": {"name": socket.gethostname(), "namespace": "default", "uid": "node"}}

    def check_once(self) -> Dict[int, Optional[str]]:
        """One probe round; returns the probe's verdicts. Cordons / uncordons and events follow the CHANGES."""
        try:
            verdicts = self.probe() or {}
        except Exception as e:  # noqa: BLE001  - a broken probe must not take GPUs away
            log.warning("GPU health probe failed: %s", e)
            return self.last
        alloc = self.agent.alloc
        known = {g.index for g in self.agent.topology.gpus}
        cordoned = alloc.cordoned
        for gpu, reason in verdicts.items():
            if gpu not in known:
                continue
            metrics.gpu_healthy.labels(gpu=str(gpu)).set(0 if reason else 1)
            if reason:
                if gpu not in cordoned:       # never replace a manual cordon's reason
                    alloc.cordon(gpu, HEALTH_PREFIX + reason)
                    log.warning("GPU %d cordoned: %s", gpu, reason)
                    if self.recorder is not None:
                        self.recorder.event(self.node_object(), "Warning", "GPUUnhealthy", f"GPU {gpu}: {reason}; no new ranks will be placed on it")
            elif cordoned.get(gpu, "").startswith(HEALTH_PREFIX):
                alloc.uncordon(gpu)
                log.info("GPU %d passes its health probe again: uncordoned", gpu)
                if self.recorder is not None:
                    self.recorder.event(self.node_object(), "Normal", "GPUHealthy", f"GPU {gpu} passes its health probe again")
        self.last = dict(verdicts)
        metrics.gpu_slots_free.set(alloc.free_gpus)
        self.agent.wake()
        return verdicts

    def start(self) -> None:
        if self.interval <= 0 or self._thread is not None:
            return

        def loop():
            while not self._stop.wait(self.interval):
                self.check_once()
        self.check_once()
        self._thread = threading.Thread(target=loop, name="gpu-health", daemon=True)
        self._thread.start()

    def stop(self) -> None:
        self._stop.set()
