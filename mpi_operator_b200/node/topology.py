"""GPU / NVLink topology discovery.

The reference asks the Kubernetes API server where capacity is; a single-box
daemon asks the hardware (BASELINE.json north-star).  Discovery order: NVML
(``pynvml`` from nvidia-ml-py: device list, UUIDs, NVLink P2P status), then
``torch.cuda``, then ``B200MPI_FAKE_GPUS=N`` for tests.  On HGX B200 every
pair is an NVSwitch peer, so any k-of-8 placement is bandwidth-equivalent; we
still record per-pair P2P capability and refuse non-peer placements.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional


@dataclass
class GPU:
    index: int
    uuid: str = ""
    name: str = ""
    memory_bytes: int = 0
    nvlink_peers: List[int] = field(default_factory=list)


@dataclass
class Topology:
    gpus: List[GPU] = field(default_factory=list)
    source: str = "none"
    multicast_supported: Optional[bool] = None

    @property
    def gpu_count(self) -> int:
        return len(self.gpus)

    def all_peers(self, indices: List[int]) -> bool:
        """True when every pair in ``indices`` can do P2P (NVSwitch: always)."""
        by = {g.index: g for g in self.gpus}
        for i in indices:
            for j in indices:
                if i != j and by[i].nvlink_peers and j not in by[i].nvlink_peers:
                    return False
        return True

    def to_dict(self) -> dict:
        return {"source": self.source, "multicast_supported": self.multicast_supported,
                "gpus": [g.__dict__ for g in self.gpus]}


def _from_nvml() -> Optional[Topology]:
    try:
        import pynvml  # nvidia-ml-py
        pynvml.nvmlInit()
    except Exception:
        return None
    try:
        n = pynvml.nvmlDeviceGetCount()
        handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in range(n)]
        gpus = []
        for i, h in enumerate(handles):
            name = pynvml.nvmlDeviceGetName(h)
            uuid = pynvml.nvmlDeviceGetUUID(h)
            mem = pynvml.nvmlDeviceGetMemoryInfo(h).total
            peers = []
            for j, hj in enumerate(handles):
                if i == j:
                    continue
                try:
                    st = pynvml.nvmlDeviceGetP2PStatus(h, hj, pynvml.NVML_P2P_CAPS_INDEX_NVLINK)
                    if st == pynvml.NVML_P2P_STATUS_OK:
                        peers.append(j)
                except Exception:
                    pass
            gpus.append(GPU(i, uuid if isinstance(uuid, str) else uuid.decode(), name if isinstance(name, str) else name.decode(), int(mem), peers))
        return Topology(gpus, "nvml")
    except Exception:
        return None
    finally:
        try:
            pynvml.nvmlShutdown()
        except Exception:
            pass


def _from_torch() -> Optional[Topology]:
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        n = torch.cuda.device_count()
        gpus = []
        for i in range(n):
            p = torch.cuda.get_device_properties(i)
            peers = [j for j in range(n) if j != i and torch.cuda.can_device_access_peer(i, j)]
            gpus.append(GPU(i, str(getattr(p, "uuid", "")), p.name, int(p.total_memory), peers))
        return Topology(gpus, "torch")
    except Exception:
        return None


def discover_topology() -> Topology:
    fake = os.environ.get("B200MPI_FAKE_GPUS")
    if fake is not None:
        n = int(fake)
        return Topology([GPU(i, f"GPU-fake-{i}", "FakeB200", 180 << 30, [j for j in range(n) if j != i]) for i in range(n)], "fake")
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    topo = _from_nvml() or _from_torch() or Topology([], "none")
    if vis not in (None, "") and topo.source == "nvml":
        try:
            keep = [int(x) for x in vis.split(",") if x.strip() != ""]
            topo.gpus = [g for g in topo.gpus if g.index in keep]
        except ValueError:
            pass
    return topo
