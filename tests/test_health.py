"""node/health.py: the NVML probe against a stand-in `pynvml` (what it flags, what it tolerates) and the allocator's cordon rules."""
import sys
import types

import pytest

from mpi_operator_b200.node.allocator import GangAllocator, SlotRequest
from mpi_operator_b200.node.health import nvml_probe
from mpi_operator_b200.node.topology import GPU, Topology


def fake_nvml(state):
    m = types.ModuleType("pynvml")
    m.NVML_MEMORY_ERROR_TYPE_UNCORRECTED, m.NVML_VOLATILE_ECC, m.NVML_FEATURE_ENABLED = 1, 0, 1

    class NVMLError(Exception):
        pass
    m.NVMLError = NVMLError
    m.nvmlInit = lambda: None
    m.nvmlShutdown = lambda: None
    m.nvmlDeviceGetCount = lambda: len(state)
    m.nvmlDeviceGetHandleByIndex = lambda i: i

    def mem(h):
        if state[h].get("lost"):
            raise NVMLError("GPU is lost")
        return types.SimpleNamespace(total=180 << 30)
    m.nvmlDeviceGetMemoryInfo = mem
    m.nvmlDeviceGetTotalEccErrors = lambda h, et, ct: state[h].get("ecc", 0)

    def retired(h):
        if state[h].get("no_retire_api"):
            raise NVMLError("Not Supported")
        return 1 if state[h].get("retire_pending") else 0
    m.nvmlDeviceGetRetiredPagesPendingStatus = retired
    m.nvmlDeviceGetRemappedRows = lambda h: (0, 0, 1 if state[h].get("remap_pending") else 0, 1 if state[h].get("remap_failed") else 0)
    return m


def test_nvml_probe_flags_lost_ecc_retirement_and_remap(monkeypatch):
    state = [{}, {"ecc": 3}, {"lost": True}, {"retire_pending": True}, {"remap_failed": True}, {"no_retire_api": True}, {"remap_pending": True}]
    monkeypatch.setitem(sys.modules, "pynvml", fake_nvml(state))
    v = nvml_probe()
    # default: only the unambiguous conditions take a GPU out of service
    assert [i for i, why in v.items() if why] == [2, 4] and "not reachable" in v[2] and "remapping failed" in v[4]
    monkeypatch.setenv("B200MPI_GPU_HEALTH_STRICT", "1")       # contained errors count as well
    v = nvml_probe()
    assert v[0] is None and v[5] is None                       # an unsupported query is skipped, not a failure
    assert "ECC" in v[1] and "retirement" in v[3] and "pending" in v[6]
    monkeypatch.setitem(sys.modules, "pynvml", None)           # no NVML at all (CPU box): nothing to say
    assert nvml_probe() == {}


def test_allocator_cordon_rules():
    a = GangAllocator(Topology([GPU(i) for i in range(4)], "fake"))
    assert a.allocate(SlotRequest("ns/p0", gpus=1)) == [0]
    assert a.cordon(0, "bad") and a.cordon(2, "manual") and not a.cordon(2, "manual")      # idempotent
    assert a.free_gpus == 2 and a.cordoned == {0: "bad", 2: "manual"}
    assert a.allocate(SlotRequest("ns/p1", gpus=3)) is None     # only GPUs 1 and 3 are usable
    assert a.allocate_gang([SlotRequest("ns/g0", gpus=1, group="g"), SlotRequest("ns/g1", gpus=1, group="g")], 2) == {"ns/g0": [1], "ns/g1": [3]}
    a.release("ns/p0")
    assert a.free_gpus == 0                                     # GPU 0 was released while cordoned: it stays out
    assert a.uncordon(0) and a.free_gpus == 1 and not a.uncordon(0)
    a.release("ns/g0")
    assert a.uncordon(2) and a.free_gpus == 3
    with pytest.raises(ValueError):
        a.cordon(9)
