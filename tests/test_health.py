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


def test_allocator_places_reservations_by_socket(monkeypatch):
    """Placement is NUMA-aware (node/allocator.py::_select): best fit on one socket, whole sockets first when a gang spans both,
    plain first-n without NUMA information. The reference leaves placement to kube-scheduler / Volcano."""
    from mpi_operator_b200.node.allocator import GangAllocator, SlotRequest
    from mpi_operator_b200.node.topology import discover_topology
    monkeypatch.setenv("B200MPI_FAKE_GPUS", "8")
    monkeypatch.setenv("B200MPI_FAKE_NUMA_NODES", "2")
    topo = discover_topology()
    assert [g.numa_node for g in topo.gpus] == [0, 0, 0, 0, 1, 1, 1, 1] and topo.to_dict()["gpus"][5]["numa_node"] == 1
    a = GangAllocator(topo)
    assert a.allocate(SlotRequest("d/a", 2)) == [0, 1]                    # both sockets fit: the lower one
    assert a.allocate(SlotRequest("d/b", 2)) == [2, 3]                    # best fit: fills socket 0 instead of opening socket 1
    gang = [SlotRequest(f"d/g-{k}", 1, group="g") for k in range(4)]
    assert a.allocate_gang(gang, 4) == {f"d/g-{k}": [4 + k] for k in range(4)}   # a 4-GPU gang still finds a whole socket
    a.release("d/a")
    a.release("d/g-0")
    a.release("d/g-1")                                                     # free: 0,1 (socket 0) and 4,5 (socket 1)
    assert a.allocate(SlotRequest("d/c", 1)) == [0]
    assert a.allocate(SlotRequest("d/d", 2)) == [4, 5]                    # does not split across sockets while one socket holds it
    assert a.allocate(SlotRequest("d/e", 2)) is None and a.free_gpus == 1
    for k in ("d/b", "d/c", "d/d", "d/g-2", "d/g-3"):
        a.release(k)
    assert a.free_gpus == 8
    a.cordon(1, "test")
    six = [SlotRequest(f"d/s-{k}", 1, group="s") for k in range(6)]
    got = a.allocate_gang(six, 6)                                          # spans both: the fuller socket whole, the rest from the other
    assert sorted(g for v in got.values() for g in v) == [0, 2, 4, 5, 6, 7]
    # no NUMA information: the first n free GPUs, as before
    monkeypatch.delenv("B200MPI_FAKE_NUMA_NODES")
    b = GangAllocator(discover_topology())
    assert b.allocate(SlotRequest("d/x", 3)) == [0, 1, 2] and b.allocate(SlotRequest("d/y", 3)) == [3, 4, 5]
