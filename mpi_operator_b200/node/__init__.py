"""Node agent: the single-box replacement for kube-scheduler, kubelet and the
batch/v1 Job controller.  Pods become process
groups pinned to GPU slots discovered from the NVLink/NVSwitch topology."""
from .agent import NodeAgent  # noqa: F401
from .allocator import GangAllocator, SlotRequest  # noqa: F401
from .topology import Topology, discover_topology  # noqa: F401
