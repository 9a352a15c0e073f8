"""``horovod.common.exceptions`` — what elastic training loops catch."""
from mpi_operator_b200.hvd.exceptions import HorovodInternalError, HostsUpdatedInterrupt  # noqa: F401


class HorovodVersionMismatchError(ImportError):
    pass
