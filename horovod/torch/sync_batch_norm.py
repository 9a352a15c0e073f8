"""``horovod.torch.sync_batch_norm.SyncBatchNorm``."""
from mpi_operator_b200.hvd.sync_batch_norm import SyncBatchNorm  # noqa: F401
