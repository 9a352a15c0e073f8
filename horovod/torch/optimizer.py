"""``horovod.torch.optimizer.DistributedOptimizer``."""
from mpi_operator_b200.hvd.optimizer import DistributedOptimizer  # noqa: F401
