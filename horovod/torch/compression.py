"""``horovod.torch.compression.Compression`` (``none`` / ``fp16``)."""
from mpi_operator_b200.hvd import Compression  # noqa: F401
