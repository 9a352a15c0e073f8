"""``import horovod.torch as hvd`` compatibility: resolves to mpi_operator_b200.hvd
(the LD/PYTHONPATH-injected replacement for the Horovod the reference's images ship)."""
from .runner import run  # noqa: F401  (horovod.run)

__version__ = "0.20.0+b200mpi"
