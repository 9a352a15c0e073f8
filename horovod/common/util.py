"""``horovod.common.util`` build probes (``nccl_built()``, ``mpi_built()``, ...)."""
from mpi_operator_b200.hvd import (ccl_built, cuda_built, ddl_built, gloo_built, gloo_enabled, mpi_built, mpi_enabled, nccl_built,  # noqa: F401
                                   rocm_built)
