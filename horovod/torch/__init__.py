from mpi_operator_b200.hvd import *  # noqa: F401,F403
from mpi_operator_b200.hvd import (Adasum, Average, Compression, DistributedOptimizer, Max, Min, Sum, elastic,  # noqa: F401
                                   allgather, allreduce, allreduce_, barrier, broadcast, broadcast_, broadcast_object,
                                   broadcast_optimizer_state, broadcast_parameters, cross_rank, cross_size, init,
                                   is_initialized, join, local_rank, local_size, nccl_built, rank, shutdown, size)
