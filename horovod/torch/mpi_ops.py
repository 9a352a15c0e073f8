"""``horovod.torch.mpi_ops``: the operation-level names some scripts import directly."""
from mpi_operator_b200.hvd import (Adasum, Average, Max, Min, Sum, allgather, allgather_async, allreduce, allreduce_, allreduce_async,  # noqa: F401
                                   allreduce_async_, alltoall, barrier, broadcast, broadcast_, broadcast_async, broadcast_async_,
                                   cross_rank, cross_size, grouped_allreduce, grouped_allreduce_async, init, is_initialized, join,
                                   local_rank, local_size, mpi_built, mpi_enabled, mpi_threads_supported, nccl_built, poll, rank,
                                   reducescatter, shutdown, size, start_timeline, stop_timeline, synchronize)
