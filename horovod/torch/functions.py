"""``horovod.torch.functions``: parameter / optimizer-state / object broadcasts."""
from mpi_operator_b200.hvd import allgather_object, broadcast_object, broadcast_optimizer_state, broadcast_parameters  # noqa: F401
