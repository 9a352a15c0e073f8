"""``horovod.torch.elastic`` (``from horovod.torch.elastic import run, TorchState, ElasticSampler``)."""
from mpi_operator_b200.hvd.elastic import *  # noqa: F401,F403
from mpi_operator_b200.hvd.elastic import (ElasticSampler, HorovodInternalError, HostsUpdatedInterrupt, ObjectState, State,  # noqa: F401
                                           TorchState, run)
