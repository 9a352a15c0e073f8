"""``import horovod.torch as hvd`` compatibility: resolves to mpi_operator_b200.hvd
(the LD/PYTHONPATH-injected replacement for the Horovod the reference's images ship)."""
