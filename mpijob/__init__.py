"""``import mpijob`` — drop-in name of the reference's Python SDK package
(sdk/python/v2beta1/mpijob/__init__.py:17-86), re-exported from
``mpi_operator_b200.sdk``."""
from mpi_operator_b200.sdk import *  # noqa: F401,F403
from mpi_operator_b200.sdk import __version__  # noqa: F401
from mpi_operator_b200.sdk import models  # noqa: F401
from mpi_operator_b200.sdk import api_client, configuration, exceptions, rest  # noqa: F401

# The reference's SDK example still imports the pre-v2beta1 name (sdk/python/v2beta1/tensorflow-mnist.py:17), which its own
# package no longer defines; keep that script runnable.
from mpi_operator_b200.sdk.models import V2beta1ReplicaSpec as V1ReplicaSpec  # noqa: E402,F401
