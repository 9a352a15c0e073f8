"""MPIJob API: types, constants, defaults, validation, scheme."""
from . import constants  # noqa: F401
from .defaults import set_defaults_mpijob  # noqa: F401
from .register import scheme  # noqa: F401
from .types import (JobCondition, JobStatus, MPIJob, MPIJobList, MPIJobSpec, ReplicaSpec, ReplicaStatus,  # noqa: F401
                    RunPolicy, SchedulingPolicy)
from .validation import validate_mpijob  # noqa: F401
