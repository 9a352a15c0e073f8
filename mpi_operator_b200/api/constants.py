"""Names, labels and defaults of the MPIJob API.

Reference: pkg/apis/kubeflow/v2beta1/constants.go:17-46 and
pkg/apis/kubeflow/v2beta1/types.go:96-102.  Values are part of the user-visible
surface and kept byte-identical.
"""
# Environment variable for the namespace when the operator runs inside a cluster.
ENV_KUBEFLOW_NAMESPACE = "KUBEFLOW_NAMESPACE"
DEFAULT_RESTART_POLICY = "Never"
DEFAULT_LAUNCHER_RESTART_POLICY = "OnFailure"
OPERATOR_NAME = "mpi-operator"

# labels
REPLICA_INDEX_LABEL = "training.kubeflow.org/replica-index"
REPLICA_TYPE_LABEL = "training.kubeflow.org/replica-type"
OPERATOR_NAME_LABEL = "training.kubeflow.org/operator-name"
JOB_NAME_LABEL = "training.kubeflow.org/job-name"
JOB_ROLE_LABEL = "training.kubeflow.org/job-role"

# spec.runPolicy.managedBy values (types.go:96-102)
KUBEFLOW_JOB_CONTROLLER = "kubeflow.org/mpi-operator"
MULTIKUEUE_CONTROLLER = "kueue.x-k8s.io/multikueue"

# group / version / kind (register.go:23-36)
GROUP_NAME = "kubeflow.org"
GROUP_VERSION = "v2beta1"
KIND = "MPIJob"
API_VERSION = GROUP_NAME + "/" + GROUP_VERSION
PLURAL = "mpijobs"
SINGULAR = "mpijob"

# enums (types.go)
REPLICA_TYPE_LAUNCHER = "Launcher"
REPLICA_TYPE_WORKER = "Worker"
CLEAN_POD_POLICY_UNDEFINED = ""
CLEAN_POD_POLICY_ALL = "All"
CLEAN_POD_POLICY_RUNNING = "Running"
CLEAN_POD_POLICY_NONE = "None"
RESTART_POLICY_ALWAYS = "Always"
RESTART_POLICY_ON_FAILURE = "OnFailure"
RESTART_POLICY_NEVER = "Never"
RESTART_POLICY_EXIT_CODE = "ExitCode"
MPI_IMPLEMENTATION_OPENMPI = "OpenMPI"
MPI_IMPLEMENTATION_INTEL = "Intel"
MPI_IMPLEMENTATION_MPICH = "MPICH"
LAUNCHER_CREATION_POLICY_AT_STARTUP = "AtStartup"
LAUNCHER_CREATION_POLICY_WAIT_FOR_WORKERS_READY = "WaitForWorkersReady"

# JobConditionType (types.go:309-340)
JOB_CREATED = "Created"
JOB_RUNNING = "Running"
JOB_RESTARTING = "Restarting"
JOB_SUCCEEDED = "Succeeded"
JOB_SUSPENDED = "Suspended"
JOB_FAILED = "Failed"

CONDITION_TRUE = "True"
CONDITION_FALSE = "False"
CONDITION_UNKNOWN = "Unknown"

# gang scheduling (podgroup.go; [EXT] volcano / scheduler-plugins constants)
VOLCANO_QUEUE_NAME_ANNOTATION = "scheduling.volcano.sh/queue-name"
VOLCANO_GROUP_NAME_ANNOTATION = "scheduling.k8s.io/group-name"
SCHED_PLUGINS_POD_GROUP_LABEL = "scheduling.x-k8s.io/pod-group"
GANG_SCHEDULER_VOLCANO = "volcano"

# single-box extension: resource name that maps onto GPU slots
GPU_RESOURCE = "nvidia.com/gpu"
