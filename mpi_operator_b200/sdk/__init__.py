"""Python SDK: the reference's ``mpijob`` package surface
— model classes with the same constructor kwargs / openapi_types / attribute_map
/ to_dict / to_str / __eq__, ApiClient, Configuration, exceptions — plus
``MPIJobClient`` which submits to the single-box daemon instead of
``kubernetes.client.CustomObjectsApi``."""
__version__ = "0.4.0"

from .api_client import ApiClient  # noqa: F401
from .client import MPIJobClient  # noqa: F401
from .configuration import Configuration  # noqa: F401
from .exceptions import (ApiAttributeError, ApiException, ApiKeyError, ApiTypeError, ApiValueError,  # noqa: F401
                         OpenApiException)
from .models import (V1Container, V1LabelSelector, V1LabelSelectorRequirement, V1ListMeta, V1ObjectMeta,  # noqa: F401
                     V1OwnerReference, V1PodSpec, V1PodTemplateSpec, V2beta1JobCondition, V2beta1JobStatus,
                     V2beta1MPIJob, V2beta1MPIJobList, V2beta1MPIJobSpec, V2beta1ReplicaSpec, V2beta1ReplicaStatus,
                     V2beta1RunPolicy, V2beta1SchedulingPolicy)
from . import meta_models as _meta_models  # noqa: E402  (the generic apimachinery models, built from a schema table)
from .meta_models import *  # noqa: E402,F401,F403
