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


def _register_model_modules() -> None:
    """The generated package has one module per model (``mpijob.models.v2beta1_mpi_job``, ``mpijob.models.v1_object_meta``,
    ``mpijob.models.io_k8s_apimachinery_pkg_apis_meta_v1_object_meta`` ...: sdk/python/v2beta1/mpijob/models/*.py). The
    classes here come from tables, so those import paths are registered as synthetic modules holding the one class."""
    import sys
    import types
    from mpi_operator_b200.sdk.meta_models import snake
    sys.modules.setdefault(__name__ + ".models", models)
    for mod in ("api_client", "configuration", "exceptions", "rest"):
        sys.modules.setdefault(f"{__name__}.{mod}", globals()[mod])
    for cls_name, cls in models.MODEL_CLASSES.items():
        mod_name = f"{__name__}.models.{snake(cls_name)}"
        if mod_name not in sys.modules:
            m = types.ModuleType(mod_name, f"{cls_name} (see mpi_operator_b200.sdk)")
            setattr(m, cls_name, cls)
            sys.modules[mod_name] = m
            setattr(models, snake(cls_name), m)


_register_model_modules()
