"""``kubernetes``-client compatibility for MPIJob scripts on a single box.

The reference's SDK example and most user scripts submit MPIJobs through the official Kubernetes Python client
(sdk/python/v2beta1/tensorflow-mnist.py:7-25,110-128: ``config.load_kube_config()``, ``client.CustomObjectsApi()
.create_namespaced_custom_object(group="kubeflow.org", version="v2beta1", plural="mpijobs", body=job)``). There is no API
server here; this package offers the slice of that client such scripts use — the pod-template models, ``CustomObjectsApi``
for ``kubeflow.org/v2beta1 mpijobs``, a read-only ``CoreV1Api`` (pods, pod logs, events), ``watch.Watch`` — on top of
``mpi_operator_b200.sdk`` and the daemon's REST API, the way ``horovod/`` stands in for Horovod.  It is NOT the Kubernetes
client: with the real package installed, put it first on ``sys.path`` and point it at a real cluster instead."""
from . import client, config, watch  # noqa: F401

__version__ = "0.0.0+mpi-operator-b200"
