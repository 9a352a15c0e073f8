#!/usr/bin/env python
"""Submit the Horovod MNIST MPIJob with the SDK (counterpart of the reference's
sdk/python/v2beta1/tensorflow-mnist.py:91-128, which posts through
kubernetes.client.CustomObjectsApi). Needs a running daemon:
    python -m mpi_operator_b200.cmd.main --listen 127.0.0.1:8087 &
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from mpijob import (MPIJobClient, V1Container, V1ObjectMeta, V1PodSpec, V1PodTemplateSpec, V2beta1MPIJob,
                    V2beta1MPIJobSpec, V2beta1ReplicaSpec, V2beta1RunPolicy)


def main():
    launcher = V1Container(
        name="mpi-launcher", image="docker.io/kubeflow/mpi-horovod-mnist", command=["mpirun"],
        args=["-np", "2", "--allow-run-as-root", "-bind-to", "none", "-map-by", "slot", "-x", "LD_LIBRARY_PATH", "-x", "PATH",
              "-mca", "pml", "ob1", "-mca", "btl", "^openib", "python", "/examples/tensorflow_mnist.py"],
        resources={"limits": {"cpu": 1, "memory": "2Gi"}})
    worker = V1Container(name="mpi-worker", image="docker.io/kubeflow/mpi-horovod-mnist",
                         resources={"limits": {"nvidia.com/gpu": 1, "cpu": 2, "memory": "4Gi"}})
    job = V2beta1MPIJob(
        api_version="kubeflow.org/v2beta1", kind="MPIJob", metadata=V1ObjectMeta(name="tensorflow-mnist", namespace="default"),
        spec=V2beta1MPIJobSpec(
            slots_per_worker=1, run_policy=V2beta1RunPolicy(clean_pod_policy="Running"),
            mpi_replica_specs={
                "Launcher": V2beta1ReplicaSpec(replicas=1, template=V1PodTemplateSpec(spec=V1PodSpec(containers=[launcher]))),
                "Worker": V2beta1ReplicaSpec(replicas=2, template=V1PodTemplateSpec(spec=V1PodSpec(containers=[worker])))}))
    api = MPIJobClient(os.environ.get("MPIJOB_SERVER", "127.0.0.1:8087"))
    api.create_namespaced_custom_object(group="kubeflow.org", version="v2beta1", namespace="default", plural="mpijobs", body=job)
    done = api.wait_for_condition("tensorflow-mnist", "Succeeded", timeout=600)
    print(api.logs("tensorflow-mnist"))
    print("conditions:", [(c["type"], c["status"]) for c in done["status"]["conditions"]])


if __name__ == "__main__":
    main()
