"""Packaging of the SDK (reference: sdk/python/v2beta1/setup.py:31-69, name kubeflow-mpi 0.4.0)."""
import os

from setuptools import setup

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
setup(
    name="kubeflow-mpi", version="0.4.0", description="Python SDK for the single-box MPIJob operator",
    packages=["mpijob", "mpi_operator_b200.sdk"], package_dir={"": ROOT},
    python_requires=">=3.9", install_requires=["pyyaml"],
)
