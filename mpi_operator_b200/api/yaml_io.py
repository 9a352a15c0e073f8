"""YAML/JSON loading of MPIJob documents (the same files users feed kubectl;
reference examples: examples/v2beta1/pi/pi.yaml, .../tensorflow-benchmarks.yaml)."""
from __future__ import annotations

import json
from typing import Iterable, List

import yaml

from .register import scheme
from .types import MPIJob


def load_all(text: str) -> List[MPIJob]:
    jobs = []
    for doc in yaml.safe_load_all(text):
        if not doc:
            continue
        if doc.get("kind") == "List":
            for item in doc.get("items", []):
                jobs.append(scheme.decode(item))
            continue
        obj = scheme.decode(doc)
        if isinstance(obj, MPIJob):
            jobs.append(obj)
        else:
            jobs.extend(obj.items)
    return jobs


def load_file(path: str) -> List[MPIJob]:
    with open(path) as f:
        return load_all(f.read())


def dump(job: MPIJob) -> str:
    return yaml.safe_dump(job.to_dict(), sort_keys=False)


def dump_json(job: MPIJob) -> str:
    return json.dumps(job.to_dict(), indent=2)
