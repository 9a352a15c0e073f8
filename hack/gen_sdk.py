#!/usr/bin/env python
"""SDK generator check (role of hack/python-sdk/main.go + gen-sdk.sh in the reference): every
definition and property of swagger.json must be present in mpi_operator_b200/sdk/models.py with the
same attribute_map; prints a diff and exits 1 otherwise. The model classes themselves are
declarative (one table, a metaclass) so "regeneration" is keeping that table in sync."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpi_operator_b200.sdk import models  # noqa: E402


def snake(name):
    return re.sub(r"(?<!^)(?=[A-Z])", "_", name).lower()


def main():
    sw = json.load(open(os.path.join(ROOT, "sdk/python/v2beta1/swagger.json")))
    bad = 0
    for name, schema in sw["definitions"].items():
        cls_name = "V2beta1" + name.split(".", 1)[1]
        cls = models.MODEL_CLASSES.get(cls_name)
        if cls is None:
            print("missing model", cls_name)
            bad += 1
            continue
        have = set(cls.attribute_map.values())
        want = set(schema["properties"])
        if have != want:
            print(f"{cls_name}: attribute_map mismatch: missing {sorted(want - have)} extra {sorted(have - want)}")
            bad += 1
    print("sdk models in sync with swagger.json" if not bad else f"{bad} problem(s)")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
