#!/usr/bin/env python
"""`make generate` / `make verify-generate`: regenerates every derived artefact from
the typed API model (the reference does this with k8s code-generators + controller-gen +
openapi-generator: hack/update-codegen.sh, hack/python-sdk/gen-sdk.sh, Makefile:87-98)."""
import json
import os
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpi_operator_b200.api import openapi  # noqa: E402


def render():
    sys.path.insert(0, os.path.join(ROOT, "hack"))
    import gen_sdk
    crd = yaml.safe_dump(openapi.crd(), sort_keys=False)
    operator_cfg = open(os.path.join(ROOT, "manifests/base/operator-config.yaml")).read()
    out = {
        "manifests/base/kubeflow.org_mpijobs.yaml": "---\n" + crd,
        "sdk/python/v2beta1/swagger.json": json.dumps(openapi.swagger(), indent=2, sort_keys=True) + "\n",
        "deploy/v2beta1/mpi-operator.yaml": "# all-in-one: CRD schema + daemon config (hack/generate.py; reference: hack/generate-manifest.sh:24-37)\n---\n" + crd + "---\n" + operator_cfg,
    }
    out.update(gen_sdk.render_docs(openapi.swagger()))  # per-model SDK docs (reference: sdk/python/v2beta1/docs/*.md)
    out.update(gen_sdk.render_meta_docs())     # ... and of the generic apimachinery models (sdk/meta_models.py)
    return out


def main():
    verify = "--verify" in sys.argv
    bad = 0
    for rel, content in render().items():
        path = os.path.join(ROOT, rel)
        if verify:
            cur = open(path).read() if os.path.exists(path) else None
            if cur != content:
                print(f"out of date: {rel} (run `make generate`)")
                bad += 1
        else:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as f:
                f.write(content)
            print("generated", rel)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
