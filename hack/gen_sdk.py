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


PY_TYPES = {"integer": "int", "string": "str", "boolean": "bool", "object": "object", "number": "float"}


def _py_type(prop):
    """Swagger property -> (python type text, markdown link or None)."""
    if "$ref" in prop:
        ref = prop["$ref"].rsplit("/", 1)[1]
        if ref.startswith("v2beta1."):
            n = "V2beta1" + ref.split(".", 1)[1]
            return n, f"[**{n}**]({n}.md)"
        return "object", None
    t = prop.get("type")
    if t == "array":
        inner, _ = _py_type(prop.get("items", {}))
        return f"list[{inner}]", None
    if t == "object" and "additionalProperties" in prop:
        inner, _ = _py_type(prop["additionalProperties"])
        return f"dict(str, {inner})", None
    if t == "string" and prop.get("format") == "date-time":
        return "datetime", None
    return PY_TYPES.get(t, "object"), None


def render_docs(sw):
    """Per-model markdown reference (the role of sdk/python/v2beta1/docs/V2beta1*.md in the reference tree,
    produced there by openapi-generator; here rendered straight from our swagger.json)."""
    out = {}
    for name, schema in sorted(sw["definitions"].items()):
        cls = "V2beta1" + name.split(".", 1)[1]
        req = set(schema.get("required", []))
        lines = [f"# {cls}", "", schema.get("description", "").strip(), "", "## Properties",
                 "Name | Type | Description | Notes", "------------ | ------------- | ------------- | -------------"]
        for jname in sorted(schema["properties"], key=snake):
            prop = schema["properties"][jname]
            t, link = _py_type(prop)
            desc = (prop.get("description") or "").replace("|", "\\|").replace("\n", " ")
            lines.append(f"**{snake(jname)}** | {link or '**' + t + '**'} | {desc} | {'' if jname in req else '[optional]'}")
        lines += ["", "[[Back to README]](../README.md)", ""]
        out[f"sdk/python/v2beta1/docs/{cls}.md"] = "\n".join(lines)
    return out


def render_meta_docs():
    """The same per-model page for the generic apimachinery models (reference: sdk/python/v2beta1/docs/V1*.md,
    IoK8sApimachineryPkg*.md, K8sIoApimachineryPkg*.md), rendered from the schema table in mpi_operator_b200/sdk/meta_models.py."""
    import re
    from mpi_operator_b200.sdk.meta_models import META_MODELS
    out = {}
    for cls_name, cls in sorted(META_MODELS.items()):
        lines = [f"# {cls_name}", "", (cls.__doc__ or "").strip().splitlines()[0] if cls.__doc__ else "", "", "## Properties",
                 "Name | Type | Description | Notes", "------------ | ------------- | ------------- | -------------"]
        for attr in sorted(cls.openapi_types):
            t = cls.openapi_types[attr]
            inner = re.sub(r"^list\[(.*)\]$|^dict\(str, (.*)\)$", lambda m: m.group(1) or m.group(2), t)
            shown = f"[**{t}**]({inner}.md)" if inner in META_MODELS else f"**{t}**"
            lines.append(f"**{attr}** | {shown} | JSON name `{cls.attribute_map[attr]}` | {'' if attr in cls.required else '[optional]'}")
        lines += ["", "[[Back to README]](../README.md)", ""]
        out[f"sdk/python/v2beta1/docs/{cls_name}.md"] = "\n".join(lines)
    return out


def main():
    if "--docs" in sys.argv:
        sw = json.load(open(os.path.join(ROOT, "sdk/python/v2beta1/swagger.json")))
        for rel, text in render_docs(sw).items():
            os.makedirs(os.path.dirname(os.path.join(ROOT, rel)), exist_ok=True)
            open(os.path.join(ROOT, rel), "w").write(text)
            print("wrote", rel)
        return
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
