"""The generic apimachinery models of the reference SDK, built from one schema table.

Reference: sdk/python/v2beta1/mpijob/models/ ships, next to the 9 ``V2beta1*`` MPIJob models, one generated file per
``meta/v1`` API type - 44 types, each under two names (``v1_*.py`` / ``io_k8s_apimachinery_pkg_apis_meta_v1_*.py``), the
three ``runtime`` / ``version`` types (``io_k8s_apimachinery_pkg_*`` and ``k8s_io_apimachinery_pkg_*``), and six
pre-v2beta1 leftovers (``v1_job_condition.py``, ``v1_job_status.py``, ``v1_replica_spec.py``, ``v1_replica_status.py``,
``v1_run_policy.py``, ``v1_scheduling_policy.py``) - 100 files, all with the same generated body
(``openapi_types`` / ``attribute_map`` / properties with required-field checks / ``to_dict`` / ``to_str`` / ``__eq__``,
e.g. v1_owner_reference.py:35-62,99-112,229-275).

Here the API shapes are DATA: the tables below hold one line per type (``jsonName:type``, ``!`` = required) and the classes
are created from it on import with ``sdk.models.OpenApiModel`` as the base, so every model gets the same constructor
kwargs, properties, validation and (de)serialisation as the MPIJob models; ``ApiClient.deserialize`` finds them through
``MODEL_CLASSES``. ``mpijob/__init__.py`` also registers one module per class (``mpijob.models.v1_object_meta`` ...), the
import paths of the generated package.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

from . import models as _m

# jsonName:type  (! = required). `T` in a type stands for the class prefix the table is instantiated with.
_META_V1 = """
APIGroup                  apiVersion:str kind:str name:str! preferredVersion:TGroupVersionForDiscovery serverAddressByClientCIDRs:list[TServerAddressByClientCIDR] versions:list[TGroupVersionForDiscovery]!
APIGroupList              apiVersion:str groups:list[TAPIGroup]! kind:str
APIResource               categories:list[str] group:str kind:str! name:str! namespaced:bool! shortNames:list[str] singularName:str! storageVersionHash:str verbs:list[str]! version:str
APIResourceList           apiVersion:str groupVersion:str! kind:str resources:list[TAPIResource]!
APIVersions               apiVersion:str kind:str serverAddressByClientCIDRs:list[TServerAddressByClientCIDR]! versions:list[str]!
ApplyOptions              apiVersion:str dryRun:list[str] fieldManager:str! force:bool! kind:str
Condition                 lastTransitionTime:datetime! message:str! observedGeneration:int reason:str! status:str! type:str!
CreateOptions             apiVersion:str dryRun:list[str] fieldManager:str fieldValidation:str kind:str
DeleteOptions             apiVersion:str dryRun:list[str] gracePeriodSeconds:int ignoreStoreReadErrorWithClusterBreakingPotential:bool kind:str orphanDependents:bool preconditions:TPreconditions propagationPolicy:str
FieldSelectorRequirement  key:str! operator:str! values:list[str]
GetOptions                apiVersion:str kind:str resourceVersion:str
GroupKind                 group:str! kind:str!
GroupResource             group:str! resource:str!
GroupVersion              group:str! version:str!
GroupVersionForDiscovery  groupVersion:str! version:str!
GroupVersionKind          group:str! kind:str! version:str!
GroupVersionResource      group:str! resource:str! version:str!
InternalEvent             Object:object! Type:str!
LabelSelector             matchExpressions:list[TLabelSelectorRequirement] matchLabels:dict(str,str)
LabelSelectorRequirement  key:str! operator:str! values:list[str]
List                      apiVersion:str items:list[object]! kind:str metadata:TListMeta
ListMeta                  continue:str remainingItemCount:int resourceVersion:str selfLink:str
ListOptions               allowWatchBookmarks:bool apiVersion:str continue:str fieldSelector:str kind:str labelSelector:str limit:int resourceVersion:str resourceVersionMatch:str sendInitialEvents:bool timeoutSeconds:int watch:bool
ManagedFieldsEntry        apiVersion:str fieldsType:str fieldsV1:object manager:str operation:str subresource:str time:datetime
ObjectMeta                annotations:dict(str,str) creationTimestamp:datetime deletionGracePeriodSeconds:int deletionTimestamp:datetime finalizers:list[str] generateName:str generation:int labels:dict(str,str) managedFields:list[TManagedFieldsEntry] name:str namespace:str ownerReferences:list[TOwnerReference] resourceVersion:str selfLink:str uid:str
OwnerReference            apiVersion:str! blockOwnerDeletion:bool controller:bool kind:str! name:str! uid:str!
PartialObjectMetadata     apiVersion:str kind:str metadata:TObjectMeta
PartialObjectMetadataList apiVersion:str items:list[TPartialObjectMetadata]! kind:str metadata:TListMeta
PatchOptions              apiVersion:str dryRun:list[str] fieldManager:str fieldValidation:str force:bool kind:str
Preconditions             resourceVersion:str uid:str
RootPaths                 paths:list[str]!
ServerAddressByClientCIDR clientCIDR:str! serverAddress:str!
Status                    apiVersion:str code:int details:TStatusDetails kind:str message:str metadata:TListMeta reason:str status:str
StatusCause               field:str message:str reason:str
StatusDetails             causes:list[TStatusCause] group:str kind:str name:str retryAfterSeconds:int uid:str
Table                     apiVersion:str columnDefinitions:list[TTableColumnDefinition]! kind:str metadata:TListMeta rows:list[TTableRow]!
TableColumnDefinition     description:str! format:str! name:str! priority:int! type:str!
TableOptions              apiVersion:str includeObject:str kind:str
TableRow                  cells:list[object]! conditions:list[TTableRowCondition] object:object
TableRowCondition         message:str reason:str status:str! type:str!
Timestamp                 nanos:int! seconds:int!
TypeMeta                  apiVersion:str kind:str
UpdateOptions             apiVersion:str dryRun:list[str] fieldManager:str fieldValidation:str kind:str
WatchEvent                object:object! type:str!
"""

# apimachinery runtime / version types (two spellings of the package prefix in the generated SDK)
_RUNTIME = """
RuntimeTypeMeta apiVersion:str kind:str
RuntimeUnknown  ContentEncoding:str! ContentType:str! apiVersion:str kind:str
VersionInfo     buildDate:str! compiler:str! emulationMajor:str emulationMinor:str gitCommit:str! gitTreeState:str! gitVersion:str! goVersion:str! major:str! minCompatibilityMajor:str minCompatibilityMinor:str minor:str! platform:str!
"""

# pre-v2beta1 names the generated package still carries (the kubeflow/common shapes: no suspend / managedBy in RunPolicy)
_LEGACY_V1 = """
JobCondition     lastTransitionTime:datetime lastUpdateTime:datetime message:str reason:str status:str! type:str!
JobStatus        completionTime:datetime conditions:list[V1JobCondition]! lastReconcileTime:datetime replicaStatuses:dict(str,V1ReplicaStatus)! startTime:datetime
ReplicaSpec      replicas:int restartPolicy:str template:V1PodTemplateSpec
ReplicaStatus    active:int failed:int labelSelector:V1LabelSelector selector:str succeeded:int
RunPolicy        activeDeadlineSeconds:int backoffLimit:int cleanPodPolicy:str schedulingPolicy:V1SchedulingPolicy ttlSecondsAfterFinished:int
SchedulingPolicy minAvailable:int minResources:dict(str,object) priorityClass:str queue:str scheduleTimeoutSeconds:int
"""

_IRREGULAR = {"serverAddressByClientCIDRs": "server_address_by_client_cidrs", "continue": "_continue"}


def snake(name: str) -> str:
    """openapi-generator's camelCase -> snake_case (acronym runs stay together: clientCIDR -> client_cidr,
    APIGroupList -> api_group_list, fieldsV1 -> fields_v1)."""
    if name in _IRREGULAR:
        return _IRREGULAR[name]
    s = re.sub(r"([A-Z]+)([A-Z][a-z])", r"\1_\2", name)
    s = re.sub(r"([a-z])([A-Z])", r"\1_\2", s)
    s = re.sub(r"([0-9])([A-Z])", r"\1_\2", s)
    return s.lower()


def _parse(table: str) -> List[Tuple[str, List[Tuple[str, str, bool]]]]:
    out = []
    for line in table.strip().splitlines():
        name, *fields = line.split()
        parsed = []
        for f in fields:
            json_name, typ = f.split(":", 1)
            required = typ.endswith("!")
            parsed.append((json_name, typ.rstrip("!").replace("(str,", "(str, "), required))
        out.append((name, parsed))
    return out


def _build(cls_name: str, fields, type_prefix: str, doc: str):
    def sub(t: str) -> str:        # `TFoo` -> `<prefix>Foo` inside list[...] / dict(str, ...) too
        return re.sub(r"\bT(?=[A-Z])", type_prefix, t)
    ns = {
        "openapi_types": {snake(j): sub(t) for j, t, _ in fields},
        "attribute_map": {snake(j): j for j, _, _ in fields},
        "required": tuple(snake(j) for j, _, r in fields if r),
        "__doc__": doc,
        "__module__": __name__,
    }
    return type(_m.OpenApiModel)(cls_name, (_m.OpenApiModel,), ns)


META_MODELS: Dict[str, type] = {}

for _name, _fields in _parse(_META_V1):
    for _prefix in ("V1", "IoK8sApimachineryPkgApisMetaV1"):
        _cls_name = _prefix + _name
        META_MODELS[_cls_name] = _build(_cls_name, _fields, _prefix, f"meta/v1 {_name} (reference: mpijob/models/{snake(_cls_name)}.py)")
for _name, _fields in _parse(_RUNTIME):
    for _prefix in ("IoK8sApimachineryPkg", "K8sIoApimachineryPkg"):
        _cls_name = _prefix + _name
        META_MODELS[_cls_name] = _build(_cls_name, _fields, _prefix, f"apimachinery {_name} (reference: mpijob/models/{snake(_cls_name)}.py)")
for _name, _fields in _parse(_LEGACY_V1):
    META_MODELS["V1" + _name] = _build("V1" + _name, _fields, "V1", f"pre-v2beta1 {_name} (reference: mpijob/models/v1_{snake(_name)}.py)")

# The hand-listed classes of sdk/models.py that the daemon's pod templates use keep their identity (V1ObjectMeta,
# V1ListMeta, V1OwnerReference, V1LabelSelector, V1LabelSelectorRequirement and their long names): only their field lists
# are widened to the full meta/v1 shape, so `isinstance` checks and `MODEL_CLASSES` users see ONE class per name.
for _cls_name, _cls in list(META_MODELS.items()):
    _have = _m.MODEL_CLASSES.get(_cls_name)
    if _have is not None:
        for _attr, _typ in _cls.openapi_types.items():
            if _attr not in _have.openapi_types:
                _have.openapi_types[_attr] = _typ
                _have.attribute_map[_attr] = _cls.attribute_map[_attr]
                setattr(_have, _attr, type(_m.OpenApiModel)._make_property(_attr, False))
        META_MODELS[_cls_name] = _have
    else:
        _m.MODEL_CLASSES[_cls_name] = _cls
        setattr(_m, _cls_name, _cls)

globals().update(META_MODELS)
__all__ = sorted(META_MODELS)
