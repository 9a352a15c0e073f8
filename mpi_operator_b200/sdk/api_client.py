"""Generic API client: (de)serialisation between SDK models and JSON + request
plumbing (reference: sdk/python/v2beta1/mpijob/api_client.py:223-263
``sanitize_for_serialization``, :265-324 ``deserialize``, :326 ``call_api``)."""
from __future__ import annotations

import datetime
import json
import re
from typing import Any

from . import models as M
from .configuration import Configuration
from .exceptions import ApiValueError
from .rest import RESTClientObject


class ApiClient:
    PRIMITIVE_TYPES = (float, bool, bytes, str, int)
    NATIVE_TYPES_MAPPING = {"int": int, "float": float, "str": str, "bool": bool, "object": object}

    def __init__(self, configuration: Configuration = None, header_name=None, header_value=None, cookie=None, pool_threads=1):
        self.configuration = configuration or Configuration.get_default_copy()
        self.rest_client = RESTClientObject(self.configuration)
        self.default_headers = {"User-Agent": "OpenAPI-Generator/0.4.0/python (b200)"}
        if header_name is not None:
            self.default_headers[header_name] = header_value
        self.cookie = cookie
        self.client_side_validation = self.configuration.client_side_validation

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        pass

    @property
    def user_agent(self):
        return self.default_headers["User-Agent"]

    @user_agent.setter
    def user_agent(self, v):
        self.default_headers["User-Agent"] = v

    def set_default_header(self, name, value):
        self.default_headers[name] = value

    # ------------------------------------------------------------- encoding --
    def sanitize_for_serialization(self, obj):
        """Model/dict/list tree -> JSON-ready structure using each model's attribute_map; None fields dropped."""
        if obj is None:
            return None
        if isinstance(obj, self.PRIMITIVE_TYPES):
            return obj
        if isinstance(obj, (list, tuple)):
            return [self.sanitize_for_serialization(x) for x in obj]
        if isinstance(obj, (datetime.datetime, datetime.date)):
            return obj.isoformat()
        if isinstance(obj, dict):
            return {k: self.sanitize_for_serialization(v) for k, v in obj.items()}
        if hasattr(obj, "openapi_types") and hasattr(obj, "attribute_map"):
            return {obj.attribute_map[attr]: self.sanitize_for_serialization(getattr(obj, attr))
                    for attr in obj.openapi_types if getattr(obj, attr) is not None}
        if hasattr(obj, "to_dict"):
            return self.sanitize_for_serialization(obj.to_dict())
        raise ApiValueError(f"cannot serialise object of type {type(obj).__name__}")

    def deserialize(self, response, response_type):
        data = response.data if hasattr(response, "data") else response
        if isinstance(data, (bytes, str)):
            try:
                data = json.loads(data)
            except ValueError:
                pass
        return self._deserialize(data, response_type)

    def _deserialize(self, data, klass):
        if data is None:
            return None
        if isinstance(klass, str):
            m = re.match(r"list\[(.*)\]", klass)
            if m:
                return [self._deserialize(x, m.group(1)) for x in data]
            m = re.match(r"dict\(([^,]*), (.*)\)", klass)
            if m:
                return {k: self._deserialize(v, m.group(2)) for k, v in data.items()}
            if klass in self.NATIVE_TYPES_MAPPING:
                klass = self.NATIVE_TYPES_MAPPING[klass]
            elif klass == "datetime":
                return data
            else:
                klass = M.MODEL_CLASSES.get(klass, object)
        if klass in (int, float, str, bool):
            try:
                return klass(data)
            except (TypeError, ValueError):
                return data
        if klass is object:
            return data
        if not isinstance(data, dict):
            return data
        kwargs = {}
        for attr, typ in klass.openapi_types.items():
            key = klass.attribute_map[attr]
            if key in data and data[key] is not None:
                kwargs[attr] = self._deserialize(data[key], typ)
        for r in klass.required:
            kwargs.setdefault(r, [] if klass.openapi_types[r].startswith("list") else ({} if klass.openapi_types[r].startswith("dict") else None))
        cfg = Configuration.get_default_copy()
        cfg.client_side_validation = False
        return klass(local_vars_configuration=cfg, **kwargs)

    # -------------------------------------------------------------- requests --
    def call_api(self, resource_path, method, path_params=None, query_params=None, header_params=None, body=None,
                 post_params=None, files=None, response_type=None, auth_settings=None, async_req=None,
                 _return_http_data_only=True, collection_formats=None, _preload_content=True, _request_timeout=None,
                 _host=None):
        for k, v in (path_params or {}).items():
            resource_path = resource_path.replace("{%s}" % k, str(v))
        headers = dict(self.default_headers)
        headers.update(self.configuration.auth_headers() if hasattr(self.configuration, "auth_headers") else {})
        headers.update(header_params or {})
        url = (_host or self.configuration.host) + resource_path
        resp = self.rest_client.request(method, url, query_params=query_params, headers=headers,
                                        body=self.sanitize_for_serialization(body), _request_timeout=_request_timeout)
        if response_type == "raw":
            return resp.data
        data = self.deserialize(resp, response_type) if response_type else json.loads(resp.data or b"null")
        return data if _return_http_data_only else (data, resp.status, resp.getheaders())

    def select_header_accept(self, accepts):
        return "application/json" if not accepts or "application/json" in [a.lower() for a in accepts] else ", ".join(accepts)

    def select_header_content_type(self, content_types):
        return "application/json"
