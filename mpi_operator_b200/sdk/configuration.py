"""Client configuration (reference: sdk/python/v2beta1/mpijob/configuration.py:32-447).
The host is the single-box daemon's REST endpoint instead of a kube-apiserver."""
from __future__ import annotations

import copy
import logging
import os
import sys


def default_token() -> str:
    """Bearer token for a daemon started with --auth-token-file: $MPIJOB_TOKEN, else the content of $MPIJOB_TOKEN_FILE."""
    tok = os.environ.get("MPIJOB_TOKEN", "")
    path = os.environ.get("MPIJOB_TOKEN_FILE", "")
    if not tok and path:
        try:
            with open(path) as f:
                tok = f.read().strip()
        except OSError:
            tok = ""
    return tok


class Configuration:
    _default = None

    def __init__(self, host: str = None, api_key=None, api_key_prefix=None, username=None, password=None,
                 discard_unknown_keys: bool = False):
        self.host = host or ("http://" + os.environ.get("MPIJOB_SERVER", "127.0.0.1:8087"))
        self.temp_folder_path = None
        self.api_key = dict(api_key or {})
        self.api_key_prefix = dict(api_key_prefix or {})
        if "authorization" not in self.api_key:   # daemon started with --auth-token-file: MPIJOB_TOKEN or MPIJOB_TOKEN_FILE
            tok = default_token()
            if tok:
                self.api_key["authorization"], self.api_key_prefix["authorization"] = tok, "Bearer"
        self.refresh_api_key_hook = None
        self.username, self.password = username, password
        self.discard_unknown_keys = discard_unknown_keys
        self.logger = {"package_logger": logging.getLogger("mpijob"), "urllib3_logger": logging.getLogger("urllib3")}
        self.logger_format = "%(asctime)s %(levelname)s %(message)s"
        self.logger_stream_handler = None
        self.logger_file_handler = None
        self._logger_file = None
        self._debug = False
        self.verify_ssl = True
        self.ssl_ca_cert = self.cert_file = self.key_file = None
        self.assert_hostname = None
        self.connection_pool_maxsize = 4
        self.proxy = self.proxy_headers = None
        self.safe_chars_for_path_param = ""
        self.retries = None
        self.client_side_validation = True
        self.timeout = 30.0

    def __deepcopy__(self, memo):
        cls = self.__class__
        result = cls.__new__(cls)
        memo[id(self)] = result
        for k, v in self.__dict__.items():
            if k not in ("logger", "logger_file_handler", "logger_stream_handler"):
                setattr(result, k, copy.deepcopy(v, memo))
        result.logger = copy.copy(self.logger)
        result.logger_file_handler = result.logger_stream_handler = None
        return result

    @classmethod
    def set_default(cls, default):
        cls._default = copy.deepcopy(default)

    @classmethod
    def get_default_copy(cls):
        return copy.deepcopy(cls._default) if cls._default is not None else Configuration()

    @property
    def debug(self):
        return self._debug

    @debug.setter
    def debug(self, value):
        self._debug = value
        for lg in self.logger.values():
            lg.setLevel(logging.DEBUG if value else logging.WARNING)

    @property
    def logger_file(self):
        return self._logger_file

    @logger_file.setter
    def logger_file(self, value):
        self._logger_file = value
        if value:
            self.logger_file_handler = logging.FileHandler(value)
            self.logger_file_handler.setFormatter(logging.Formatter(self.logger_format))
            for lg in self.logger.values():
                lg.addHandler(self.logger_file_handler)

    def get_api_key_with_prefix(self, identifier, alias=None):
        if self.refresh_api_key_hook is not None:
            self.refresh_api_key_hook(self)
        key = self.api_key.get(identifier, self.api_key.get(alias) if alias is not None else None)
        if key:
            prefix = self.api_key_prefix.get(identifier)
            return f"{prefix} {key}" if prefix else key

    def get_basic_auth_token(self):
        import base64
        return "Basic " + base64.b64encode(f"{self.username or ''}:{self.password or ''}".encode()).decode()

    def auth_settings(self):
        tok = self.get_api_key_with_prefix("authorization")
        return {"BearerToken": {"type": "api_key", "in": "header", "key": "authorization", "value": tok}} if tok else {}

    def auth_headers(self) -> dict:
        tok = self.get_api_key_with_prefix("authorization")
        return {"Authorization": tok} if tok else {}

    def to_debug_report(self):
        return ("Python SDK Debug Report:\n"
                f"OS: {sys.platform}\nPython Version: {sys.version}\nVersion of the API: v2beta1\nSDK Package Version: 0.4.0")

    def get_host_settings(self):
        return [{"url": self.host, "description": "single-box MPIJob operator daemon"}]

    def get_host_from_settings(self, index, variables=None, servers=None):
        servers = servers or self.get_host_settings()
        try:
            return servers[index or 0]["url"]
        except IndexError:
            raise ValueError(f"Invalid index {index} when selecting the host settings. Must be less than {len(servers)}")
