"""HTTP transport (reference: sdk/python/v2beta1/mpijob/rest.py:49-298, there urllib3;
here the stdlib http.client — the daemon is always a plain local HTTP endpoint)."""
from __future__ import annotations

import http.client
import json
import urllib.parse

from .exceptions import ApiException, ApiValueError, ForbiddenException, NotFoundException, ServiceException, UnauthorizedException


class RESTResponse:
    def __init__(self, status, reason, data, headers):
        self.status, self.reason, self.data, self._headers = status, reason, data, headers

    def getheaders(self):
        return dict(self._headers)

    def getheader(self, name, default=None):
        return dict(self._headers).get(name, default)


class RESTClientObject:
    def __init__(self, configuration, pools_size=4, maxsize=None):
        self.configuration = configuration

    def request(self, method, url, query_params=None, headers=None, body=None, post_params=None,
                _preload_content=True, _request_timeout=None):
        method = method.upper()
        if method not in ("GET", "HEAD", "DELETE", "POST", "PUT", "PATCH", "OPTIONS"):
            raise ApiValueError(f"unsupported method {method}")
        if post_params and body:
            raise ApiValueError("body parameter cannot be used with post_params parameter.")
        u = urllib.parse.urlparse(url)
        path = u.path + ("?" + urllib.parse.urlencode(query_params) if query_params else ("?" + u.query if u.query else ""))
        headers = dict(headers or {})
        data = None
        if body is not None:
            data = body if isinstance(body, (bytes, str)) else json.dumps(body)
            headers.setdefault("Content-Type", "application/json")
        timeout = _request_timeout if isinstance(_request_timeout, (int, float)) else self.configuration.timeout
        conn = http.client.HTTPConnection(u.hostname, u.port or 80, timeout=timeout)
        try:
            conn.request(method, path, body=data, headers=headers)
            r = conn.getresponse()
            resp = RESTResponse(r.status, r.reason, r.read(), r.getheaders())
        except OSError as e:
            raise ApiException(status=0, reason=f"{type(e).__name__}: {e} (is the mpi-operator daemon running at {url}?)")
        finally:
            conn.close()
        if not 200 <= resp.status <= 299:
            cls = {401: UnauthorizedException, 403: ForbiddenException, 404: NotFoundException}.get(resp.status)
            if cls is None:
                cls = ServiceException if 500 <= resp.status <= 599 else ApiException
            raise cls(http_resp=resp)
        return resp

    def GET(self, url, **kw): return self.request("GET", url, **kw)  # noqa: E704,N802
    def HEAD(self, url, **kw): return self.request("HEAD", url, **kw)  # noqa: E704,N802
    def OPTIONS(self, url, **kw): return self.request("OPTIONS", url, **kw)  # noqa: E704,N802
    def DELETE(self, url, **kw): return self.request("DELETE", url, **kw)  # noqa: E704,N802
    def POST(self, url, **kw): return self.request("POST", url, **kw)  # noqa: E704,N802
    def PUT(self, url, **kw): return self.request("PUT", url, **kw)  # noqa: E704,N802
    def PATCH(self, url, **kw): return self.request("PATCH", url, **kw)  # noqa: E704,N802
