"""SDK exception hierarchy (reference: sdk/python/v2beta1/mpijob/exceptions.py:16-163)."""


class OpenApiException(Exception):
    """The base exception class for all SDK errors."""


class ApiTypeError(OpenApiException, TypeError):
    def __init__(self, msg, path_to_item=None, valid_classes=None, key_type=None):
        self.path_to_item, self.valid_classes, self.key_type = path_to_item, valid_classes, key_type
        super().__init__(f"{msg} at {render_path(path_to_item)}" if path_to_item else msg)


class ApiValueError(OpenApiException, ValueError):
    def __init__(self, msg, path_to_item=None):
        self.path_to_item = path_to_item
        super().__init__(f"{msg} at {render_path(path_to_item)}" if path_to_item else msg)


class ApiAttributeError(OpenApiException, AttributeError):
    def __init__(self, msg, path_to_item=None):
        self.path_to_item = path_to_item
        super().__init__(f"{msg} at {render_path(path_to_item)}" if path_to_item else msg)


class ApiKeyError(OpenApiException, KeyError):
    def __init__(self, msg, path_to_item=None):
        self.path_to_item = path_to_item
        super().__init__(f"{msg} at {render_path(path_to_item)}" if path_to_item else msg)


class ApiException(OpenApiException):
    def __init__(self, status=None, reason=None, http_resp=None, body=None):
        if http_resp is not None:
            self.status, self.reason, self.body, self.headers = http_resp.status, http_resp.reason, http_resp.data, http_resp.getheaders()
        else:
            self.status, self.reason, self.body, self.headers = status, reason, body, None
        super().__init__(str(self))

    def __str__(self):
        msg = f"({self.status})\nReason: {self.reason}\n"
        if self.headers:
            msg += f"HTTP response headers: {self.headers}\n"
        if self.body:
            msg += f"HTTP response body: {self.body}\n"
        return msg


class NotFoundException(ApiException):
    pass


class UnauthorizedException(ApiException):
    pass


class ForbiddenException(ApiException):
    pass


class ServiceException(ApiException):
    pass


def render_path(path_to_item):
    """Returns a string representation of a path, e.g. ['a'][0]['b']."""
    return "".join(f"[{p}]" if isinstance(p, int) else f"['{p}']" for p in (path_to_item or []))
