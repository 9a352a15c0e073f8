"""API error model (the subset of k8s.io/apimachinery/pkg/api/errors the
reference's controller branches on: IsNotFound, IsAlreadyExists, IsConflict,
and the 401/403 watch errors that are fatal at controller.go:374-388)."""
from __future__ import annotations


class ApiError(Exception):
    def __init__(self, reason: str, message: str, code: int):
        super().__init__(message)
        self.reason, self.message, self.code = reason, message, code

    def to_status(self) -> dict:
        return {"kind": "Status", "apiVersion": "v1", "status": "Failure", "message": self.message,
                "reason": self.reason, "code": self.code}


def not_found(resource: str, name: str) -> ApiError:
    return ApiError("NotFound", f'{resource} "{name}" not found', 404)


def already_exists(resource: str, name: str) -> ApiError:
    return ApiError("AlreadyExists", f'{resource} "{name}" already exists', 409)


def conflict(resource: str, name: str, why: str = "the object has been modified; please apply your changes to the latest version and try again") -> ApiError:
    return ApiError("Conflict", f'Operation cannot be fulfilled on {resource} "{name}": {why}', 409)


def invalid(resource: str, name: str, why: str) -> ApiError:
    return ApiError("Invalid", f'{resource} "{name}" is invalid: {why}', 422)


def bad_request(why: str) -> ApiError:
    return ApiError("BadRequest", why, 400)


def forbidden(why: str) -> ApiError:
    return ApiError("Forbidden", why, 403)


def unauthorized(why: str = "Unauthorized") -> ApiError:
    return ApiError("Unauthorized", why, 401)


def is_not_found(e: Exception) -> bool:
    return isinstance(e, ApiError) and e.reason == "NotFound"


def is_already_exists(e: Exception) -> bool:
    return isinstance(e, ApiError) and e.reason == "AlreadyExists"


def is_conflict(e: Exception) -> bool:
    return isinstance(e, ApiError) and e.reason == "Conflict"
