"""``kubernetes.client.rest.ApiException`` (scripts catch it around create/get calls)."""
from mpi_operator_b200.sdk.exceptions import ApiException  # noqa: F401
