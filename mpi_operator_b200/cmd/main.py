"""``python -m mpi_operator_b200.cmd.main`` — the operator binary
(reference: cmd/mpi-operator/main.go:42-53)."""
import sys

from .options import parse
from .server import run


def main(argv=None) -> int:
    return run(parse(argv))


if __name__ == "__main__":
    sys.exit(main())
