"""NVTX ranges for Nsight Systems / ncu timelines (SURVEY.md §5.1 [NEW]); opt-in with ``B200MPI_NVTX=1`` so that the default
path pays nothing. ``with nvtx.range("backward"):`` around host-side phases; device work launched inside shows up under it."""
from __future__ import annotations

import contextlib
import os

ENABLED = os.environ.get("B200MPI_NVTX", "0") == "1"


@contextlib.contextmanager
def range(name: str):   # noqa: A001 (mirrors torch.cuda.nvtx.range)
    if not ENABLED:
        yield
        return
    import torch
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def mark(name: str) -> None:
    if ENABLED:
        import torch
        torch.cuda.nvtx.mark(name)
