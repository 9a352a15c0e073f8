import os
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
HERE = os.path.dirname(os.path.abspath(__file__))


def test_multiprocess_collectives():
    sys.path.insert(0, HERE)
    from mp_launch import launch
    n = min(torch.cuda.device_count(), 8)
    rcs = launch(n, [os.path.join(HERE, "mp_worker.py")], timeout=240)
    assert rcs == [0] * n
