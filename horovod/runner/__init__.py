"""``horovod.runner``: programmatic launch. ``horovod.runner.run(fn, args=(), kwargs={}, np=2)`` runs ``fn`` on ``np`` ranks under
the native mpirun (each rank executes it after ``hvd.init()`` is available to it) and returns the per-rank results."""
from __future__ import annotations

import os
import pickle
import subprocess
import sys
import tempfile
from typing import Any, Callable, Dict, List, Optional

_RUNNER = """
import pickle, sys
sys.path[:0] = {path!r}
fn, args, kwargs = pickle.load(open({inp!r}, 'rb'))
import os
rank = int(os.environ.get('HOROVOD_RANK', os.environ.get('OMPI_COMM_WORLD_RANK', '0')))
out = fn(*args, **kwargs)
pickle.dump(out, open({outdir!r} + '/result.' + str(rank), 'wb'))
"""


def run(func: Callable, args=(), kwargs: Optional[Dict[str, Any]] = None, np: int = 1, hosts: Optional[str] = None,
        env: Optional[Dict[str, str]] = None, use_mpi: Optional[bool] = None, use_gloo: Optional[bool] = None, verbose: bool = False,
        start_timeout: Optional[int] = None) -> List[Any]:
    from mpi_operator_b200.cmd.horovodrun import MPIRUN
    try:
        import cloudpickle as pk
    except ImportError:   # plain pickle works for importable top-level functions
        pk = pickle
    with tempfile.TemporaryDirectory(prefix="hvdrun-") as d:
        inp = os.path.join(d, "call.pkl")
        with open(inp, "wb") as f:
            pk.dump((func, tuple(args), dict(kwargs or {})), f)
        code = _RUNNER.format(path=[p for p in sys.path if p], inp=inp, outdir=d)
        argv = [str(MPIRUN), "-np", str(np)] + (["-H", hosts] if hosts else []) + [sys.executable, "-c", code]
        e = dict(os.environ)
        e.update(env or {})
        if start_timeout:
            e["B200MPI_INIT_TIMEOUT_MS"] = str(start_timeout * 1000)
        r = subprocess.run(argv, env=e, capture_output=not verbose, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"horovod.runner.run: ranks failed with exit code {r.returncode}\\n{(r.stderr or '')[-2000:]}")
        return [pickle.load(open(os.path.join(d, f"result.{k}"), "rb")) for k in range(np)]
