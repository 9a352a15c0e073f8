"""Collective trace -> Chrome trace (chrome://tracing / Perfetto).

``Communicator.trace(True)`` makes the runtime record one entry per collective (op, bytes, algorithm,
CTAs, host enqueue time); ``trace_dump(path)`` appends them as JSONL (one file may hold many ranks).
This converts that JSONL into the Trace Event format: one process per rank, one instant/complete event
per collective — the Horovod-Timeline analogue the reference only lists as "consider" (ROADMAP.md:14).

    python -m mpi_operator_b200.utils.trace trace.jsonl trace.json
"""
from __future__ import annotations

import json
import sys
from typing import Iterable, List


def to_chrome_trace(records: Iterable[dict]) -> dict:
    recs = sorted(records, key=lambda r: (r.get("rank", 0), r.get("t_ns", 0)))
    t0 = min((r.get("t_ns", 0) for r in recs), default=0)
    events: List[dict] = []
    ranks = sorted({r.get("rank", 0) for r in recs})
    for rk in ranks:
        events.append({"ph": "M", "pid": rk, "name": "process_name", "args": {"name": f"rank {rk}"}})
    by_rank = {rk: [r for r in recs if r.get("rank", 0) == rk] for rk in ranks}
    for rk, rs in by_rank.items():
        for i, r in enumerate(rs):
            ts = (r["t_ns"] - t0) / 1e3
            nxt = (rs[i + 1]["t_ns"] - t0) / 1e3 if i + 1 < len(rs) else ts + 1.0
            events.append({"ph": "X", "pid": rk, "tid": 0, "ts": ts, "dur": max(min(nxt - ts, 50.0), 0.5),
                           "name": f"{r['op']}[{r.get('algo', 'auto')}]", "cat": "collective",
                           "args": {"bytes": r.get("bytes", 0), "blocks": r.get("blocks", 0)}})
    return {"traceEvents": events, "displayTimeUnit": "ns"}


def load_jsonl(path: str) -> List[dict]:
    out = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                out.append(json.loads(line))
    return out


def summarize(records: Iterable[dict]) -> dict:
    agg: dict = {}
    for r in records:
        k = (r["op"], r.get("algo", "auto"))
        a = agg.setdefault(k, {"calls": 0, "bytes": 0})
        a["calls"] += 1
        a["bytes"] += r.get("bytes", 0)
    return {f"{op}[{algo}]": v for (op, algo), v in sorted(agg.items())}


def main(argv=None) -> int:
    argv = argv or sys.argv[1:]
    if len(argv) < 2:
        print(__doc__)
        return 2
    recs = load_jsonl(argv[0])
    with open(argv[1], "w") as f:
        json.dump(to_chrome_trace(recs), f)
    print(json.dumps(summarize(recs), indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
