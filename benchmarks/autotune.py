"""Derive the allreduce algorithm crossover table from measured sweeps
(profiles/allreduce_sweep_n*_f32.json) and write mpi_operator_b200/runtime/tuning.json, which
Communicator.create() applies ("the algorithm crossover table produced by measurement
and stored in a tuning JSON"). Usage: python benchmarks/autotune.py profiles/allreduce_sweep_n*_f32.json"""
import json
import os
import sys


def crossovers(rows):
    by_size = {}
    for r in rows:
        base = r["algo"].split("@")[0]
        if base in ("oneshot", "twoshot", "nvls"):
            cur = by_size.setdefault(r["bytes"], {})
            cur[base] = min(cur.get(base, 1e9), r["ms_median_max_over_ranks"])
    sizes = sorted(by_size)
    oneshot_max = 0
    for s in sizes:
        t = by_size[s]
        if "oneshot" in t and t["oneshot"] <= min(v for k, v in t.items() if k != "oneshot" or len(t) == 1):
            oneshot_max = s
    nvls_min = None
    for s in reversed(sizes):
        t = by_size[s]
        if "nvls" in t and "twoshot" in t and t["nvls"] <= t["twoshot"]:
            nvls_min = s
        elif "nvls" in t and "twoshot" in t and s > oneshot_max:
            break
    blocks = {}
    for r in rows:
        if "@" in r["algo"] and r["bytes"] >= (64 << 20):
            base, nb = r["algo"].split("@")
            blocks.setdefault(base, {}).setdefault(int(nb), []).append(r["busbw_gbs"])
    best_blocks = {b: max(v, key=lambda k: sum(v[k]) / len(v[k])) for b, v in blocks.items()}
    return {"oneshot_max_bytes": oneshot_max, "nvls_min_bytes": nvls_min if nvls_min is not None else -1, "best_blocks": best_blocks}


def main():
    out = {}
    for path in sys.argv[1:]:
        d = json.load(open(path))
        out[str(d["world"])] = crossovers(d["rows"])
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mpi_operator_b200", "runtime", "tuning.json")
    json.dump({"source": [os.path.basename(p) for p in sys.argv[1:]], "by_world": out}, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
