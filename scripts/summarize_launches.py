"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel GPU time for ONE step."""
import collections
import csv
import json
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def to_ns(d):
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    return v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)


def main(path, out_json=None):
    rows = load(path)
    names = [d["Kernel Name"] for d in rows]
    marks = [i for i, n in enumerate(names) if "timestep_embedding" in n]
    # a step starts at its first timestep_embedding launch; take the last complete one
    starts = [m for j, m in enumerate(marks) if j == 0 or m - marks[j - 1] > 50]
    if len(starts) >= 2:
        lo, hi = starts[-2], starts[-1]
    else:                       # the capture window was exactly one step (bench.py --profile-step)
        lo, hi = 0, len(rows)
    step = rows[lo:hi]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for d in step:
        n = d["Kernel Name"].split("(")[0]
        if "gemm_tcgen05" in n or "gemm2_kernel" in n:
            n = f"{n.split('<')[0]} grid={d['Grid Size']}"
        ns = to_ns(d)
        agg[n][0] += 1
        agg[n][1] += ns
        tot += ns
    out = {"launches": len(step), "total_ms": tot / 1e6,
           "kernels": [{"name": k, "count": v[0], "ms": v[1] / 1e6, "share": v[1] / tot}
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    gemm = sum(k["ms"] for k in out["kernels"] if "gemm_tcgen05" in k["name"] or "gemm2_kernel" in k["name"])
    out["gemm_tcgen05_ms"], out["gemm_tcgen05_share"] = gemm, gemm / (tot / 1e6)
    print(f"launches {out['launches']}  total {out['total_ms']:.2f} ms  gemm_tcgen05 {gemm:.2f} ms ({out['gemm_tcgen05_share']:.1%})")
    for k in out["kernels"][:32]:
        print(f"{k['name'][:72]:72s} n={k['count']:5d} ms={k['ms']:8.2f} avg_us={k['ms'] / k['count'] * 1e3:8.1f} {k['share']:6.1%}")
    if out_json:
        json.dump(out, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
