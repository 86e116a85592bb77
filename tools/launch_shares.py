"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share of the command's GPU time, and the
launches of ONE -similar step (from one mih2_hist_all_kernel launch to the next) in order.

    python tools/launch_shares.py profiles/launches_bench_r02.csv > profiles/launch_shares_r02.json
"""
import csv
import json
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*", "", name) if name.startswith("cub::") or "at::native" in name else name
    rows.append((name, float(r["Metric Value"]) / 1e6, r["Grid Size"], r["Block Size"]))
total = sum(ms for _, ms, _, _ in rows)
agg = {}
for name, ms, _, _ in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
starts = [i for i, r in enumerate(rows) if "mih2_hist_all_kernel" in r[0]]
step = []
if len(starts) >= 3:
    a, b = starts[-2], starts[-1]  # a device-resident timed step (the e2e steps at the end add copies, not kernels)
    step = [{"kernel": n, "ms": round(ms, 4), "grid": g, "block": bl} for n, ms, g, bl in rows[a:b] if "at::native" not in n]
out = {
    "command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 --no-extras "
               "(10^7 rows, dht 5)",
    "note": "cold-cache, serialised launch times under the profiler: compare SHARES, not absolutes. Final round-2 build.",
    "total_ms": total,
    "kernels": [{"kernel": k, "launches": v[0], "ms": v[1], "share": v[1] / total} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])],
    "one_similar_step": {"ms": sum(x["ms"] for x in step), "launches": step},
}
print(json.dumps(out, indent=1))
