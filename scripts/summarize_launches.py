"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name:
python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    tot[name][0] += 1
    tot[name][1] += v * scale
total = sum(v[1] for v in tot.values())
print(f"total {total:.3f} ms over {sum(v[0] for v in tot.values())} launches (ncu per-launch times: cold-cache, serialised; compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {ms:10.3f} {100 * ms / total:6.1f}% {1e3 * ms / n:9.1f}")
