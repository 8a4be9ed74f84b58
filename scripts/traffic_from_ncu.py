"""profiles/rNN_traffic.json from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` log of one eager step (scripts/ncu_step.py): per-launch DRAM traffic averaged by kernel family.

  python scripts/traffic_from_ncu.py gpurun_out/traffic.csv profiles/r01_traffic.json
"""
import collections
import csv
import json
import re
import sys

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "nsecond": 1e-9, "us": 1e-6,
         "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3}
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = collections.defaultdict(dict)
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    per[(name, r["ID"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * SCALE.get(r["Metric Unit"], 1)
fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (name, _), m in per.items():
    if "gemm_tc_kernel" in name:
        f = "conv3x3_igemm" if re.search(r"<\d+, \d+, (\d), \d>", name).group(1) == "1" else "gemm"
    elif "attn_fwd" in name:
        f = "attn_fwd"
    elif "attn_bwd" in name:
        f = "attn_bwd"
    else:
        continue
    a = fam[f]
    a[0] += 1
    a[1] += m.get("dram__bytes_read.sum", 0)
    a[2] += m.get("dram__bytes_write.sum", 0)
    a[3] += m.get("gpu__time_duration.sum", 0)
out = {f: {"launches": a[0], "dram_read_bytes_per_launch": a[1] / a[0], "dram_write_bytes_per_launch": a[2] / a[0],
           "traffic_bytes_per_launch": (a[1] + a[2]) / a[0], "ncu_ms_total": a[3] * 1e3} for f, a in fam.items()}
out["_source"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one eager TextBoost step "
                  "(scripts/ncu_step.py), SD-1.5 B=8, KPL on; per-launch averages by kernel family (cold caches)")
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
