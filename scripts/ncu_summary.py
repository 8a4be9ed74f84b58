"""Summarise an .ncu-rep here (no GPU needed): per kernel, the headline metrics and the top stall lines.

  python scripts/ncu_summary.py gpurun_out/rNN_x.ncu-rep [--top 12] > profiles/rNN_x_summary.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__registers_per_thread",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
col = {h: i for i, h in enumerate(hdr)}
kern = []
for r in raw[2:]:
    kern.append(r)
    print(f"== [{r[col['ID']]}] {r[col['Kernel Name']][:100]}")
    for w in WANT:
        if w in col:
            print(f"   {w:68s} {r[col[w]]:>14s} {units[col[w]]}")
STALLS = "smsp__pcsamp_warps_issue_stalled_"
print("\n== stall reasons (pc sampling, % of samples) and issue rate")
for r in kern:
    items = [(h[len(STALLS):], float(r[i] or 0)) for h, i in col.items() if h.startswith(STALLS) and "not_issued" not in h]
    tot = sum(v for _, v in items) or 1
    topr = ", ".join(f"{n} {100 * v / tot:.0f}%" for n, v in sorted(items, key=lambda x: -x[1])[:6])
    ipc = r[col["sm__inst_executed.avg.per_cycle_active"]] if "sm__inst_executed.avg.per_cycle_active" in col else "?"
    inst = r[col["smsp__inst_executed.sum"]] if "smsp__inst_executed.sum" in col else "?"
    print(f"   [{r[col['ID']]}] IPC/SM {ipc}  warp-instr {inst}: {topr}")
src = run("--page", "source", "--csv")
blocks = src.split('"Kernel Name",')[1:]
for k, blk in enumerate(blocks):
    rows = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
    h = rows[1]
    ends = [i for i, r in enumerate(rows) if i > 1 and r and r[0] == "Address"]  # a second view repeats the table
    body = [r for r in rows[2:(ends[0] if ends else len(rows))] if len(r) == len(h)]
    i_src, i_s = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
    i_ex = h.index("Instructions Executed")
    tot = sum(int(r[i_s]) for r in body) or 1
    print(f"\n== stall samples, kernel {k}: {rows[0][1][:90]}  (total {tot})")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][i_s]))[:top]
    for i in sorted(order):
        r = body[i]
        print(f"   {i:5d} {int(r[i_s]):7d} {100 * int(r[i_s]) / tot:5.1f}%  exec={r[i_ex]:>8s}  {r[i_src].strip()[:100]}")
