"""Summarises an ncu capture for profiles/: usage
    ncu -i prof.ncu-rep --page raw --csv > raw.csv
    ncu -i prof.ncu-rep --page source --csv --print-source=sass > sass.csv
    python profiles/ncu_summary.py raw.csv sass.csv launches.csv "title" > profiles/rNN_x_summary.md
"""
import collections
import csv
import sys

raw, sass, launches, title = sys.argv[1:5]
n_warps = int(sys.argv[5]) if len(sys.argv) > 5 else 32768
print(f"# {title}\n")
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) > mv:
        try:
            tot[r[kn]] += float(r[mv].replace(",", "")); cnt[r[kn]] += 1
        except ValueError:
            pass
s = sum(tot.values())
print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`), share of GPU time\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
    print(f"| `{k[:95]}` | {cnt[k]} | {v / 1e3:.1f} | {v / s * 100:.1f}% |")
rows = list(csv.reader(open(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
print("\n## `ncu --set full` of the step kernel (first captured launch)\n\n| metric | value |\n|---|---|")
for w in want:
    for i, hh in enumerate(hdr):
        if hh == w:
            print(f"| {w} [{units[i]}] | {d[i]} |")
rows = list(csv.reader(open(sass)))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"rows": []}; kern.append(cur); continue
    if r and r[0] == "Address":
        cur["hdr"] = r; continue
    if cur is not None and len(r) > 5:
        cur["rows"].append(r)
k = kern[0]
h = k["hdr"]
ie, src = h.index("Instructions Executed"), h.index("Source")
total = sum(float(r[ie]) for r in k["rows"])
hist = collections.Counter()
for r in k["rows"]:
    t = r[src].split()
    op = t[1] if t[0].startswith("@") else t[0]
    hist[op.split(".")[0]] += float(r[ie])
print(f"| SASS instructions in the kernel | {len(k['rows'])} |")
print(f"| warp instructions executed per warp-step ({n_warps} warps) | {total / n_warps:.0f} |")
print("| top opcodes per warp-step | " + ", ".join(f"{op} {c / n_warps:.0f}" for op, c in hist.most_common(24)) + " |")
st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
ss = collections.Counter()
for r in k["rows"]:
    for i in st:
        try:
            ss[h[i]] += float(r[i])
        except ValueError:
            pass
tt = sum(ss.values())
print("| warp stall samples | " + ", ".join(f"{a[6:]} {v / tt * 100:.1f}%" for a, v in ss.most_common(8)) + " |")
