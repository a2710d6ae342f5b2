"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python tools/launch_summary.py gpurun_out/launches.csv 'command line that was profiled' > profiles/rNN_launch_list_summary.csv"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
agg = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv or "gpu__time_duration" not in r[mn]:
        continue
    name = re.sub(r"\(.*$", "", r[kn]).replace("qie::", "").strip()
    t = float(r[mv].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
unit_ns = True
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
scale = 1e-6 if tot > 1e6 else 1e-3      # ns or us -> ms
print(f"# ncu launch list — `--metrics gpu__time_duration.sum --clock-control none`")
print(f"# command: {sys.argv[2] if len(sys.argv) > 2 else ''}  (cold-cache, serialised: compare SHARES)")
print(f"# total kernel time {tot * scale:.1f} ms over {n} launches")
print("kernel,launches,total_ms,share,avg_ms")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{c},{t * scale:.3f},{t / tot:.4f},{t * scale / c:.4f}")
