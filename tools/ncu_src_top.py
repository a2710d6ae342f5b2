"""Top stall-sampled SASS instructions of one kernel of an ncu report's source page:
   ncu -i rep.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv; python tools/ncu_src_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "# Samples" in r)
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = [r for r in rows if len(r) == len(hdr) and r[iS].isdigit()]
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for j in stall:
        agg[hdr[j]] = agg.get(hdr[j], 0) + int(r[j])
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:n]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[j]), hdr[j]) for j in stall), reverse=True)[:2]
    print(i, r[iS], r[iEx], r[iSrc].strip()[:80], st)
