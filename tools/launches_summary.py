"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.

    python tools/launches_summary.py gpurun_out/launches.csv [first_id last_id]   # optional ID window = one step
"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
with open(path) as f:
    rows = list(csv.reader(l for l in f if l.startswith('"')))
hdr, data = rows[0], rows[1:]
ik, iv, iid, ig, ib = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "ID", "Grid Size", "Block Size"))
agg = OrderedDict()
tot = 0.0
for d in data:
    if not (lo <= int(d[iid]) <= hi):
        continue
    name = d[ik].split("(")[0].replace("void ", "")[:70]
    ns = float(d[iv].replace(",", ""))
    e = agg.setdefault(name, [0, 0.0, d[ig], d[ib]])
    e[0] += 1
    e[1] += ns
    tot += ns
print(f"| kernel | launches | total us | share | grid | block |\n|---|---|---|---|---|---|")
for name, (n, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot:.1f}% | {g} | {b} |")
print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot / 1e3:.1f} | 100% | | |")
