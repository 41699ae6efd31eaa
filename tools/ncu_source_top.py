"""Top stalled SASS instructions of one kernel in an .ncu-rep (needs `ncu` on PATH; no GPU).

    python tools/ncu_source_top.py <report.ncu-rep> <kernel-regex> [launch-skip] [top-n]
"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the page is: a "Kernel Name" row, a header row, then one row per SASS instruction (possibly repeated per kernel)
hdr = None
data = []
seen = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        seen += 1
        if seen > 1:
            break  # the page repeats the listing per view; keep the first
        print("kernel:", r[1][:140])
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
s0, s1 = hdr.index("stall_barrier"), hdr.index("stall_wait")
tot = sum(int(r[i_s] or 0) for r in data)
print("total samples", tot, "instructions", len(data), "warp-instr executed", sum(int(r[i_ex] or 0) for r in data))
agg = {}
for r in data:
    for j in range(s0, s1 + 1):
        agg[hdr[j]] = agg.get(hdr[j], 0) + int(r[j] or 0)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for n, r in enumerate(data):
    r.append(n)
for r in sorted(data, key=lambda r: -int(r[i_s] or 0))[:topn]:
    st = {hdr[j][6:]: int(r[j]) for j in range(s0, s1 + 1) if r[j] not in ("0", "")}
    print(f"{int(r[i_s]):6d} {100.0 * int(r[i_s]) / max(tot, 1):5.1f}%  #{r[-1]:<5d} exec {r[i_ex]:>8s}  {r[i_src][:70]:70s} {st}")
