"""SASS evidence: per-kernel counts of the mnemonics that prove tcgen05 / TMEM / TMA / DSMEM / packed-fp32 use
(cuobjdump -sass of the built libpcst.so; no GPU needed).   python tools/sass_table.py > profiles/r02/sass_mnemonics.md"""
import os
import re
import subprocess

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pointcloud_style_transfer_b200", "csrc", "libpcst.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
pats = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "STAS", "CREDUX", "UCGABAR", "FFMA2", "FADD2", "FMUL2",
        "FMNMX3", "DMUL", "DADD", "BAR.SYNC", "ATOMS", "ATOMG", "RED.E"]
print("# SASS evidence per kernel (cuobjdump -sass libpcst.so, sm_100a): mnemonic counts\n")
print("`UTCHMMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit -> mbarrier, `LDTM` = tcgen05.ld (TMEM -> registers), `UTCATOMSWS` = TMEM "
      "alloc / dealloc, `UBLKCP` = cp.async.bulk (1-D TMA), `SYNCS` = mbarrier operations, `STAS` = st.async into a peer CTA's shared "
      "memory (DSMEM), `CREDUX` = warp-wide integer reduction, `UCGABAR` = barrier.cluster, `FFMA2 / FADD2 / FMUL2` = packed fp32x2 "
      "arithmetic, `FMNMX3` = 3-input min / max, `DMUL / DADD` = non-fused fp64 (the exact kNN ranking).\n")
print("| kernel | SASS instructions | " + " | ".join(pats) + " |")
print("|---|---|" + "---|" * len(pats))
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    n = len(re.findall(r"/\*[0-9a-f]{4,6}\*/\s+[@A-Z]", f))
    if n < 40:
        continue
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    d = d.replace("void ", "").replace("pcst::", "")
    cnt = [len(re.findall(r"\b" + re.escape(p), f)) for p in pats]
    rows.append((d, n, cnt))
for d, n, cnt in sorted(rows):
    print(f"| `{d[:64]}` | {n} | " + " | ".join(str(c) if c else "" for c in cnt) + " |")
