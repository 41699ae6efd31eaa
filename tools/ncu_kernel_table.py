"""Per-kernel summary table of an `ncu --set full` report (needs `ncu` on PATH; no GPU).

    python tools/ncu_kernel_table.py gpurun_out/prof_full.ncu-rep
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[0], rows[2:]
cols = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe busy %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe busy %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
]
have = [(hdr.index(c), n) for c, n in cols if c in hdr]
units = rows[1]
print("| " + " | ".join(n for _, n in have) + " |")
print("|" + "---|" * len(have))
for r in data:
    cells = []
    for i, n in have:
        v = r[i]
        if n == "kernel":
            v = "`" + v.split("(")[0].replace("void ", "").replace("pcst::", "")[:44] + "`"
        elif n in ("us",):
            v = f"{float(v.replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}[units[i]]:.1f}"
        elif "MB" in n:
            f = float(v.replace(",", ""))
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[i], 1e-6)
            v = f"{f * scale:.2f}"
        elif "%" in n:
            v = f"{float(v.replace(',', '')):.1f}"
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
