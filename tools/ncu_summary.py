"""Summarise an `ncu --set full` report (one row per captured launch) as a markdown table.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM write MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
print("| kernel | " + " | ".join(n for _, n in COLS) + " |")
print("|---|" + "---|" * len(COLS))
units = rows[1]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[idx["Kernel Name"]]
    name = name.split("(")[0].replace("segb200::", "").replace("void ", "")
    cells = []
    for key, _ in COLS:
        v = r[idx[key]] if key in idx else ""
        try:
            f = float(v.replace(",", ""))
            u = units[idx[key]] if key in idx else ""
            if key.startswith("dram__bytes"):
                f = f * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
            if key == "gpu__time_duration.sum":
                f = f * {"us": 1.0, "ns": 1e-3, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
            cells.append(f"{f:.1f}" if abs(f) < 1e6 and f != int(f) else f"{f:.0f}")
        except ValueError:
            cells.append(v)
    print(f"| `{name}` | " + " | ".join(cells) + " |")
