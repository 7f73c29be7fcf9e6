"""ncu_multi_summary.py report.ncu-rep - one markdown table row per kernel of a `ncu --set full` report."""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = [("gpu__time_duration.sum", "time"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("smsp__inst_executed.sum", "warp instructions")]
print("| kernel | " + " | ".join(k[1] for k in keys) + " |")
print("|---|" + "---|" * len(keys))
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0][:70]
    cells = []
    for k, _ in keys:
        if k in hdr:
            i = hdr.index(k)
            v = r[i]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {units[i]}".strip())
        else:
            cells.append("-")
    print(f"| `{name}` | " + " | ".join(cells) + " |")
