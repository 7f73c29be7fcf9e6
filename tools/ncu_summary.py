"""ncu_summary.py report.ncu-rep out.md [title] - key metrics + top stall lines of the first kernel in a report."""
import csv, subprocess, sys, io

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__inst_executed.sum.per_cycle_elapsed"]
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none; per-launch, cold cache, serialised)", "",
         "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        lines.append(f"| {k} | {vals[i]} | {units[i]} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[2:] if len(r) == len(h)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
lines += ["", f"## Top stall sites ({tot} warp samples)", "", "| samples | % | executed | SASS | dominant stalls |", "|---|---|---|---|---|"]
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
    st = sorted(((int(r[ix[n]] or 0), n) for n in h[30:47]), reverse=True)[:2]
    s = int(r[ix["# Samples"]] or 0)
    lines.append(f"| {s} | {100.0 * s / max(tot, 1):.1f} | {r[ix['Instructions Executed']]} | `{r[1].strip()[:60]}` | "
                 + ", ".join(f"{n}={c}" for c, n in st if c) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:30]))
