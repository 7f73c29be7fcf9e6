"""launch_table.py launches.csv [skip_prefixes...] - markdown table (kernel, launches, average us) from an
`ncu --metrics gpu__time_duration.sum --csv` launch list; torch's own elementwise kernels are dropped."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"]
        if name.startswith("void at::") or name.startswith("at::"):
            continue
        key = (name.split("(")[0][:90], d.get("Grid Size", ""), d.get("Block Size", ""))
        v = float(d["Metric Value"]) / (1e3 if d.get("Metric Unit") == "ns" else 1.0)
        agg.setdefault(key, []).append(v)
print("| kernel | grid | block | launches | average us |")
print("|---|---|---|---|---|")
for (name, grid, block), v in agg.items():
    print(f"| `{name}` | {grid} | {block} | {len(v)} | {sum(v) / len(v):.1f} |")
