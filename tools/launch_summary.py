"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one diffusion step, time per kernel class."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
data = [(row[ki], row[gi], float(row[vi].replace(",", ""))) for row in r]
idx = [i for i, d in enumerate(data) if "timestep_embedding" in d[0]]
a, b = idx[0], idx[1]
step = data[a:b]
tot = sum(d[2] for d in step)
print(f"launches in one step: {len(step)}   sum of kernel durations: {tot / 1000:.1f} us (cold-cache, serialised: compare shares)")
agg = collections.defaultdict(lambda: [0, 0.0])
for name, grid, ns in step:
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    short = re.sub(r"<.*", "", short) if "at::" in short else short
    agg[short][0] += 1
    agg[short][1] += ns
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1000:9.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:3d}  {k[:100]}")
if len(sys.argv) > 2:
    for name, grid, ns in step:
        print(f"{ns/1000:8.1f} us  {grid:18s} {re.sub(r'[(].*', '', name)[:90]}")
