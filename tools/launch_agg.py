"""Aggregate an ncu gpu__time_duration.sum CSV launch list by kernel name: python tools/launch_agg.py file.csv [top-N single launches]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
data = [(re.sub(r"\(.*", "", row[ki]).replace("void ", "").replace("fdm::", ""), row[gi], float(row[vi].replace(",", ""))) for row in r]
tot = sum(d[2] for d in data)
print(f"{len(data)} launches, sum {tot / 1e6:.3f} ms (cold-cache, serialised)")
agg = collections.defaultdict(lambda: [0, 0.0])
for n, g, ns in data:
    agg[n][0] += 1
    agg[n][1] += ns
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {k[:110]}")
for n, g, ns in sorted(data, key=lambda d: -d[2])[:int(sys.argv[2]) if len(sys.argv) > 2 else 0]:
    print(f"   {ns / 1e3:9.1f} us  {g:20s} {n[:90]}")
