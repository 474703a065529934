"""One denoiser step from an ncu multi-metric launch list (long CSV: one row per launch and metric) -> wide CSV + per-kernel summary.
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file L.csv python bench.py ...
python tools/launch_table.py L.csv OUT.csv > OUT_summary.txt     (the step = the launches between two timestep_embedding kernels)"""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
I = {h: i for i, h in enumerate(hdr)}
scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}
L = collections.OrderedDict()
for r in rows:
    d = L.setdefault(r[I["ID"]], {"name": r[I["Kernel Name"]], "grid": r[I["Grid Size"]]})
    d[r[I["Metric Name"]]] = float(r[I["Metric Value"]].replace(",", "")) * scale.get(r[I["Metric Unit"]], 1.0)
ids = list(L)
marks = [i for i, k in enumerate(ids) if "timestep_embedding" in L[k]["name"]]
step = ids[marks[-2]:marks[-1]] if len(marks) >= 2 else ids
T, RD, WR = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
TP = next((k for k in L[ids[0]] if k.startswith("sm__pipe_tensor")), None)
with open(sys.argv[2], "w") as f:
    w = csv.writer(f)
    w.writerow(["ID", "Kernel Name", "Grid Size", T + " [ns]", RD + " [B]", WR + " [B]", "sm__pipe_tensor_cycles_active [% of peak, active]"])
    for k in step:
        d = L[k]
        w.writerow([k, d["name"], d["grid"], d.get(T, 0), d.get(RD, 0), d.get(WR, 0), d.get(TP, 0) if TP else ""])
tot = sum(L[k].get(T, 0) for k in step)
print(f"# one denoiser forward schedule + posterior update: {len(step)} launches, sum {tot / 1e3:.1f} us")
print("# per-launch times are cold-cache and serialised (ncu flushes caches between launches): compare SHARES with bench.py's roofline_table")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for k in step:
    d = L[k]
    n = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("fdm::", "")
    n = re.sub(r"<.*", "", n) if "at::" in n else n
    a = agg[n]
    a[0] += 1; a[1] += d.get(T, 0); a[2] += d.get(RD, 0); a[3] += d.get(WR, 0); a[4] = max(a[4], d.get(TP, 0) if TP else 0)
print(f"{'kernel':56s}{'n':>4s}{'us':>11s}{'share':>8s}{'DRAM rd MB':>12s}{'DRAM wr MB':>12s}{'tensor%max':>11s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:55]:56s}{a[0]:4d}{a[1] / 1e3:11.1f}{100 * a[1] / tot:7.1f}%{a[2] / 1e6:12.1f}{a[3] / 1e6:12.1f}{a[4]:11.1f}")
