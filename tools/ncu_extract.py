"""Extract the judged metrics of an .ncu-rep into a small text table.  python tools/ncu_extract.py REPORT.ncu-rep > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem"]
print(f"# {rep}: ncu --set full --clock-control none (per-launch, cold cache, serialised)")
for k, r in enumerate(data):
    print(f"\n## launch {k}")
    for w in want:
        if w in idx:
            v = r[idx[w]]
            if w == "Kernel Name":
                v = v[:110]
            print(f"{w:70s} {v} {units[idx[w]]}")
