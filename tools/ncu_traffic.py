"""profiles/r01_halo_traffic.json from an `ncu --set full` report: mean DRAM bytes (read + write) per conv_halo_kernel launch.
python tools/ncu_traffic.py REPORT.ncu-rep WORKLOAD > profiles/r01_halo_traffic.json"""
import csv
import io
import json
import subprocess
import sys

rep, workload = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n, dur = 0.0, 0, 0.0
for r in data:
    if "conv_halo_kernel" not in r[idx["Kernel Name"]]:
        continue
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[idx[key]].replace(",", "")) * scale[units[idx[key]]]
    dur += float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
    n += 1
print(json.dumps({"workload": workload, "kernel": "conv_halo_kernel", "launches": n, "dram_bytes_per_launch": tot / max(n, 1),
                  "mean_duration_us_under_ncu": dur / max(n, 1),
                  "note": "ncu --set full --clock-control none, cold cache per launch (ncu flushes L2 between kernels); writes mostly stay in the 126 MB L2"}))
