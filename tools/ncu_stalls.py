"""Top SASS lines by warp-stall samples of one kernel in an .ncu-rep (needs --import-source on / --set full).
python tools/ncu_stalls.py REPORT.ncu-rep KERNEL_REGEX [N]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several launches may be concatenated: each starts with a "Kernel Name" row followed by a header row
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks[:1]:
    h = b["hdr"]
    si, src = h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    data, tot = [], 0.0
    for k, r in enumerate(b["rows"]):
        try:
            v = float(r[si])
        except (ValueError, IndexError):
            continue
        tot += v
        top = max(stall_cols, key=lambda i: float(r[i] or 0))
        data.append((v, k, r[src].strip()[:110], h[top]))
    print(f"# {b['name'][:100]}: {int(tot)} samples")
    for v, k, s, why in sorted(data, reverse=True)[:n]:
        print(f"{100 * v / tot:5.1f}%  #{k:4d}  {why:22s} {s}")
