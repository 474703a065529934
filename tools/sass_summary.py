"""Per-kernel count of the Blackwell tensor-core / TMA SASS mnemonics in libfdm_sm100.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), UBLKCP (bulk copy), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier); 2CTA = instructions in their cta_group::2 form.
python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200", "libfdm_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
counts, cur, k = collections.OrderedDict(), None, -1
MN = ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "HMMA", "FFMA", "2CTA")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        cur = names[k] if k < len(names) else m.group(1)
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for mn in MN:
            if op.startswith(mn):
                counts[cur][mn] += 1
        counts[cur]["total"] += 1
        if ".2CTA" in line:  # cta_group::2 forms (UTCHMMA.2CTA, UTMALDG.*.2CTA, UTCBAR.2CTA.MULTICAST)
            counts[cur]["2CTA"] += 1
print(f"# cuobjdump -sass {os.path.basename(lib)} (sm_100a): instruction counts per kernel")
print(f"{'kernel':78s} " + " ".join(f"{m:>8s}" for m in MN) + "    total")
for name, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTCHMMA"]):
    if c["total"] == 0:
        continue
    print(f"{name[:78]:78s} " + " ".join(f"{c[m]:8d}" for m in MN) + f" {c['total']:8d}")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'ALL KERNELS':78s} " + " ".join(f"{tot[m]:8d}" for m in MN) + f" {tot['total']:8d}")
