#!/bin/sh
# ORACLE — TEST INFRASTRUCTURE.  Recipe that ships the UNMODIFIED reference to the GPU box.
#
# The reference (plai-group/latent-flexible-video-diffusion-modeling) is a pure-Python tree: nothing to compile.  This
# script copies its package and scripts from where they lie (/root/reference, read-only, build container only) into
# oracle/_ref/ — git-ignored (never part of the history), NOT gpurun-ignored (travels to the B200 box like our built .so).
# Run by __graft_entry__.build(); a no-op where /root/reference does not exist (the GPU box uses the shipped copy).
#
# Consumers (checker / baseline only, never the product): oracle/ref_loader.py -> tests/, bench.py --impl reference,
# bench.py's cpu_baseline and gpu_eager legs, __graft_entry__.smoke().
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${FDM_REFERENCE:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/improved_diffusion" ]; then
  echo "make_ref: no reference checkout at $SRC (GPU box?) - keeping $DST as shipped"
  exit 0
fi
mkdir -p "$DST"
for d in improved_diffusion scripts; do
  rm -rf "$DST/$d"
  cp -r "$SRC/$d" "$DST/$d"
done
cp "$SRC/LICENSE" "$DST/LICENSE" 2>/dev/null || true
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} + 2>/dev/null || true
( cd "$SRC" && find improved_diffusion scripts -name '*.py' | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "make_ref: copied $(wc -l < "$DST/SHA256SUMS") reference files into $DST"
