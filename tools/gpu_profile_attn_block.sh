# ncu --set full of the first launches of the attention-block kernels of a cfg4 step (the 16x16 level comes first in launch order):
# norm+qkv linear, the three temporal-attention kernels, proj_out (per-tap kernel), spatial attention.  Extracts only (the report stays on the box).
mkdir -p gpurun_out
CMD="python bench.py --workload cfg4-sampling --steps 3 --warmup 3 --no-train --no-e2e --no-cpu-baseline --no-gpu-eager"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nl_qkv_kernel|attn_rows_kernel|rpe_bias_kernel|rpe_pv_kernel|attn_spatial_tc_kernel|conv_tc_kernel" -c 12 -f -o gpurun_out/r02_attn_block $CMD > gpurun_out/ncu_attn_block.log 2>&1
tail -2 gpurun_out/ncu_attn_block.log
python tools/ncu_extract.py gpurun_out/r02_attn_block.ncu-rep > gpurun_out/r02_attn_block_ncu_full.txt 2>&1
for k in nl_qkv_kernel attn_rows_kernel attn_spatial_tc_kernel conv_tc_kernel rpe_pv_kernel; do
  python tools/ncu_stalls.py gpurun_out/r02_attn_block.ncu-rep $k 14 > gpurun_out/r02_attn_block_stalls_$k.txt 2>&1
done
rm -f gpurun_out/*.ncu-rep
