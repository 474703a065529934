# ncu --set full capture of the dominant kernels of one denoiser step (eager warm-up pass of bench.py):
# the first launches of conv_halo_kernel (top-level 64->64 and level-1 128->128 convs), conv_tc_kernel, and both attention kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel|attn_temporal_kernel|attn_spatial_tc_kernel|gn_apply_kernel" -c 14 -f -o gpurun_out/r01_top_kernels $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
