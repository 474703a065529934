# ncu --set full capture of the dominant kernels of one denoiser step (eager warm-up pass of bench.py):
# every conv_halo_kernel launch of the step plus the first temporal/spatial attention and GroupNorm-apply launches
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel" -c 15 -f -o gpurun_out/r01_conv_halo $CMD > gpurun_out/ncu_full.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"attn_temporal_kernel|attn_spatial_tc_kernel|gn_apply_kernel|conv_tc_kernel|temporal_gn_kernel" -c 24 -f -o gpurun_out/r01_other_kernels $CMD >> gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
