# ncu --set full of every conv_halo_kernel launch of one denoiser step (eager warm-up pass of bench.py) -> traffic per launch
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel" -c 15 -f -o gpurun_out/r01_conv_halo $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
