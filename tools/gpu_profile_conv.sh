# ncu --set full capture of the first conv_tc launches of one denoiser step (eager warm-up pass of bench.py)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 8 -f -o gpurun_out/conv_tc $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
