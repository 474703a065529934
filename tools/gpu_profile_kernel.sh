# ncu --set full of the first launches of ONE kernel (regex in $K, count in $N) of a denoiser step: K=attn_temporal N=2 bash tools/gpu_profile_kernel.sh
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"${K}" -c ${N:-2} -f -o gpurun_out/r01_${TAG:-kernel} $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
