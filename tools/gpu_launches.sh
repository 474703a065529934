# per-launch device-time list (ncu) of one workload: WL=<workload> TAG=<suffix> bash tools/gpu_launches.sh
mkdir -p gpurun_out
CMD="python bench.py --workload ${WL:-cfg4-sampling} --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 340 --csv --log-file gpurun_out/launches_${TAG:-x}.csv $CMD > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200
