# GPU tests, a short bench, and the per-launch device-time list (ncu) of the same bench command
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/t5.log; cat gpurun_out/t5.log
CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none ${NCU_CACHE:-} -s 300 -c 320 --csv --log-file gpurun_out/launches${NCU_TAG:-}.csv $CMD > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200
