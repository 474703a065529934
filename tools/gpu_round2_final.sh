# Round-2 closing measurements on one B200 (every artefact lands in gpurun_out/, copied to profiles/r02_* afterwards):
#   sh tools/gpu_round2_final.sh A   tests, bench records (default / reference arm / cfg5 / cfg3), step profiles, cfg5 batch sweep
#   sh tools/gpu_round2_final.sh B   ncu launch list (multi-metric) of the cfg4 step, ncu --set full of the halo / head kernels
mkdir -p gpurun_out
if [ "$1" = "A" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
  timeout 600 python bench.py > gpurun_out/final_bench_default.log 2>&1; tail -1 gpurun_out/final_bench_default.log > gpurun_out/final_bench_default.json
  timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.log 2>&1; tail -1 gpurun_out/final_bench_reference.log > gpurun_out/final_bench_reference.json
  for wl in cfg5-sampling cfg3-sampling; do
    timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-train --no-cpu-baseline > gpurun_out/final_bench_$wl.log 2>&1
    tail -1 gpurun_out/final_bench_$wl.log > gpurun_out/final_bench_$wl.json
  done
  for wl in cfg4-sampling cfg5-sampling cfg3-sampling; do timeout 300 python tools/step_profile.py $wl > gpurun_out/final_step_profile_$wl.txt 2>&1; done
  timeout 900 sh tools/cfg5_sweep.sh 1 > gpurun_out/final_cfg5_sweep_1gpu.jsonl 2> gpurun_out/final_cfg5_sweep.err
  for f in gpurun_out/final_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("parity") or {}).get("ok"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
  done
else
  CMD="python bench.py --workload cfg4-sampling --steps 3 --warmup 3 --no-train --no-e2e --no-cpu-baseline --no-gpu-eager"
  $CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -s 300 -c 340 --csv --log-file gpurun_out/final_launches_long.csv $CMD > gpurun_out/ncu_list.log 2>&1
  tail -2 gpurun_out/ncu_list.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel|conv_head_kernel" -c 16 -f -o gpurun_out/r02_conv_halo_head $CMD > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
  CMD5="python bench.py --workload cfg5-sampling --steps 3 --warmup 3 --no-train --no-e2e --no-cpu-baseline --no-gpu-eager"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel" -c 24 -f -o gpurun_out/r02_conv_halo_cfg5 $CMD5 > gpurun_out/ncu_full5.log 2>&1
  tail -2 gpurun_out/ncu_full5.log
  # the reports are too large to travel back (64 MiB limit): extract the judged tables here and keep the text only
  for r in r02_conv_halo_head r02_conv_halo_cfg5; do
    python tools/ncu_extract.py gpurun_out/$r.ncu-rep > gpurun_out/${r}_ncu_full.txt 2>&1
    python tools/ncu_stalls.py gpurun_out/$r.ncu-rep conv_halo_kernel 16 > gpurun_out/${r}_ncu_stalls.txt 2>&1
  done
  python tools/ncu_traffic.py gpurun_out/r02_conv_halo_head.ncu-rep cfg4-sampling > gpurun_out/final_halo_traffic.json 2>&1
  rm -f gpurun_out/*.ncu-rep
  du -sh gpurun_out
fi
