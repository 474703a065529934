# A/B of the spatial attention kernel: FDM_SA_SPLIT_MIN_L (two threads per score row for L >= this; 0 = never) x FDM_SA_VBAR (V on its own barrier)
mkdir -p gpurun_out
OPT="--steps 30 --warmup 5 --no-train --no-e2e --no-cpu-baseline --no-gpu-eager"
run() { FDM_SA_SPLIT_MIN_L=$1 FDM_SA_VBAR=$2 timeout 200 python bench.py --workload $3 $OPT 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split_min_l=$1 vbar=$2','$3',d['ms_per_step'],d['ms_per_step_hot'],d['parity']['eps_rel_l2_vs_cpu_oracle'],d['parity']['ok'])"; }
run 0 0 cfg4-sampling; run 128 0 cfg4-sampling; run 128 1 cfg4-sampling; run 0 0 cfg4-sampling; run 128 0 cfg4-sampling
run 128 0 cfg5-sampling
FDM_SA_SPLIT_MIN_L=128 FDM_SA_VBAR=0 timeout 200 python tools/step_profile.py cfg4-sampling 2>/dev/null | grep -E "attn_spatial|^#"
FDM_SA_SPLIT_MIN_L=128 FDM_SA_VBAR=0 timeout 100 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attn_spatial_vs_torch" 2>&1 | tail -1
