#!/bin/sh
# BASELINE.json configs[4]: long-context sampling (nc=128, 4x64x64 latents, K=40), batch sweep B in {1,2,4,8,16} videos per GPU.
#   N=1:   sh tools/cfg5_sweep.sh 1 > profiles/r02_cfg5_sweep_1gpu.jsonl
#   N>1:   sh tools/cfg5_sweep.sh 8   (launches torchrun per batch size; one process per GPU, batch sharded, no collective)
N=${1:-1}
for B in 1 2 4 8 16; do
  ARGS="--workload cfg5-sampling --batch $B --gpus $N --steps 8 --warmup 3 --no-train --no-e2e --no-cpu-baseline --no-gpu-eager"
  if [ "$N" = "1" ]; then
    python bench.py $ARGS | tail -1
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py $ARGS | tail -1
  fi
done
