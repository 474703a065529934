"""Where the gap between the device-resident training step (bench.py `train`) and NativeTrainStep with host batches comes from:
the same runner at cfg3 with (a) host videos, (b) videos already on the device, (c) the timed loop under torch's profiler
(top GPU kernels outside the two CUDA graphs).  Usage: python tools/train_step_gap.py"""
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
for env in ({}, {"FDM_TSB_DEVICE_POOL": "1"}):
    r = subprocess.run([sys.executable, os.path.join(here, "train_step_bench.py"), "cfg3", "30"],
                       env=dict(os.environ, FDM_TSB_ARMS="native", **env), capture_output=True, text=True)
    print(env, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:])
