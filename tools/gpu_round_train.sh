# Round artefacts of the training path: ncu launch lists of one native training step (cfg2 / cfg3), ncu --set full of the
# tcgen05 wgrad / attention-backward kernels and the GroupNorm backward (a few launches each: gpurun merges back <= 64 MiB),
# all AFTER the plain command exited 0.
mkdir -p gpurun_out
python tools/train_launches.py cfg3 > gpurun_out/plain_train.log 2>&1 || exit 1
for c in cfg2 cfg3; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_${c}_train.csv python tools/train_launches.py $c > gpurun_out/ncu_train.log 2>&1
done
ncu --profile-from-start off --set full --clock-control none -k regex:"wgrad_tc_kernel" -s 20 -c 6 -f -o gpurun_out/r01_wgrad_tc python tools/train_launches.py cfg3 > gpurun_out/ncu_full_train.log 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:"attn_spatial_bwd_tc|attn_temporal_bwd" -c 8 -f -o gpurun_out/r01_attn_bwd python tools/train_launches.py cfg3 >> gpurun_out/ncu_full_train.log 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:"gn_bwd_kernel" -c 4 -f -o gpurun_out/r01_gn_bwd python tools/train_launches.py cfg3 >> gpurun_out/ncu_full_train.log 2>&1
tail -2 gpurun_out/ncu_full_train.log; ls -la gpurun_out/*.ncu-rep
