#!/bin/bash
mkdir -p gpurun_out
for pat in "G.dec5.0" "G.enc1"; do
  tag=$(echo $pat | tr -d '. ')
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_fprop \
    -o gpurun_out/r2t_fprop_$tag -f python scripts/layer_bench.py fprop "$pat" > gpurun_out/r2t_ncu_$tag.log 2>&1
  tail -1 gpurun_out/r2t_ncu_$tag.log
done
