#!/bin/bash
mkdir -p gpurun_out
python scripts/layer_bench.py wgrad > gpurun_out/r2l_layer_bench.txt 2>&1
cat gpurun_out/r2l_layer_bench.txt
for pat in "G.dec4" "D.d5" "G.enc1"; do
  tag=$(echo $pat | tr -d '. ')
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_wgrad \
    -o gpurun_out/r2l_wgrad_$tag -f python scripts/layer_bench.py wgrad "$pat" > gpurun_out/r2l_ncu_$tag.log 2>&1
  tail -1 gpurun_out/r2l_ncu_$tag.log
done
