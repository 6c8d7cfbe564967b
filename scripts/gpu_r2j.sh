#!/bin/bash
# round-2 pass J: per-layer wgrad / stride-2 dgrad baseline (CUDA events) + ncu --set full of three wgrad launches
mkdir -p gpurun_out
python scripts/layer_bench.py all > gpurun_out/r2j_layer_bench.txt 2>&1
cat gpurun_out/r2j_layer_bench.txt
for pat in "G.dec1" "D.d2" "G.enc1"; do
  tag=$(echo $pat | tr -d '. ')
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_wgrad \
    -o gpurun_out/r2j_wgrad_$tag -f python scripts/layer_bench.py wgrad "$pat" > gpurun_out/r2j_ncu_$tag.log 2>&1
  tail -2 gpurun_out/r2j_ncu_$tag.log
done
ls -la gpurun_out/*.ncu-rep
