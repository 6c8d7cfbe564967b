#!/bin/bash
mkdir -p gpurun_out
python scripts/layer_bench.py fprop 2>&1 | tee gpurun_out/r2r_fprop_bench.txt
echo "== no patch16/64"; UEGAN_NO_PATCH16=1 UEGAN_NO_PATCH64=1 python scripts/layer_bench.py fprop 2>&1 | tee gpurun_out/r2r_fprop_bench_old.txt
for pat in "G.enc1" "G.dec5.0"; do
  tag=$(echo $pat | tr -d '. ')
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_fprop \
    -o gpurun_out/r2r_fprop_$tag -f python scripts/layer_bench.py fprop "$pat" > gpurun_out/r2r_ncu_$tag.log 2>&1
  tail -1 gpurun_out/r2r_ncu_$tag.log
done
