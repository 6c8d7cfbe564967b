#!/bin/bash
# ncu --set full of representative launches of the final round-2 build (one launch each, after warm-up); the reports are
# reduced on the box to their raw-metric and source-page CSVs (the .ncu-rep files exceed the 64 MiB return limit)
mkdir -p gpurun_out
for spec in "fprop:G.dec5.0:conv_fprop" "fprop:VGG conv1_2:conv_fprop" "wgrad:G.dec4:conv_wgrad" "wgrad:G.dec1:conv_wgrad" "wgrad:G.enc1:conv_wgrad"; do
  what=${spec%%:*}; rest=${spec#*:}; pat=${rest%%:*}; kr=${rest#*:}
  tag=$(echo "${what}_${pat}" | tr -d '. ' )
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$kr \
    -o /tmp/r3n_$tag -f python scripts/layer_bench.py $what "$pat" > gpurun_out/r3n_ncu_$tag.log 2>&1
  ncu -i /tmp/r3n_$tag.ncu-rep --page raw --csv > gpurun_out/r3n_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/r3n_$tag.ncu-rep --page source --csv > gpurun_out/r3n_${tag}_source.csv 2>/dev/null
done
ls -la gpurun_out/r3n_*
