mkdir -p gpurun_out
timeout 120 python scripts/ncu_layers.py 16 > gpurun_out/r1g_layers_times.txt 2>&1; cat gpurun_out/r1g_layers_times.txt | tail -12
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r1g_layers python scripts/ncu_layers.py 16 > gpurun_out/r1g_ncu_layers.log 2>&1
tail -3 gpurun_out/r1g_ncu_layers.log; ls -la gpurun_out/*.ncu-rep
