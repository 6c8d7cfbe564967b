#!/bin/bash
# 8-GPU pass (configs[3]): training bench, global batch 128, CUDA graph with the peer-memory reductions
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r3v_topo.txt 2>&1
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --lib-baseline 0 > gpurun_out/r3v_bench_train_n$N.json 2> gpurun_out/r3v_bench_train_n$N.err
echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3v_bench_train_n$N.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["config"]["cuda_graph"], d["config"]["host_enqueue_ms_per_step"], d.get("replicas"), d["clocks"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r3v_bench_train_n$N.err").read()[-3000:])
PY
grep -iE "warn|peer|nccl" gpurun_out/r3v_bench_train_n$N.err | head -5
