#!/bin/bash
# 2-GPU pass (round-2 final build): data-parallel equivalence over peer memory, then the training bench (CUDA graph, 2 streams) at N = 2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/ddp_equivalence.py > gpurun_out/r3f_ddp_equivalence_peer.log 2>&1
echo "peer exit $?"; grep -E "reductions via|rel |post-step|checksums|DDP_EQUIVALENCE|Error|error" gpurun_out/r3f_ddp_equivalence_peer.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --lib-baseline 0 > gpurun_out/r3f_bench_train_n2.json 2> gpurun_out/r3f_bench_train_n2.err
echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r3f_bench_train_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["config"]["cuda_graph"], d["config"]["host_enqueue_ms_per_step"], d.get("replicas"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r3f_bench_train_n2.err").read()[-3000:])
PY
