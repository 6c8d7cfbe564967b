#!/bin/bash
# same 2-GPU box: training bench at N = 1, then N = 2 (scaling on identical hardware), then the data-parallel equivalence
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_ddp_host.py -m gpu -q -x -p no:cacheprovider -k "batched_generator or two_gpus" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --lib-baseline 0 > gpurun_out/r4y_bench_train_n1.json 2> gpurun_out/r4y_bench_train_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 --lib-baseline 0 > gpurun_out/r4y_bench_train_n2.json 2> gpurun_out/r4y_bench_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/ddp_equivalence.py > gpurun_out/r4y_ddp_equivalence_peer.log 2>&1
grep -E "reductions via|rel |post-step|checksums|DDP_EQUIVALENCE|weight_u" gpurun_out/r4y_ddp_equivalence_peer.log | tail -14
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.load(open(f"gpurun_out/r4y_bench_train_n{n}.json"))
        print(n, {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["config"]["cuda_graph"], d.get("replicas"), d["clocks"])
    except Exception as e:
        print(n, "parse failed", e); print(open(f"gpurun_out/r4y_bench_train_n{n}.err").read()[-2000:])
PY
