#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/ddp_equivalence.py 2>&1 | grep -E "reductions via|rel |post-step|checksums|DDP_EQUIVALENCE|mismatch"
