#!/bin/bash
# r4f: programmatic dependent launch on every kernel of the library: full parity suite, then same-box A/B (UEGAN_PDL=0|1)
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do
for v in 0 1; do
UEGAN_PDL=$v timeout 300 python bench.py --steps 20 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train pdl=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
done
for v in 0 1; do
UEGAN_PDL=$v timeout 300 python bench.py --workload inference --steps 30 --warmup 5 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('infer pdl=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
