#!/bin/bash
# same-box A/B of the epilogue-written reflection halo
for rep in 1 2; do
for v in 0 1; do
UEGAN_NO_EPILOGUE_HALO=$v python bench.py --workload inference --steps 30 --warmup 5 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('infer no_epi_halo=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
done
for v in 0 1 0 1; do
UEGAN_NO_EPILOGUE_HALO=$v python bench.py --steps 20 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train no_epi_halo=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
