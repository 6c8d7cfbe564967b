#!/bin/bash
# r4i: same-box check that the inference path did not regress: old .so (before the launch_pdl plumbing) vs current, epilogue halo on/off
cp uegan_b200/libuegan_sm100.so /tmp/new.so
for rep in 1 2; do
for v in old new; do
if [ $v = old ]; then cp scratch_ab/old.so uegan_b200/libuegan_sm100.so; else cp /tmp/new.so uegan_b200/libuegan_sm100.so; fi
for h in 0 1; do
UEGAN_NO_EPILOGUE_HALO=$h timeout 300 python bench.py --workload inference --steps 30 --warmup 5 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('infer $v no_epi_halo=$h', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
done
done
cp /tmp/new.so uegan_b200/libuegan_sm100.so
