mkdir -p gpurun_out
for m in all none 64 32; do
  UEGAN_FUSED_STATS=$m timeout 200 python bench.py --steps 6 --warmup 3 > gpurun_out/r1h_train_fs$m.json 2> gpurun_out/r1h_train_fs$m.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1h_train_fs$m.json")); r=d["roofline"]
    print("FUSED_STATS=$m", round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms", r["by_kind_ms_tflops"])
except Exception as e:
    print("FUSED_STATS=$m failed", e); print(open("gpurun_out/r1h_train_fs$m.err").read()[-600:])
PY
done
for m in all none; do
UEGAN_FUSED_STATS=$m timeout 200 python bench.py --workload inference --steps 20 --warmup 5 > gpurun_out/r1h_infer_fs$m.json 2>/dev/null; head -c 160 gpurun_out/r1h_infer_fs$m.json; echo
done
