"""Sums dram__bytes_read.sum + dram__bytes_write.sum of the tcgen05 GEMM launches (conv_fprop*, conv_rowsum*, conv_wgrad*) of
ONE step from an ncu per-launch CSV (scripts/gpu_r2g.sh) and records it in profiles/r2_step_traffic.json, which bench.py reads
for `roofline.traffic`.    python scripts/ncu_traffic.py <launches.csv> <train|inference> <f16|tf32> <source note>"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, workload, prec, note = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
rows = collections.OrderedDict()
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    e = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    try:
        e[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        pass
gemm = [e for e in rows.values() if any(k in e["name"] for k in ("conv_fprop_kernel", "conv_rowsum_kernel", "conv_wgrad"))]
tot = lambda es, k: sum(e.get(k, 0.0) for e in es)
out = {
    "gemm_launches": len(gemm), "all_launches": len(rows),
    "gemm_dram_bytes_per_step": tot(gemm, "dram__bytes_read.sum") + tot(gemm, "dram__bytes_write.sum"),
    "all_dram_bytes_per_step": tot(rows.values(), "dram__bytes_read.sum") + tot(rows.values(), "dram__bytes_write.sum"),
    "gemm_ms_serialised": tot(gemm, "gpu__time_duration.sum") / 1e6, "all_ms_serialised": tot(rows.values(), "gpu__time_duration.sum") / 1e6,
    "source": note,
}
dst = os.path.join(ROOT, "profiles", "r2_step_traffic.json")
d = json.load(open(dst)) if os.path.exists(dst) else {}
d[f"{workload}_{prec}"] = out
json.dump(d, open(dst, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
