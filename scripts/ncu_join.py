"""Joins an ncu per-launch CSV (any --metrics list) of scripts/ncu_step.py with the wrapper trace it wrote
(UEGAN_TRACE_OUT): every uegan:: kernel launch gets the layer / tensor description of the call that issued it.
    python scripts/ncu_join.py launches.csv trace.json > table.md"""
import csv, json, sys, collections

rows = collections.OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    e = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
    try:
        e[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        e[r["Metric Name"]] = r["Metric Value"]
trace = json.load(open(sys.argv[2]))
calls = [d for d, n in trace for _ in range(n)]
ours = [e for e in rows.values() if "uegan::" in e["name"]]
ok = len(calls) == len(ours)
sys.stderr.write(f"{len(rows)} launches, {len(ours)} uegan::, trace {len(calls)} -> {'aligned' if ok else 'MISALIGNED'}\n")
for i, e in enumerate(ours):
    e["call"] = calls[i] if ok else "?"
tot = sum(e.get("gpu__time_duration.sum", 0) for e in rows.values()) / 1e6
agg = collections.OrderedDict()
for e in rows.values():
    short = e["name"].split("(")[0].replace("void ", "").replace("uegan::", "")[:48]
    key = (short, e.get("call", ""))
    a = agg.setdefault(key, {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "tp": 0.0})
    t = e.get("gpu__time_duration.sum", 0) / 1e6
    a["n"] += 1; a["ms"] += t
    a["rd"] += e.get("dram__bytes_read.sum", 0) / 1e6
    a["wr"] += e.get("dram__bytes_write.sum", 0) / 1e6
    a["tp"] += e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) * t
print(f"total {tot:.2f} ms in {len(rows)} launches (serialised, cold cache: compare shares)\n")
print("| kernel | call | n | ms | share | DRAM rd MB | wr MB | GB/s | tensor pipe % |")
print("|---|---|---|---|---|---|---|---|---|")
for (k, c), a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    if a["ms"] < 0.02:
        continue
    gbs = (a["rd"] + a["wr"]) / 1e3 / (a["ms"] * 1e-3) if a["ms"] else 0
    print(f"| {k} | {c} | {a['n']} | {a['ms']:.3f} | {100*a['ms']/tot:.1f}% | {a['rd']:.0f} | {a['wr']:.0f} | {gbs:.0f} | {a['tp']/a['ms'] if a['ms'] else 0:.1f} |")
