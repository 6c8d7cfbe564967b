#!/usr/bin/env python
"""bench.py -- UEGAN hot path on B200.  One JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload inference|train]

Workload (config.workload): BASELINE.json configs[1] "Generator inference, synthetic 3x512x512 batch=32, 1xB200"
while the native training step (configs[2]) is being built; one "step" = one pass of the Generator over one batch of
32 synthetic 512x512 images per GPU (weak scaling: every rank runs its own batch, no data-path collective --
inference is embarrassingly parallel).

* `value`      : images/s, all ranks, inputs resident in HBM, CUDA events, max over ranks.
* `e2e`        : same metric through the public API (`uegan_b200.models.Generator.__call__`) with the batch in pinned
                 HOST memory: H2D of the batch and D2H of the enhanced images are inside the timed region every step.
* `roofline`   : dominant kernel = conv_fprop_kernel (tcgen05 implicit GEMM); achieved = algorithmic conv FLOPs of all
                 its launches in a step / their summed CUDA-event durations (measured in a separate instrumented pass).
* `cpu_baseline`: the oracle port (oracle/uegan_oracle.py, fp32 torch-CPU restatement of models.py:44-74) on the box's
                 host cores, bounded sample.
* `--impl reference`: the same oracle port as the reference arm (the reference is Python; /root/reference does not
                 exist on the GPU box), all host threads, same metric/unit/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_GFLOP_PER_IMAGE = 67.747  # conv FLOPs (2*MAC) of one Generator forward at 512x512, SURVEY.md 8(d)
BATCH = 32
RES = 512


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the samples taken under load (upper half: idle samples at the edges pull the median down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_generator_rate(batch, iters, threads):
    """images/s of the CPU oracle port of Generator.forward at 512x512 (bounded sample)."""
    import torch
    from oracle import uegan_oracle as O
    torch.set_num_threads(threads)
    gp = O.make_generator_params(32, 0, "o1")
    x = O.make_images((batch, 3, RES, RES), 0)
    with torch.no_grad():
        O.generator_forward(gp, x[:1])  # warm-up
        t0 = time.perf_counter()
        for _ in range(iters):
            O.generator_forward(gp, x)
        dt = time.perf_counter() - t0
    return batch * iters / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 2
    times = []
    import torch
    from oracle import uegan_oracle as O
    torch.set_num_threads(threads)
    gp = O.make_generator_params(32, 0, "o1")
    x = O.make_images((batch, 3, RES, RES), 0)
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.generator_forward(gp, x)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    val = batch / (ms / 1e3)
    sample = f"oracle port of Generator.forward, {batch}x3x{RES}x{RES} per step, {threads} threads, fp32"
    print(json.dumps({
        "impl": "reference", "metric": "512x512 images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "generator_inference_3x512x512 (BASELINE.json configs[1])", "batch_per_step": batch,
                   "note": "CPU steps are a bounded sample (batch 2) of the batch-32 GPU step"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_native(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import uegan_oracle as O  # synthetic weights/images only (checker code is not timed as product)
    from uegan_b200 import kernels as K
    from uegan_b200.models import Generator

    G = Generator(32, "none", "LeakyReLU", False)
    G.load_state_dict(O.make_generator_params(32, 0, "o1"))
    G = G.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(rank)
    x = torch.rand(BATCH, 3, RES, RES, device="cuda", generator=g) * 2 - 1
    x_host = x.cpu().pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = K.launches()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), K.launches() - l0

    def step_resident():
        return G(x)

    def step_e2e():
        xd = x_host.cuda(non_blocking=True)
        out = G(xd)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 1))

    # ---- instrumented pass: CUDA events around every conv launch (dominant kernel), rank 0
    roof = None
    if rank == 0:
        K._Counters.conv_events = []
        with torch.no_grad():
            for _ in range(3):
                G(x)
        torch.cuda.synchronize()
        ev = K._Counters.conv_events
        K._Counters.conv_events = None
        conv_ms = sum(a.elapsed_time(b) for a, b, *_ in ev) / 3
        conv_flops = sum(f for _, _, f, *_ in ev) / 3
        hbm, tf_burst, tf_sus, src = peaks()
        peak = tf_burst / 2  # kind::tf32 runs at half the bf16 rate; MEASURED_PEAKS.json holds the bf16 figure
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        per_layer = {}
        for a, b, f, xt, cout, k, s in ev:
            key = f"{xt.c}->{cout} k{k}s{s} @{xt.h}"
            per_layer[key] = per_layer.get(key, 0.0) + a.elapsed_time(b) / 3
        roof = {"bound": "tensor", "kernel": "conv_fprop_kernel<tf32> (all 25 launches of a step)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": f"{src}: bf16_tflops {tf_burst} / 2 for kind::tf32",
                "conv_ms_per_step": conv_ms, "step_ms": ms_total / args.steps,
                "conv_share_of_step": conv_ms / (ms_total / args.steps),
                "per_layer_ms": {k: round(v, 4) for k, v in sorted(per_layer.items(), key=lambda kv: -kv[1])}}

    if rank == 0:
        cpu_threads = os.cpu_count() or 1
        cpu_val, cpu_dt = oracle_generator_rate(2, 2, cpu_threads)
        ms_step = ms_total / args.steps
        value = world * BATCH / (ms_step * 1e-3)
        e2e = world * BATCH / (ms_e2e / args.steps * 1e-3)
        print(json.dumps({
            "metric": "512x512 images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "generator_inference_3x512x512 (BASELINE.json configs[1])", "batch_per_gpu": BATCH,
                       "global_batch": world * BATCH, "parallelism": f"replicas x{world}",
                       "l2": "activation traffic per step (GBs) >> 126 MB L2; no flush needed",
                       "achieved_tflops_per_gpu": G_GFLOP_PER_IMAGE * value / world / 1e3},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "cpu_baseline": {"value": cpu_val, "unit": "images/s", "cores": cpu_threads, "kind": "port",
                             "sample": f"oracle Generator.forward, 2 iterations of 2x3x512x512 ({cpu_dt:.1f} s)"},
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
