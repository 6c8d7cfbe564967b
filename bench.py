#!/usr/bin/env python
"""bench.py -- UEGAN hot path on B200.  One JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload train|inference]

Default workload = the configuration BASELINE.json's metric is quoted on: configs[2] "full G+D+VGG-perceptual training
step, 3x512x512 batch=16" per GPU (configs[3] at N=8: global batch 128).  One "step" = one iteration of the loop body
of trainer.py:75-119 (2 G forwards, 5 D forwards, 2 VGG forwards, all backwards, two Adam steps) on one batch of 16
synthetic 512x512 image pairs per GPU; weak scaling.  N > 1: torchrun, one process per GPU, two flat NCCL gradient
all-reduces per step (+ the tiny all-reduces of the relativistic means).
`--workload inference` = configs[1]: Generator forward on 32 x 3x512x512 per GPU (replicas, no collective).

* `value`      : images/s, all ranks, inputs resident in HBM, CUDA events, max over ranks.
* `e2e`        : same metric through the public API (`uegan_b200.trainer.Trainer.train_step` /
                 `models.Generator.__call__`) with the batch in pinned HOST memory: H2D of the batch and D2H of the
                 result (5 loss scalars / enhanced images) inside the timed region every step.
* `roofline`   : dominant kernel family = the tcgen05 implicit-GEMM convolutions (conv_fprop_kernel for fprop and
                 dgrad, conv_wgrad_kernel); achieved = algorithmic conv FLOPs of their launches in a step / their summed
                 CUDA-event durations (separate instrumented pass); peak = FLOP-weighted tensor peak (tf32 launches at
                 half the measured bf16 figure, fp16 launches at the full figure).
* `cpu_baseline` / `--impl reference`: the oracle port (oracle/uegan_oracle.py, fp32 torch-CPU restatement of the same
                 step) on the box's host cores -- the reference is Python and /root/reference is not on the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_GFLOP_PER_IMAGE = 67.747     # conv FLOPs of one Generator forward at 512x512
TRAIN_GFLOP_PER_IMAGE = 1098.9  # algorithmic conv FLOPs of one training step per image at 512x512
RES = 512


def _conv_tables():
    """(cin, cout, h_in, h_out) of every convolution on the path at 512 x 512 (1x1 up-convs hoisted before the
    bilinear x2, GAM fuse with its live half only -- the forms this implementation executes)."""
    G = [(3, 32, 512, 512), (32, 64, 512, 256), (64, 128, 256, 128), (128, 256, 128, 64), (256, 512, 64, 32),
         (512, 512, 32, 32),                                                     # ga5.fuse (live half)
         (512, 256, 32, 32), (256, 256, 64, 64), (512, 256, 64, 64),             # upsample1 conv, ga4.fuse, dec1
         (256, 128, 64, 64), (128, 128, 128, 128), (256, 128, 128, 128),         # upsample2, ga3, dec2
         (128, 64, 128, 128), (64, 64, 256, 256), (128, 64, 256, 256),           # upsample3, ga2, dec3
         (64, 32, 256, 256), (32, 32, 512, 512), (64, 32, 512, 512),             # upsample4, ga1, dec4
         (32, 32, 512, 512), (32, 3, 512, 512)]                                  # dec5.0, dec5.1
    D = [(3, 32, 512, 256), (32, 64, 256, 128), (64, 128, 128, 64), (128, 256, 64, 32), (256, 512, 32, 16),
         (32, 1, 256, 256), (64, 1, 128, 128), (128, 1, 64, 64), (256, 1, 32, 32), (512, 1, 16, 16)]
    V = [(3, 64, 512, 512), (64, 64, 512, 512), (64, 128, 256, 256), (128, 128, 256, 256), (128, 256, 128, 128),
         (256, 256, 128, 128), (256, 256, 128, 128), (256, 256, 128, 128), (256, 512, 64, 64), (512, 512, 64, 64),
         (512, 512, 64, 64), (512, 512, 64, 64), (512, 512, 32, 32)]
    return G, D, V


def training_step_algorithmic_bytes(bytes_per_elem=2):
    """Algorithmic HBM bytes of ONE training step per image at 512 x 512: every GEMM reads its two activation-sized
    operands / writes its output exactly once at 16-bit storage (fprop: x + y; dgrad: dz + dx; wgrad: x + dz), everything
    else (padding, activations, norms, up/down-sampling, losses) fused away, weights (35 MB per step, not per image)
    excluded.  Passes per step (trainer.py:75-119): G 2 fprop, 2 wgrad, 2 dgrad (no dgrad into the image); D 5 fprop,
    3 wgrad + 3 dgrad without the input layer (D step), 1 full dgrad (G step); VGG 2 fprop, 1 dgrad."""
    G, D, V = _conv_tables()
    io = lambda t: sum(ci * hi * hi + co * ho * ho for ci, co, hi, ho in t)
    first = lambda t: t[0][0] * t[0][2] ** 2 + t[0][1] * t[0][3] ** 2
    g = io(G) * (2 + 2) + (io(G) - first(G)) * 2
    d = io(D) * (5 + 3) + (io(D) - first(D)) * 3 + io(D) * 1
    v = io(V) * (2 + 1)
    return (g + d + v) * bytes_per_elem


def generator_algorithmic_bytes(bytes_per_elem=2):
    return sum(ci * hi * hi + co * ho * ho for ci, co, hi, ho in _conv_tables()[0]) * bytes_per_elem


def measured_traffic(train, precision):
    """DRAM bytes per step of the GEMM launches, from the committed ncu pass of THIS build
    (profiles/r2_step_traffic.json, written by scripts/ncu_traffic.py from the per-launch CSV: dram__bytes_read.sum +
    dram__bytes_write.sum); None when no capture of this workload / precision is committed."""
    p = os.path.join(ROOT, "profiles", "r2_step_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    e = d.get(f"{'train' if train else 'inference'}_{precision}")
    if not e:
        return None, None
    return e.get("gemm_dram_bytes_per_step"), e.get("source")


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
        sm, mx, reasons, power = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); power.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []  # samples under load (the idle edges pull a plain median down)
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def train_args(batch):
    return types.SimpleNamespace(
        g_conv_dim=32, d_conv_dim=32, g_norm_fun="none", d_norm_fun="none", g_act_fun="LeakyReLU",
        d_act_fun="LeakyReLU", g_use_sn=False, d_use_sn=True, adv_loss_type="rahinge", init_type="", optimizer_type="adam",
        g_lr=1e-4, d_lr=4e-4, beta1=0.5, beta2=0.999, alpha=0.9, lr_decay=False, pool_size=0, adv_input=True,
        lambda_adv=0.10, lambda_percep=1.0, lambda_idt=0.10, idt_loss_type="l1", save_root_dir="/tmp/uegan_b200",
        version="bench", model_save_path="models", train_batch_size=batch, total_epochs=1, pretrained_model=0.0,
        model_save_epoch=1, info_step=100, cuda_graph=False)


# ------------------------------------------------------------------------------------------------ CPU / library baselines
def reference_available():
    from oracle.make_ref import ref_dir
    return ref_dir() is not None


def baseline_step_fn(workload, batch, device="cpu"):
    """One iteration of the workload by the REFERENCE's own code (oracle/ref_step.py over the vendored oracle/_ref:
    kind "reference"), or, if the vendored copy is missing, by the oracle port (kind "port", CPU only)."""
    if reference_available():
        from oracle import ref_step
        return ref_step.make_step(workload, batch, device, RES), "reference"
    if device != "cpu":
        return None, None
    import torch
    from oracle import uegan_oracle as O
    gp = O.make_generator_params(32, 0, "o1")
    x = O.make_images((batch, 3, RES, RES), 0)
    if workload == "inference":
        def step():
            with torch.no_grad():
                O.generator_forward(gp, x)
        return step, "port"
    dp, vp = O.make_discriminator_params(32, 1, "o1"), O.make_vgg_params()
    g_opt, d_opt = O.AdamState(O._trainable(gp)), O.AdamState(O._trainable(dp))
    y = O.make_images((batch, 3, RES, RES), 1)

    def step():
        O.train_step(gp, dp, vp, g_opt, d_opt, x, y)
    return step, "port"


def cpu_rate(workload, batch, iters, threads, warm=0):
    import torch
    torch.set_num_threads(threads)
    step, kind = baseline_step_fn(workload, batch)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = time.perf_counter() - t0
    return batch * iters / dt, dt, kind


def gpu_library_rate(workload, batch, iters=3, warm=2):
    """The reference's own PyTorch code on THIS GPU: eager ATen + cuDNN, fp32 tensors with torch's default TF32 convolutions,
    cudnn.benchmark=True (main.py:16) -- the existing kernels this repo has to beat on the same box."""
    import torch
    step, kind = baseline_step_fn(workload, batch, "cuda")
    if step is None:
        return None
    try:
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        return {"value": batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "batch": batch, "kind": kind,
                "what": "unmodified reference models.py/losses.py + torch.optim.Adam on cuda: eager PyTorch "
                        f"{torch.__version__} + cuDNN {torch.backends.cudnn.version()}, cudnn.benchmark=True, "
                        f"allow_tf32(conv)={torch.backends.cudnn.allow_tf32}, {iters} steps after {warm} warm-up"}
    except Exception as e:  # noqa: BLE001 -- a baseline that cannot run must not take the native measurement down with it
        return {"value": None, "error": f"{type(e).__name__}: {e}"[:300]}
    finally:
        del step
        torch.cuda.empty_cache()


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    batch = 1 if args.workload == "train" else 2
    import torch
    torch.set_num_threads(threads)
    step, kind = baseline_step_fn(args.workload, batch)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        step()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    val = batch / (ms / 1e3)
    what = "training step (trainer.py:75-119)" if args.workload == "train" else "Generator.forward"
    src = ("the unmodified reference modules (oracle/_ref: models.py, losses.py, torch.optim.Adam)" if kind == "reference"
           else "oracle port")
    sample = f"{src}, {what}, {batch}x3x{RES}x{RES} per step, {threads} threads, fp32"
    print(json.dumps({
        "impl": "reference", "metric": "512x512 training images/sec" if args.workload == "train" else "512x512 images/sec",
        "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args.workload), "batch_per_step": batch,
                                        "note": "CPU steps are a bounded sample of the GPU arm's per-step batch"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(w):
    return ("train_step_G+D+VGG_3x512x512_b16 (BASELINE.json configs[2]; configs[3] at 8 GPUs)" if w == "train"
            else "generator_inference_3x512x512_b32 (BASELINE.json configs[1])")


def _claim_stdout():
    """Returns a duplicate of the original stdout and points fd 1 at stderr for the rest of the process."""
    sys.stdout.flush()
    fd = os.dup(1)
    os.dup2(2, 1)
    return fd


# ------------------------------------------------------------------------------------------------ GPU: config 5
def run_sweep(args):
    """BASELINE.json configs[4]: mixed-resolution Generator inference sweep (short edge 256 / 512 / 1024, 2:3 aspect),
    batch sized to fill HBM, one replica per GPU, no communication."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    json_fd = _claim_stdout()
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import uegan_oracle as O
    from uegan_b200 import kernels as K
    from uegan_b200.models import Generator
    G = Generator(32, "none", "LeakyReLU", False)
    G.load_state_dict(O.make_generator_params(32, 0, "o1"))
    G = G.cuda().eval()
    free, _ = torch.cuda.mem_get_info()
    rows, tot_pix, tot_ms = [], 0.0, 0.0
    for (h, w) in ((256, 384), (512, 768), (1024, 1536)):
        per_img = (1500.0 if G.precision == "tf32" else 800.0) * h * w  # bytes of activation buffers per image (all Generator tensors)
        b = int(min(256, max(1, 0.5 * free / per_img)))
        x = torch.rand(b, 3, h, w, device="cuda") * 2 - 1
        with torch.no_grad():
            for _ in range(max(args.warmup, 3)):
                G(x)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                G(x)
            e1.record()
            torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        rows.append({"h": h, "w": w, "batch_per_gpu": b, "ms_per_step": ms, "images_per_s": world * b / (ms * 1e-3),
                     "mpix_per_s": world * b * h * w / 1e6 / (ms * 1e-3)})
        tot_pix += world * b * h * w / 1e6
        tot_ms += ms
        G._plans.clear()
        del x
        torch.cuda.empty_cache()
    if rank == 0:
        os.write(json_fd, (json.dumps({
            "metric": "Generator inference megapixels/sec (mixed-resolution sweep)", "value": tot_pix / (tot_ms * 1e-3),
            "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": G.precision,
            "data": "synthetic", "config": {"workload": "generator_inference_sweep_256_512_1024 (BASELINE.json configs[4])",
                                            "parallelism": f"replicas x{world}", "sweep": rows},
            "gpu_launches": K.launches()}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ GPU (native)
def run_native(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    group = None
    # stdout carries exactly ONE JSON line: native libraries (NCCL's "NCCL version ..." banner, symmetric-memory init) write
    # to file descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
    # duplicate of the original stdout at the end
    json_fd = _claim_stdout()
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    from oracle import uegan_oracle as O  # deterministic synthetic WEIGHTS only; nothing of the oracle is timed here
    from uegan_b200 import kernels as K
    train = args.workload == "train"
    batch = args.batch or (16 if train else 32)
    gen = torch.Generator(device="cuda").manual_seed(rank)
    x = torch.rand(batch, 3, RES, RES, device="cuda", generator=gen) * 2 - 1
    x_host = x.cpu().pin_memory()
    if train:
        from uegan_b200.trainer import Trainer
        targs = train_args(batch)
        # The data-parallel reductions run over peer memory inside our own kernels (uegan_b200.peer / optim), so the
        # captured step contains no NCCL call and the CUDA graph is used at every world size; with the NCCL fallback
        # (no symmetric memory) the step runs eagerly at N > 1.
        targs.peer_reduce = bool(args.peer)
        T = Trainer(None, targs, process_group=group if world > 1 else None,
                    vgg_state_dict=O.make_vgg_params())
        use_graph = bool(args.graph) and (world == 1 or T.comm is not None)
        targs.cuda_graph = use_graph
        T.G.load_state_dict(O.make_generator_params(32, 0, "o1"))
        T.D.load_state_dict(O.make_discriminator_params(32, 1, "o1"))
        y = torch.rand(batch, 3, RES, RES, device="cuda", generator=gen) * 2 - 1
        y_host = y.cpu().pin_memory()
        graphed = False
        if use_graph:
            try:
                T.capture(x, y)
                graphed = True
            except Exception as e:  # noqa: BLE001 -- keep measuring (eagerly) if capture is not possible
                sys.stderr.write(f"[bench] CUDA graph capture failed, running eagerly: {type(e).__name__}: {e}\n")
                torch.cuda.synchronize()

        def step_resident():
            if graphed:
                T.replay(x, y)
            else:
                T.train_step(x, y, sync_scalars=False)

        def step_e2e():
            if graphed:
                # every step: H2D of a batch from pinned host memory, the graph, D2H of the five loss scalars.  The copy of
                # batch k + 1 is issued right after graph k is launched and overlaps it (Trainer.prefetch: a one-batch-ahead
                # input pipeline); the first batch is staged by the warm-up call.
                if not getattr(T, "_staged", False):
                    T.prefetch(x_host, y_host)
                out = T.replay(None, None, sync_scalars=False)
                T.prefetch(x_host, y_host)
                keys = list(out.keys())
                return dict(zip(keys, torch.stack([out[k].detach().reshape(()).float() for k in keys]).tolist()))
            xd, yd = x_host.cuda(non_blocking=True), y_host.cuda(non_blocking=True)
            return T.train_step(xd, yd, sync_scalars=True)  # 5 x .item(): the reference's D2H reads (trainer.py:98-119)
        h2d, d2h = 2 * x_host.numel() * 4, 5 * 4
        gflop = TRAIN_GFLOP_PER_IMAGE
        ctx = torch.enable_grad
    else:
        from uegan_b200.models import Generator
        G = Generator(32, "none", "LeakyReLU", False)
        G.load_state_dict(O.make_generator_params(32, 0, "o1"))
        G = G.cuda().eval()
        # e2e = the deployment path: uint8 HWC images in pinned host memory -> H2D -> uegan_pack_input_u8 -> Generator ->
        # uegan_unpack_output_u8 -> D2H, double-buffered so that the copies overlap the compute (uegan_b200.io.U8Pipeline;
        # SURVEY.md 8f N3).  Every step moves a fresh batch in and its result out inside the timed region.
        from uegan_b200 import io as IO
        img_host = ((x_host + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
        out_host = [torch.empty_like(img_host).pin_memory() for _ in range(2)]
        pipe = IO.U8Pipeline(G)
        e2e_i = [0]

        def step_resident():
            return G(x)

        def step_e2e():
            pipe.submit(img_host, out_host[e2e_i[0] % 2])
            e2e_i[0] += 1
        h2d, d2h = img_host.numel(), img_host.numel()
        gflop = G_GFLOP_PER_IMAGE
        ctx = torch.no_grad
        graphed = False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, steps, warmup, finish=None):
        with ctx():
            for _ in range(warmup):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = K.launches()
            e0.record()
            h0 = time.perf_counter()
            for _ in range(steps):
                fn()
            host_ms[0] = (time.perf_counter() - h0) * 1e3 / steps  # time to ENQUEUE a step (no sync inside fn)
            if finish is not None:
                finish()  # e.g. the timing stream waits for the last device->host copy of a pipelined run
            e1.record()
            barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), K.launches() - l0

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, launches = timed(step_resident, args.steps, args.warmup)
    host_enqueue_ms = host_ms[0]
    clocks = sampler.stop() if sampler else None
    fin = None if train else (lambda: torch.cuda.current_stream().wait_stream(pipe.s_out))
    ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 1), finish=fin)

    roof = None
    # instrumented pass after the timed regions: CUDA events around every GEMM launch.  Every rank runs it (the step
    # contains collectives); rank 0 reports.
    reps = 2
    K._Counters.conv_events = []
    # (the timed step overlaps the weight-gradient launches with the data-gradient chain on a second stream; this pass runs
    # everything on one stream so that each event pair brackets one launch running alone)
    from uegan_b200 import autograd as _AG
    side_was, _AG._Side.enabled = _AG._Side.enabled, False
    with ctx():
        for _ in range(reps):
            if train:
                T.train_step(x, y, sync_scalars=False)  # eager: events wrap individual launches
            else:
                step_resident()
    barrier()
    ev = K._Counters.conv_events
    K._Counters.conv_events = None
    _AG._Side.enabled = side_was
    if rank == 0:
        hbm, tf_burst, tf_sus, src = peaks()
        tot_ms = tot_fl = t_at_peak = 0.0
        per, groups = {}, {}
        for a, b, f, xt, cout, k, s, kind, dt in ev:
            ms = a.elapsed_time(b) / reps
            pk = tf_burst / 2 if dt == 0 else tf_burst
            tot_ms += ms; tot_fl += f / reps; t_at_peak += (f / reps) / (pk * 1e12)
            key = f"{kind} {'tf32' if dt == 0 else 'f16'} {xt.c}->{cout} k{k}s{s} @{xt.h}"
            e = per.setdefault(key, [0.0, 0.0]); e[0] += ms; e[1] += f / reps
            g = groups.setdefault(f"{kind} {'tf32' if dt == 0 else 'f16'}", [0.0, 0.0]); g[0] += ms; g[1] += f / reps
        achieved = tot_fl / (tot_ms * 1e-3) / 1e12
        peak = tot_fl / t_at_peak / 1e12
        prec = "f16" if (T.G.precision if train else G.precision) == "f16" else "tf32"
        traffic, traffic_src = measured_traffic(train, prec)
        alg_bytes = batch * (training_step_algorithmic_bytes() if train else generator_algorithmic_bytes())
        top = sorted(per.items(), key=lambda kv: -kv[1][0])[:24]
        roof = {"bound": "tensor", "kernel": "conv_fprop_kernel (fprop+dgrad) + conv_wgrad_kernel, all launches of a step",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                # DRAM bytes of these launches per step (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu
                # pass of this build, next to the algorithmic bytes of the same step
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": alg_bytes,
                "traffic_over_algorithmic": (traffic / alg_bytes) if (traffic and alg_bytes) else None,
                "peak_source": f"{src}: FLOP-weighted mix of bf16_tflops {tf_burst} (fp16 launches) and half of it (tf32 launches)",
                "gemm_ms_per_step": tot_ms, "step_ms": ms_total / args.steps, "gemm_launches_per_step": len(ev) // reps,
                "gemm_share_of_step": tot_ms / (ms_total / args.steps),
                "algorithmic_gflop_per_step_measured": tot_fl / 1e9,
                "by_kind_ms_tflops": {k: [round(v[0], 2), round(v[1] / (v[0] * 1e-3) / 1e12, 1)] for k, v in groups.items()},
                "top_layers_ms_tflops": {k: [round(v[0], 3), round(v[1] / (v[0] * 1e-3) / 1e12, 1)] for k, v in top}}

    replicas = None
    if train:
        # data-parallel consistency: after all steps every rank must hold bit-identical weights (the rank-ordered peer-memory
        # sums guarantee it; with NCCL the all-reduce does).  One checksum per rank, gathered; rank 0 reports.
        import hashlib
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().flatten() for p in list(T.G.parameters()) + list(T.D.parameters())]).cpu().numpy()
        digest = hashlib.sha256(flat.tobytes()).hexdigest()[:16]
        if world > 1:
            allh = [None] * world
            dist.all_gather_object(allh, digest)
        else:
            allh = [digest]
        replicas = {"weights_sha256_16": allh[0], "identical_on_all_ranks": len(set(allh)) == 1,
                    "grad_reduce": ("single GPU" if world == 1 else
                                    ("peer memory (NVLink), fused into uegan_adam_step_peers" if T.comm is not None
                                     else "NCCL all-reduce + uegan_adam_step"))}
    gd_prec = (T.G.precision if train else G.precision)
    gd_prec = "f16 operands with per-tensor 2^k scales" if gd_prec == "f16" else "tf32"
    if rank == 0:
        cpu_threads = os.cpu_count() or 1
        cpu_batch = 1 if train else 2
        # the reference's own PyTorch/cuDNN path on this GPU (N = 1 runs only: the scaling runs stay short)
        gpu_lib = gpu_library_rate(args.workload, batch) if (world == 1 and args.lib_baseline) else None
        cpu_val, cpu_dt, cpu_kind = cpu_rate(args.workload, cpu_batch, 1 if train else 2, cpu_threads, warm=0 if train else 1)
        ms_step = ms_total / args.steps
        value = world * batch / (ms_step * 1e-3)
        e2e = world * batch / (ms_e2e / args.steps * 1e-3)
        hbm, tf_burst, tf_sus, src = peaks()
        sys.stdout.flush()
        os.write(json_fd, (json.dumps({
            "metric": "512x512 training images/sec" if train else "512x512 images/sec", "value": value,
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": (f"{gd_prec} (G, D) + f16 (VGG), fp32 accumulate" if train else gd_prec), "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "batch_per_gpu": batch, "global_batch": world * batch,
                       "parallelism": (f"dp{world}: per optimizer one fused gradient-reduction + Adam kernel over "
                                       f"{'peer memory' if (world > 1 and T.comm is not None) else ('NCCL' if world > 1 else 'one GPU')}"
                                       if train else f"replicas x{world}"),
                       "l2": "per-step activation traffic (tens of GB) >> 126 MB L2; no flush needed",
                       "host_enqueue_ms_per_step": host_enqueue_ms, "cuda_graph": bool(train and graphed),
                       "streams": (2 if (train and __import__("uegan_b200.autograd", fromlist=["_Side"])._Side.enabled) else 1),
                       "achieved_tflops_per_gpu": gflop * value / world / 1e3,
                       "frac_of_bf16_peak": gflop * value / world / 1e3 / tf_burst},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "cpu_baseline": {"value": cpu_val, "unit": "images/s", "cores": cpu_threads, "kind": cpu_kind,
                             "sample": f"{'unmodified reference modules (oracle/_ref)' if cpu_kind == 'reference' else 'oracle port'}"
                                       f" {'training step' if train else 'Generator.forward'}, "
                                       f"{1 if train else 2} iteration(s) of {cpu_batch}x3x512x512 ({cpu_dt:.1f} s)"},
            "gpu_library_baseline": gpu_lib, "replicas": replicas,
        }) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "inference", "sweep"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default 16 train / 32 inference)")
    ap.add_argument("--lib-baseline", dest="lib_baseline", type=int, default=1,
                    help="also time the unmodified reference on the GPU (eager PyTorch + cuDNN), N = 1 only")
    ap.add_argument("--peer", type=int, default=1, help="N > 1: gradient / loss reductions over peer memory inside our "
                    "kernels (1) or NCCL all-reduce (0)")
    ap.add_argument("--graph", type=int, default=1, help="capture the training step into a CUDA graph (1) or run eagerly (0)")
    args = ap.parse_args()
    if args.workload == "sweep":
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
