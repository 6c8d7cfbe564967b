"""Host-side wrappers over the C ABI: NHWC-with-halo activation buffers (torch owns the memory, the kernels are
ours) and one Python function per exported entry point.  torch is plumbing here: allocator + stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

_TORCH_DTYPE = {L.F32: torch.float32, L.BF16: torch.bfloat16, L.F16: torch.float16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Counters:
    launches = 0          # kernels of libuegan_sm100.so launched through this module
    conv_events = None    # when a list: (start, end, flops) CUDA-event triples around every conv launch
    trace = None          # when a list: (description, launches) per wrapper call, in issue order (profiling only)


def _count(n: int, what: str = "", *tensors):
    _Counters.launches += n
    tr = _Counters.trace
    if tr is not None:
        dims = " ".join(f"{t.n}x{t.h}x{t.w}x{t.c}h{t.halo}{'f' if t.dtype == L.F32 else 'h'}" for t in tensors
                        if t is not None)
        tr.append((f"{what} {dims}".strip(), n))


def launches() -> int:
    return _Counters.launches


class NHWC:
    """Activation buffer: n x (h+2*halo) x (w+2*halo) x c, dtype F32 (tf32 math) or BF16, plus zeroed slack so
    a 128-byte TMA window that starts at the last pixel never leaves the allocation."""

    SLACK = 512

    def __init__(self, n, h, w, c, halo=0, dtype=L.F32, device="cuda", zero=False, scale=None):
        """scale: optional 1-element fp32 CUDA tensor (a ScaleBook slot): stored values = scale * true values."""
        es = 4 if dtype == L.F32 else 2
        assert (c * es) % 16 == 0, "channel vector must be a multiple of 16 bytes"
        self.n, self.h, self.w, self.c, self.halo, self.dtype = n, h, w, c, halo, dtype
        elems = n * (h + 2 * halo) * (w + 2 * halo) * c
        alloc = torch.zeros if zero else torch.empty
        self.buf = alloc(elems + self.SLACK // es, dtype=_TORCH_DTYPE[dtype], device=device)
        if not zero:
            self.buf[elems:].zero_()
        self.elems = elems
        self.scale = scale
        self.ct = L.Tensor(self.buf.data_ptr(), n, h, w, c, halo, dtype, scale.data_ptr() if scale is not None else None)

    def ref(self):
        return C.byref(self.ct)

    def as_haloed(self, halo: int) -> "NHWC":
        """The same memory read as a tensor whose outer `halo` pixels are a halo: a halo-0 buffer of extent
        (h + 2*halo) x (w + 2*halo) (the padded gradient a dgrad launch writes) becomes interior h x w + halo."""
        assert self.halo == 0 and self.h > 2 * halo and self.w > 2 * halo
        v = object.__new__(NHWC)
        v.n, v.h, v.w, v.c, v.halo, v.dtype = self.n, self.h - 2 * halo, self.w - 2 * halo, self.c, halo, self.dtype
        v.buf, v.elems, v.scale = self.buf, self.elems, self.scale
        v.ct = L.Tensor(self.buf.data_ptr(), v.n, v.h, v.w, v.c, halo, self.dtype,
                        self.scale.data_ptr() if self.scale is not None else None)
        return v

    def padded_view(self) -> torch.Tensor:
        return self.buf[: self.elems].view(self.n, self.h + 2 * self.halo, self.w + 2 * self.halo, self.c)

    def interior_nchw(self) -> torch.Tensor:
        """Interior as a float32 NCHW torch tensor of TRUE values (test/debug readback; not on the hot path)."""
        v = self.padded_view()
        p = self.halo
        v = v[:, p:p + self.h, p:p + self.w, :]
        v = v.permute(0, 3, 1, 2).float().contiguous()
        return v / self.scale if self.scale is not None else v


class ScaleBook:
    """Device-resident power-of-two scales (uegan_tensor.scale) of a set of tensors and the table uegan_scale_update walks.
    `slot()` hands out 1-element views of one fp32 buffer (all 1.0 initially); `track()` registers the buffer whose stored
    values a slot describes; `update()` is ONE launch that re-derives every tracked slot from a strided sample."""

    TARGET = 1024.0      # 2^10: 64x below fp16's 65504 (samples underestimate the maximum; magnitudes drift between passes)
    SAMPLES = 1 << 16

    def __init__(self, device, capacity=512):
        self.buf = torch.ones(capacity, dtype=torch.float32, device=device)
        self.n = 0
        self.entries = []
        self.table = None

    def slot(self) -> torch.Tensor:
        assert self.n < self.buf.numel(), "ScaleBook capacity"
        t = self.buf[self.n:self.n + 1]
        self.n += 1
        return t

    def track(self, data_ptr: int, numel: int, dtype: int, slot: torch.Tensor, true_values: bool = False):
        """true_values: the buffer holds unscaled values (fp32 master weights); the slot is the scale their packed copy gets."""
        self.entries.append((int(data_ptr), int(numel), int(dtype), slot.data_ptr(), 1 if true_values else 0))
        self.table = None

    def track_nhwc(self, t: "NHWC"):
        self.track(t.buf.data_ptr(), t.elems, t.dtype, t.scale)

    def clear_tracked(self):
        self.entries, self.table = [], None

    def update(self):
        if not self.entries:
            return
        if self.table is None:
            arr = (L.ScaleEntry * len(self.entries))()
            for i, (ptr, numel, dtype, sl, tv) in enumerate(self.entries):
                arr[i].data, arr[i].numel, arr[i].dtype, arr[i].reserved, arr[i].scale = ptr, numel, dtype, tv, sl
            raw = bytes(arr)
            self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.buf.device)
        L.check(L.load().uegan_scale_update(self.table.data_ptr(), len(self.entries), self.TARGET, self.SAMPLES, _stream()),
                "scale_update")
        _count(1, "scale_update")

    def values(self) -> torch.Tensor:
        return self.buf[:self.n].detach().cpu()


def packed_weight(weight: torch.Tensor, cin_stored: int, dtype: int, cin_first: int = 0, cin: Optional[int] = None,
                  transpose_flip: bool = False, out: Optional[torch.Tensor] = None,
                  w_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Conv2d.weight (OIHW fp32, CUDA) -> K-major packed operand of the implicit GEMM.  `out`: re-pack in place
    (the operand buffers must keep their addresses across optimizer steps for CUDA-graph replay).  w_scale: device scalar
    the values are multiplied by (fp16 operands of weights whose magnitude is far from 1; pass it to the conv as well)."""
    lib = L.load()
    w = weight.detach()
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
    o, i_total, k, k2 = w.shape
    assert k == k2
    if cin is None:
        cin = i_total - cin_first
    cout = i_total if transpose_flip else o
    if transpose_flip:
        cin = o
    nbytes = lib.uegan_packed_weight_bytes(cout, cin_stored, k, dtype)
    buf = out if out is not None else torch.empty(nbytes + 256, dtype=torch.uint8, device=w.device)
    assert buf.numel() >= nbytes
    if w_scale is not None:
        assert not transpose_flip
        L.check(lib.uegan_pack_conv_weight_scaled(w.data_ptr(), buf.data_ptr(), cout, i_total, cin_first, cin, cin_stored, k,
                                                  dtype, w_scale.data_ptr(), _stream()), "pack_conv_weight_scaled")
    else:
        L.check(lib.uegan_pack_conv_weight(w.data_ptr(), buf.data_ptr(), cout, i_total, cin_first, cin, cin_stored, k,
                                           dtype, int(transpose_flip), _stream()), "pack_conv_weight")
    _count(1, "pack_weight")
    return buf


def conv_fprop(x: NHWC, w_packed: torch.Tensor, cout: int, k: int, stride: int, pad: int, y: Optional[NHWC] = None,
               y_c_off: int = 0, bias: Optional[torch.Tensor] = None, alpha: Optional[torch.Tensor] = None,
               act: int = L.ACT_NONE, mul: Optional[NHWC] = None, out_nchw: Optional[torch.Tensor] = None,
               residual_nchw: Optional[torch.Tensor] = None, in_stats: Optional[torch.Tensor] = None,
               aux_nchw: Optional[torch.Tensor] = None, w_scale: Optional[torch.Tensor] = None,
               premul: Optional[NHWC] = None, reflect_halo: bool = False, halo_launch: bool = False):
    """premul: with `mul`, the output before the multiplication is stored there as well.
    reflect_halo: also write y's reflection-padding halo: mirrored stores in the conv epilogue, or -- where the epilogue
    cannot, or with halo_launch -- a halo_fill launch after the conv."""
    lib = L.load()
    d = L.ConvDesc()
    d.x = x.ct
    d.y_premul = C.pointer(premul.ct) if premul is not None else None
    fill_after = False
    if reflect_halo and y is not None and y.halo > 0:
        import os
        if (not halo_launch and y.dtype != L.F32 and y.h > 2 * y.halo + 1 and y.w > 2 * y.halo + 1 and in_stats is None
                and os.environ.get("UEGAN_NO_EPILOGUE_HALO") != "1"):
            d.y_reflect_halo = 1
        else:
            fill_after = True
    if y is not None:
        d.y = y.ct
    d.y_c_off, d.cout, d.k, d.stride, d.pad, d.act = y_c_off, cout, k, stride, pad, act
    d.w_packed = w_packed.data_ptr()
    d.w_scale = w_scale.data_ptr() if w_scale is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.alpha = alpha.data_ptr() if alpha is not None else None
    d.mul = C.pointer(mul.ct) if mul is not None else None
    d.out_nchw = out_nchw.data_ptr() if out_nchw is not None else None
    d.residual_nchw = residual_nchw.data_ptr() if residual_nchw is not None else None
    d.in_stats = in_stats.data_ptr() if in_stats is not None else None
    d.aux_nchw = aux_nchw.data_ptr() if aux_nchw is not None else None
    ev = _Counters.conv_events
    if ev is not None:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
    L.check(lib.uegan_conv2d_fprop(C.byref(d), _stream()), "conv2d_fprop")
    if ev is not None:
        s1.record()
        ho = (x.h + 2 * pad - k) // stride + 1
        wo = (x.w + 2 * pad - k) // stride + 1
        cin = 3 if x.c * (4 if x.dtype == L.F32 else 2) == 16 else x.c
        ev.append((s0, s1, 2.0 * x.n * ho * wo * cout * k * k * cin, x, cout, k, stride, "fprop", x.dtype))
    _count(1, f"fprop ->{cout} k{k}s{stride}{' +stats' if in_stats is not None else ''}", x)
    if fill_after:
        halo_fill(y, L.PAD_REFLECT)


def rowsum_supported(cout: int, cin_stored: int, k: int, dtype: int) -> bool:
    return bool(L.load().uegan_conv2d_rowsum_supported(cout, cin_stored, k, dtype))


def packed_weight_rowsum(weight: torch.Tensor, cin_stored: int, out: Optional[torch.Tensor] = None, dtype: int = L.F32,
                         w_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Operand of the row-sum kernel (csrc/conv_rowsum.cu): rows (s, o), columns (r, c); tf32, or fp16 times *w_scale."""
    lib = L.load()
    w = weight.detach()
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
    o, i_total, k, _ = w.shape
    nbytes = lib.uegan_packed_weight_rowsum_bytes(o, cin_stored, k)
    buf = out if out is not None else torch.empty(nbytes + 256, dtype=torch.uint8, device=w.device)
    assert buf.numel() >= nbytes
    if dtype != L.F32 or w_scale is not None:
        L.check(lib.uegan_pack_conv_weight_rowsum_scaled(w.data_ptr(), buf.data_ptr(), o, i_total, 0, i_total, cin_stored, k,
                                                         dtype, w_scale.data_ptr() if w_scale is not None else None,
                                                         _stream()), "pack_conv_weight_rowsum_scaled")
    else:
        L.check(lib.uegan_pack_conv_weight_rowsum(w.data_ptr(), buf.data_ptr(), o, i_total, 0, i_total, cin_stored, k,
                                                  _stream()), "pack_conv_weight_rowsum")
    _count(1, "pack_weight_rowsum")
    return buf


def conv_planar(x: NHWC, weight: torch.Tensor, cache, key, k: int, pad: int, bias, alpha, act: int,
                out_nchw: torch.Tensor, residual_nchw: Optional[torch.Tensor] = None,
                aux_nchw: Optional[torch.Tensor] = None, w_scale: Optional[torch.Tensor] = None):
    """Stride-1 conv with a tiny output-channel count written as fp32 NCHW planes (G's last conv, D's heads): the
    row-sum kernel when the shape qualifies, else the generic implicit GEMM."""
    cout = weight.shape[0]
    if rowsum_supported(cout, x.c, k, x.dtype):
        wp = cache.get((key, "rowsum", x.dtype), weight,
                       lambda out=None: packed_weight_rowsum(weight, x.c, out=out, dtype=x.dtype, w_scale=w_scale))
        d = L.ConvDesc()
        d.x = x.ct
        d.cout, d.k, d.stride, d.pad, d.act = cout, k, 1, pad, act
        d.w_packed = wp.data_ptr()
        d.w_scale = w_scale.data_ptr() if w_scale is not None else None
        d.bias = bias.data_ptr() if bias is not None else None
        d.alpha = alpha.data_ptr() if alpha is not None else None
        d.out_nchw = out_nchw.data_ptr()
        d.residual_nchw = residual_nchw.data_ptr() if residual_nchw is not None else None
        d.aux_nchw = aux_nchw.data_ptr() if aux_nchw is not None else None
        ev = _Counters.conv_events
        if ev is not None:
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
        L.check(L.load().uegan_conv2d_fprop_rowsum(C.byref(d), _stream()), "conv2d_fprop_rowsum")
        if ev is not None:
            s1.record()
            ho, wo = x.h + 2 * pad - k + 1, x.w + 2 * pad - k + 1
            ev.append((s0, s1, 2.0 * x.n * ho * wo * cout * k * k * x.c, x, cout, k, 1, "fprop", x.dtype))
        _count(1, f"fprop_rowsum ->{cout} k{k}", x)
        return
    wp = cache.get((key, x.dtype), weight, lambda out=None: packed_weight(weight, x.c, x.dtype, out=out, w_scale=w_scale))
    conv_fprop(x, wp, cout, k, 1, pad, None, 0, bias, alpha, act, None, out_nchw, residual_nchw, aux_nchw=aux_nchw,
               w_scale=w_scale)


def pack_input(x_nchw: torch.Tensor, dst: NHWC, pad_mode: int = L.PAD_REFLECT, scale=None, shift=None):
    assert x_nchw.is_cuda and x_nchw.dtype == torch.float32 and x_nchw.is_contiguous() and x_nchw.shape[1] == 3
    assert tuple(x_nchw.shape) == (dst.n, 3, dst.h, dst.w)
    L.check(L.load().uegan_pack_input(x_nchw.data_ptr(), dst.ref(), pad_mode, L.float3(scale), L.float3(shift),
                                      _stream()), "pack_input")
    _count(1, "pack_input", dst)


def halo_fill(t: NHWC, pad_mode: int = L.PAD_REFLECT):
    L.check(L.load().uegan_halo_fill(t.ref(), pad_mode, _stream()), "halo_fill")
    _count(1, "halo_fill", t)


def instance_norm(src: NHWC, dst: NHWC, dst_c_off: int, stats_ws: torch.Tensor, eps: float = 1e-5):
    assert stats_ws.dtype == torch.float64 and stats_ws.numel() >= 3 * src.n * src.c
    L.check(L.load().uegan_instance_norm(src.ref(), dst.ref(), dst_c_off, eps, stats_ws.data_ptr(), _stream()),
            "instance_norm")
    _count(3, "instance_norm", src)


def instance_norm_apply(src: NHWC, dst: NHWC, dst_c_off: int, stats_ws: torch.Tensor, eps: float = 1e-5):
    """InstanceNorm whose sums were accumulated by the producing conv's epilogue (conv_fprop(in_stats=...))."""
    assert stats_ws.dtype == torch.float64 and stats_ws.numel() >= 3 * src.n * src.c
    L.check(L.load().uegan_instance_norm_apply(src.ref(), dst.ref(), dst_c_off, eps, stats_ws.data_ptr(), _stream()),
            "instance_norm_apply")
    _count(2, "instance_norm_apply", src)


def fused_stats_ok(ho: int, wo: int, cout: Optional[int] = None) -> bool:
    """Whether the InstanceNorm statistics ride in the producing conv's epilogue (conv_fprop(in_stats=...)); needs every
    128-pixel tile inside one image.  Default: NO -- the epilogue's per-tile butterfly (~150 instructions per 16-column
    chunk) makes the launch epilogue-bound, and a separate strip-reduce pass over the stored tensor at 5-7 TB/s is cheaper:
    measured on B200 (r1h A/B, training step) all-fused 72.2 ms, none 70.5 ms, cout <= 32 only 70.6 ms; VGG fprop 8.25 ->
    5.89 ms.  UEGAN_FUSED_STATS = all | none | <max cout> selects the policy."""
    import os
    mode = os.environ.get("UEGAN_FUSED_STATS", "none")
    if mode == "none":
        return False
    limit = (1 << 30) if mode == "all" else int(mode)
    return ho * wo >= 128 and wo >= 8 and (cout is None or cout <= limit)


def spectral_sigma(w: torch.Tensor, u: torch.Tensor, v: torch.Tensor, train: bool, sigma_out: torch.Tensor,
                   ws: torch.Tensor):
    rows = w.shape[0]
    cols = w.numel() // rows
    assert ws.numel() >= rows + cols + 8 and sigma_out.numel() >= 2
    L.check(L.load().uegan_spectral_sigma(w.data_ptr(), u.data_ptr(), v.data_ptr(), rows, cols, int(train),
                                          sigma_out.data_ptr(), ws.data_ptr(), _stream()), "spectral_sigma")
    _count(4 if train else 2, "spectral_sigma")


def spectral_sigma_batch(ws_list, us, vs, train: bool, sigma_outs, scratches, u_used=None, v_used=None):
    """spectral_sigma for several independent layers in one launch per phase (4 launches, or 2 in eval mode).
    u_used / v_used: optional per-layer fp32 buffers that receive the u / v this pass ends with."""
    n = len(ws_list)
    rows = (C.c_int32 * n)(*[w.shape[0] for w in ws_list])
    cols = (C.c_int32 * n)(*[w.numel() // w.shape[0] for w in ws_list])
    for w, sg, sc in zip(ws_list, sigma_outs, scratches):
        assert sc.numel() >= w.shape[0] + w.numel() // w.shape[0] + 8 and sg.numel() >= 2
    L.check(L.load().uegan_spectral_sigma_batch(n, _ptr_array(ws_list), _ptr_array(us), _ptr_array(vs), rows, cols, int(train),
                                                _ptr_array(sigma_outs), _ptr_array(scratches),
                                                _ptr_array(u_used) if u_used is not None else None,
                                                _ptr_array(v_used) if v_used is not None else None, _stream()),
            "spectral_sigma_batch")
    _count(4 if train else 2, "spectral_sigma_batch")


def upsample2x(src: NHWC, dst: NHWC, dst_c_off: int = 0):
    L.check(L.load().uegan_upsample2x(src.ref(), dst.ref(), dst_c_off, _stream()), "upsample2x")
    _count(1, "upsample2x", src)


def cat_build(u: NHWC, z: NHWC, stats_ws: torch.Tensor, dst: NHWC, eps: float = 1e-5) -> int:
    """dst = [bilinear x2 of u | InstanceNorm(z)] in one pass over whole pixels; statistics of z are computed here.
    Returns the device address of z's (mean, rstd) pairs inside stats_ws (kept for the backward pass)."""
    mr = instance_norm_stats(z, stats_ws, eps=eps)
    L.check(L.load().uegan_cat_build(u.ref(), z.ref(), mr, dst.ref(), _stream()), "cat_build")
    _count(1, "cat_build", dst)
    return mr


def cat_build_ok() -> bool:
    """Opt-in (UEGAN_CAT_BUILD=1): measured r3p the one-pass concat is SLOWER than the two half-line passes (inference 6031 ->
    5522 img/s, training 40.6 -> 41.3 ms): its warps diverge between the 4-load bilinear half and the 1-load normalise half."""
    import os
    return os.environ.get("UEGAN_CAT_BUILD") == "1"


def maxpool2x2(src: NHWC, dst: NHWC):
    L.check(L.load().uegan_maxpool2x2(src.ref(), dst.ref(), _stream()), "maxpool2x2")
    _count(1, "maxpool2x2", src)


def unpack_nchw(src: NHWC, c_off: int, c_count: int) -> torch.Tensor:
    out = torch.empty(src.n, c_count, src.h, src.w, dtype=torch.float32, device=src.buf.device)
    L.check(L.load().uegan_unpack_nchw(src.ref(), c_off, c_count, out.data_ptr(), _stream()), "unpack_nchw")
    return out


def zero_(t: torch.Tensor):
    """cudaMemsetAsync on the current stream (a memset node in a captured graph, not an ATen fill kernel)."""
    L.check(L.load().uegan_memset_zero(t.data_ptr(), t.numel() * t.element_size(), _stream()), "memset_zero")


def device_error() -> int:
    """Synchronises and returns the watchdog word (0 = no bounded wait expired)."""
    return L.load().uegan_device_error()


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
GAN_MODES = {"rahinge": 0, "rals": 1}
REC_TYPES = {"l1": 0, "smoothl1": 1, "l2": 2}


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def gan_loss_fwd(mode: int, for_d: bool, real, fake, ws: torch.Tensor, loss_out: torch.Tensor, group=None):
    """real / fake: lists of contiguous fp32 CUDA prediction maps.  ws: >= 48 float64 (kept for gan_loss_bwd).
    With a torch.distributed `group` the relativistic means and the normalisation run over the global batch: two
    all-reduces of 16 / 32 doubles between the three phases."""
    assert len(real) == len(fake) and ws.dtype == torch.float64 and ws.numel() >= 48
    for t in list(real) + list(fake):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    counts = (C.c_int64 * len(real))(*[t.numel() for t in real])
    for r, f in zip(real, fake):
        assert r.numel() == f.numel()
    lib = L.load()
    if group is None:
        L.check(lib.uegan_gan_loss_fwd(mode, int(for_d), len(real), _ptr_array(real), _ptr_array(fake), counts,
                                       ws.data_ptr(), loss_out.data_ptr(), _stream()), "gan_loss_fwd")
        _count(3, "gan_loss_fwd")
        return 1
    rp, fp = _ptr_array(real), _ptr_array(fake)
    if hasattr(group, "peer_ptrs"):  # uegan_b200.peer.PeerComm: sums over peer memory, no NCCL call
        world = group.world
        red = lambda lo, hi: group.reduce(ws, lo, hi)
    else:
        import torch.distributed as dist
        world = dist.get_world_size(group)
        red = lambda lo, hi: dist.all_reduce(ws[lo:hi], group=group)
    L.check(lib.uegan_gan_loss_phase(0, mode, int(for_d), len(real), rp, fp, counts, world, ws.data_ptr(), None,
                                     _stream()), "gan_loss_phase0")
    red(0, 16)
    L.check(lib.uegan_gan_loss_phase(1, mode, int(for_d), len(real), rp, fp, counts, world, ws.data_ptr(), None,
                                     _stream()), "gan_loss_phase1")
    red(16, 48)
    L.check(lib.uegan_gan_loss_phase(2, mode, int(for_d), len(real), rp, fp, counts, world, ws.data_ptr(),
                                     loss_out.data_ptr(), _stream()), "gan_loss_phase2")
    _count(3)
    return world


def gan_loss_bwd(mode: int, for_d: bool, real, fake, ws: torch.Tensor, d_real, d_fake, gscale_dev=None,
                 world: int = 1):
    counts = (C.c_int64 * len(real))(*[t.numel() for t in real])
    L.check(L.load().uegan_gan_loss_bwd(mode, int(for_d), len(real), _ptr_array(real), _ptr_array(fake), counts,
                                        ws.data_ptr(), _ptr_array(d_real) if d_real is not None else None,
                                        _ptr_array(d_fake) if d_fake is not None else None,
                                        gscale_dev.data_ptr() if gscale_dev is not None else None,
                                        -float(world) if world > 1 else 1.0, _stream()), "gan_loss_bwd")
    _count(1, "gan_loss_bwd")


def instance_norm_stats(src: NHWC, stats_ws: torch.Tensor, sums_ready: bool = False, eps: float = 1e-5) -> int:
    """Returns the device address of the per-(n,c) (mean, rstd) float pairs inside stats_ws."""
    assert stats_ws.dtype == torch.float64 and stats_ws.numel() >= 3 * src.n * src.c
    out = C.c_void_p()
    L.check(L.load().uegan_instance_norm_stats(src.ref(), eps, stats_ws.data_ptr(), int(sums_ready), C.byref(out),
                                               _stream()), "instance_norm_stats")
    _count(1 if sums_ready else 2, "instance_norm_stats", src)
    return out.value


def in_mse_fwd(x: NHWC, y: NHWC, mr_x: int, mr_y: int, weight: float, accum: torch.Tensor, loss: torch.Tensor):
    L.check(L.load().uegan_in_mse_fwd(x.ref(), y.ref(), mr_x, mr_y, float(weight), accum.data_ptr(), loss.data_ptr(),
                                      _stream()), "in_mse_fwd")
    _count(2, "in_mse_fwd", x)


def in_mse_joint(x: NHWC, y: NHWC, eps: float, weight: float, ws: torch.Tensor, accum: torch.Tensor, loss: torch.Tensor):
    """One PerceptualLoss term from ONE pass over the taps x, y (InstanceNorm statistics of both + MSE + backward sums).
    Returns the device addresses (mean/rstd of x, mean/rstd of y, backward sums) inside ws (>= 9 * n * c doubles)."""
    assert ws.dtype == torch.float64 and ws.numel() >= 9 * x.n * x.c
    mx, my, sm = C.c_void_p(), C.c_void_p(), C.c_void_p()
    L.check(L.load().uegan_in_mse_joint(x.ref(), y.ref(), float(eps), float(weight), ws.data_ptr(), accum.data_ptr(),
                                        loss.data_ptr(), C.byref(mx), C.byref(my), C.byref(sm), _stream()), "in_mse_joint")
    _count(3, "in_mse_joint", x)
    return mx.value, my.value, sm.value


def in_mse_bwd_apply(x: NHWC, y: NHWC, mr_x: int, mr_y: int, weight: float, gscale, deep: Optional[NHWC], dx: NHWC,
                     sums: int):
    L.check(L.load().uegan_in_mse_bwd_apply(x.ref(), y.ref(), mr_x, mr_y, float(weight),
                                            gscale.data_ptr() if gscale is not None else None,
                                            deep.ref() if deep is not None else None, dx.ref(), sums, _stream()),
            "in_mse_bwd_apply")
    _count(1, f"in_mse_bwd_apply{' +deep' if deep is not None else ''}", x)


def msrec_loss(pred: torch.Tensor, gt: torch.Tensor, rec_type: int, scales: int, accum: torch.Tensor,
               loss: torch.Tensor, grad: Optional[torch.Tensor] = None, grad_scale: float = 1.0, gscale_dev=None):
    assert pred.shape == gt.shape and pred.is_cuda and pred.dtype == torch.float32
    assert pred.is_contiguous() and gt.is_contiguous()
    n, c, h, w = pred.shape
    L.check(L.load().uegan_msrec_loss(pred.data_ptr(), gt.data_ptr(), n, c, h, w, rec_type, scales, accum.data_ptr(),
                                      loss.data_ptr(), grad.data_ptr() if grad is not None else None,
                                      float(grad_scale), gscale_dev.data_ptr() if gscale_dev is not None else None,
                                      _stream()), "msrec_loss")
    _count(2, "msrec_loss")


# ------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------
def packed_weight_dgrad(weight: torch.Tensor, cout_stored: int, dtype: int, stride: int = 1, pi: int = 0, pj: int = 0,
                        cin_first: int = 0, cin: Optional[int] = None, out: Optional[torch.Tensor] = None,
                        w_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Operand of the data-gradient GEMM of a conv with `weight` (OIHW): see uegan_pack_conv_weight_dgrad."""
    lib = L.load()
    w = weight.detach()
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
    o, i_total, k, _ = w.shape
    if cin is None:
        cin = i_total - cin_first
    kq = (k + stride - 1) // stride
    nbytes = lib.uegan_packed_weight_bytes(cin, cout_stored, kq, dtype)
    buf = out if out is not None else torch.empty(nbytes + 256, dtype=torch.uint8, device=w.device)
    assert buf.numel() >= nbytes
    if w_scale is not None:
        L.check(lib.uegan_pack_conv_weight_dgrad_scaled(w.data_ptr(), buf.data_ptr(), o, i_total, cin_first, cin,
                                                        cout_stored, k, stride, pi, pj, dtype, w_scale.data_ptr(),
                                                        _stream()), "pack_conv_weight_dgrad_scaled")
    else:
        L.check(lib.uegan_pack_conv_weight_dgrad(w.data_ptr(), buf.data_ptr(), o, i_total, cin_first, cin, cout_stored, k,
                                                 stride, pi, pj, dtype, _stream()), "pack_conv_weight_dgrad")
    _count(1, "pack_weight_dgrad")
    return buf


def conv_generic(x: NHWC, w_packed: torch.Tensor, cout: int, k: int, stride: int, pad: int, y: NHWC, y_c_off: int = 0,
                 bias=None, alpha=None, act: int = L.ACT_NONE, mask: Optional[NHWC] = None, mask_act: int = L.ACT_NONE,
                 y_mul: int = 1, y_off_h: int = 0, y_off_w: int = 0, real_taps: Optional[int] = None,
                 w_scale: Optional[torch.Tensor] = None, y_cls_c: int = 0):
    """conv_fprop with the dgrad-only options (activation-derivative mask, strided output view)."""
    lib = L.load()
    d = L.ConvDesc()
    d.x, d.y = x.ct, y.ct
    d.y_c_off, d.cout, d.k, d.stride, d.pad, d.act = y_c_off, cout, k, stride, pad, act
    d.w_packed = w_packed.data_ptr()
    d.w_scale = w_scale.data_ptr() if w_scale is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.alpha = alpha.data_ptr() if alpha is not None else None
    d.mask = C.pointer(mask.ct) if mask is not None else None
    d.mask_act = mask_act
    d.y_mul, d.y_off_h, d.y_off_w, d.y_cls_c = y_mul, y_off_h, y_off_w, y_cls_c
    ev = _Counters.conv_events
    if ev is not None:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
    L.check(lib.uegan_conv2d_fprop(C.byref(d), _stream()), "conv2d (dgrad)")
    if ev is not None:
        s1.record()
        # algorithmic MACs of a data-gradient launch = its share of the forward conv's MACs (real taps only; a merged
        # launch covers all four parity classes: real_taps = the forward kernel's k*k, channels = one class)
        taps = real_taps if real_taps is not None else k * k
        cc = y_cls_c if y_cls_c else cout
        real_c = 3 if cc == 16 and y.c == 16 else cc
        ev.append((s0, s1, 2.0 * x.n * x.h * x.w * x.c * real_c * taps, x, cc, k, stride, "dgrad", x.dtype))
    _count(1, f"dgrad ->{cout} k{k} ymul{y_mul}{' x4cls' if y_cls_c else ''}{' +mask' if mask is not None else ''}", x)


def conv_dgrad(dz: NHWC, weight: torch.Tensor, k: int, stride: int, dxp: NHWC, cache=None, key=None, alpha=None,
               cin_first: int = 0, cin: Optional[int] = None, mask: Optional[NHWC] = None, mask_act: int = L.ACT_NONE,
               w_scale: Optional[torch.Tensor] = None):
    """Data gradient of y = conv(xpad, weight, stride) w.r.t. the PADDED input: dxp (extent of xpad, halo 0).
    dz: output gradient with a zero halo of ceil(k/stride) - 1.  stride 2 = four parity-class launches."""
    import os
    kq = (k + stride - 1) // stride
    assert dz.halo >= kq - 1, "dz needs a zero halo of ceil(k/stride)-1"
    cin_n = (weight.shape[1] - cin_first) if cin is None else cin
    cout_arg = (cin_n + 15) // 16 * 16  # RGB input (3 channels): the packed operand's rows 3..15 are zero
    assert dxp.c >= cout_arg
    if stride == 2 and mask is None and os.environ.get("UEGAN_NO_DGRAD_MERGE") != "1":
        # all four parity classes in ONE launch: N = 4 * cout_arg columns (class, channel) over the shared kq x kq window
        def fn4(out=None):
            lib = L.load()
            cb = lib.uegan_packed_weight_bytes(cin_n, dz.c, kq, dz.dtype)
            buf = out if out is not None else torch.empty(4 * cb + 256, dtype=torch.uint8, device=weight.device)
            wd = weight.detach()
            assert wd.is_cuda and wd.dtype == torch.float32 and wd.is_contiguous()
            L.check(lib.uegan_pack_conv_weight_dgrad4(wd.data_ptr(), buf.data_ptr(), wd.shape[0], wd.shape[1], cin_first, cin_n,
                                                      dz.c, k, dz.dtype, w_scale.data_ptr() if w_scale is not None else None,
                                                      _stream()), "pack_conv_weight_dgrad4")
            _count(1, "pack_weight_dgrad4")
            return buf
        wp = cache.get((key, "dg4", dz.dtype, dz.c), weight, fn4) if cache is not None else fn4()
        conv_generic(dz, wp, 4 * cout_arg, kq, 1, kq - 1, dxp, 0, None, alpha, L.ACT_NONE, None, L.ACT_NONE, y_mul=2,
                     real_taps=k * k, w_scale=w_scale, y_cls_c=cout_arg)
        return
    for pi in range(stride):
        for pj in range(stride):
            fn = lambda out=None: packed_weight_dgrad(weight, dz.c, dz.dtype, stride, pi, pj, cin_first, cin, out=out,
                                                      w_scale=w_scale)
            wp = cache.get((key, "dg", pi, pj, dz.dtype, dz.c), weight, fn) if cache is not None else fn()
            nr = len(range(pi, k, stride)) * len(range(pj, k, stride))  # taps of the forward kernel in this class
            conv_generic(dz, wp, cout_arg, kq, 1, kq - 1, dxp, 0, None, alpha, L.ACT_NONE, mask, mask_act,
                         y_mul=stride, y_off_h=pi, y_off_w=pj, real_taps=nr, w_scale=w_scale)


class _WgradWs:
    """Split-K workspace of the weight-gradient kernels.  UEGAN_DETERMINISTIC=1 (default): every k-slice stores its partial
    plane here and a second kernel sums the planes in slice order -- bit-reproducible runs, as the reference asks of cuDNN
    (utils.py:154 cudnn.deterministic=True).  UEGAN_DETERMINISTIC=0: fp32 atomics, no workspace."""
    BYTES = 96 << 20
    buf = {}

    @classmethod
    def get(cls, device):
        import os
        if os.environ.get("UEGAN_DETERMINISTIC", "1") == "0":
            return None, 0
        key = str(device)
        b = cls.buf.get(key)
        if b is None:
            b = cls.buf[key] = torch.empty(cls.BYTES // 4, dtype=torch.float32, device=device)
        return b.data_ptr(), cls.BYTES


def zwin_ok(cout: int, x: NHWC, dz: NHWC, k: int, stride: int = 1) -> bool:
    """Whether the weight gradient of (x, dz) can read the stacked gradient as a sliding window over a ZERO-haloed dz."""
    import os
    return (os.environ.get("UEGAN_NO_ZWIN") != "1" and x.dtype == dz.dtype and dz.h == x.h and dz.w == x.w
            and bool(L.load().uegan_conv2d_wgrad_zwin_supported(cout, dz.c, dz.halo, x.c, k, stride, x.dtype)))


def conv_wgrad(x: NHWC, dz: NHWC, dw: torch.Tensor, k: int, stride: int, pad: int, cin_first: int = 0,
               cin: Optional[int] = None, alpha=None, scale: float = 1.0, dz_zero_halo: bool = False):
    """dw (OIHW fp32, pre-zeroed or accumulating) += scale * alpha * wgrad(x, dz).
    dz_zero_halo: the caller guarantees that dz's halo holds zeros (a dgrad operand); stride-1 layers with few output
    channels then read the horizontally stacked gradient as a sliding window over dz (uegan_conv2d_wgrad_zwin)."""
    cout, cin_total = dw.shape[0], dw.shape[1]
    cin_n = (cin_total - cin_first) if cin is None else cin
    assert dw.is_cuda and dw.dtype == torch.float32 and dw.is_contiguous()
    lib = L.load()
    zwin = dz_zero_halo and pad == (k - 1) // 2 and zwin_ok(cout, x, dz, k, stride)
    ev = _Counters.conv_events
    if ev is not None:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
    if zwin:
        L.check(lib.uegan_conv2d_wgrad_zwin(x.ref(), dz.ref(), cout, cin_n, cin_total, cin_first, k, pad, dw.data_ptr(),
                                            alpha.data_ptr() if alpha is not None else None, float(scale),
                                            *_WgradWs.get(dw.device), _stream()), "conv2d_wgrad_zwin")
    else:
        L.check(lib.uegan_conv2d_wgrad(x.ref(), dz.ref(), cout, cin_n, cin_total, cin_first, k, stride, pad,
                                       dw.data_ptr(), alpha.data_ptr() if alpha is not None else None, float(scale),
                                       *_WgradWs.get(dw.device), _stream()), "conv2d_wgrad")
    if ev is not None:
        s1.record()
        real_cin = 3 if x.c * (4 if x.dtype == L.F32 else 2) == 16 else cin_n
        ev.append((s0, s1, 2.0 * dz.n * dz.h * dz.w * cout * real_cin * k * k, x, cout, k, stride, "wgrad", x.dtype))
    _count(lib.uegan_wgrad_last_launches(), f"wgrad{'_zwin' if zwin else ''} cout{cout} cin{cin_n} k{k}s{stride}", x, dz)


def head_bwd(dout: torch.Tensor, out: torch.Tensor, x, mode: int, dz: NHWC):
    assert dout.is_contiguous() and out.is_contiguous() and dout.dtype == torch.float32
    L.check(L.load().uegan_head_bwd(dout.data_ptr(), out.data_ptr(), x.data_ptr() if x is not None else None,
                                    dout.shape[1], mode, dz.ref(), _stream()), "head_bwd")
    _count(1, "head_bwd", dz)


def grad_combine(dst: NHWC, channels: int, src_a: Optional[NHWC] = None, pad_a: int = 0, pad_mode_a: int = L.PAD_REFLECT,
                 add_b: Optional[NHWC] = None, add_c: Optional[NHWC] = None, mask: Optional[NHWC] = None,
                 act: int = L.ACT_NONE, mul: Optional[NHWC] = None, dst_c_off: int = 0, a_c_off: int = 0, b_c_off: int = 0,
                 c_c_off: int = 0, mask_c_off: int = 0, mul_c_off: int = 0, dst2: Optional[NHWC] = None,
                 mul2: Optional[NHWC] = None):
    """dst2 / mul2: optional second output (halo 0) = mul2 * (the sum before mask / mul), from the same pass."""
    r = lambda t: t.ref() if t is not None else None
    if dst2 is not None:
        L.check(L.load().uegan_grad_combine2(dst.ref(), dst_c_off, channels, r(src_a), a_c_off, pad_a, pad_mode_a, r(add_b),
                                             b_c_off, r(add_c), c_c_off, r(mask), mask_c_off, act, r(mul), mul_c_off,
                                             dst2.ref(), 0, mul2.ref(), 0, _stream()), "grad_combine2")
    else:
        L.check(L.load().uegan_grad_combine(dst.ref(), dst_c_off, channels, r(src_a), a_c_off, pad_a, pad_mode_a, r(add_b),
                                            b_c_off, r(add_c), c_c_off, r(mask), mask_c_off, act, r(mul), mul_c_off,
                                            _stream()), "grad_combine")
    _count(1, f"grad_combine c{channels} pad{pad_a if src_a is not None else '-'}"
              f"{' +b' if add_b is not None else ''}{' +c' if add_c is not None else ''}"
              f"{' mask' if mask is not None else ''}{' mul' if mul is not None else ''}", dst, src_a)


def fold_inplace(dxp: NHWC, pad: int) -> NHWC:
    """Reflect-pad adjoint in place on the padded gradient `dxp` (halo 0, extent (h+2p) x (w+2p)); returns the view
    (interior h x w, ZERO halo p) that downstream wgrad / dgrad / elementwise kernels consume."""
    v = dxp.as_haloed(pad)
    L.check(L.load().uegan_fold_inplace(v.ref(), _stream()), "fold_inplace")
    _count(1, "fold_inplace", v)
    return v


def dz_hstack(dz: NHWC, cout: int, k: int, e: NHWC):
    L.check(L.load().uegan_dz_hstack(dz.ref(), cout, k, e.ref(), _stream()), "dz_hstack")
    _count(1, "dz_hstack", e)


def conv_wgrad_hstack(x: NHWC, e: NHWC, dw: torch.Tensor, k: int, pad: int, alpha=None, scale: float = 1.0):
    """dw (OIHW fp32, pre-zeroed or accumulating) += wgrad from the stacked gradient e = dz_hstack(dz)."""
    cout, cin_total = dw.shape[0], dw.shape[1]
    assert dw.is_cuda and dw.dtype == torch.float32 and dw.is_contiguous()
    ev = _Counters.conv_events
    if ev is not None:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
    L.check(L.load().uegan_conv2d_wgrad_hstack(x.ref(), e.ref(), cout, cin_total, cin_total, 0, k, pad, dw.data_ptr(),
                                               alpha.data_ptr() if alpha is not None else None, float(scale),
                                               *_WgradWs.get(dw.device), _stream()),
            "conv2d_wgrad_hstack")
    if ev is not None:
        s1.record()
        ev.append((s0, s1, 2.0 * x.n * x.h * x.w * cout * cin_total * k * k, x, cout, k, 1, "wgrad", x.dtype))
    _count(L.load().uegan_wgrad_last_launches(), f"wgrad_hstack cout{cout} cin{cin_total} k{k}", x, e)


def hstack_ok(cout: int, cin_stored: int, k: int) -> bool:
    import os
    return (os.environ.get("UEGAN_NO_HSTACK") != "1" and cout in (1, 3) and k * cout <= 21 and k % 2 == 1
            and cin_stored % 32 == 0)


def channel_sum(src: NHWC, out: torch.Tensor, c_off: int = 0, channels: Optional[int] = None, accumulate: bool = False):
    """out[c] (+)= sum over pixels of src[.., c_off + c]; `channels` stored channels are reduced, the first out.numel()
    of them are written."""
    channels = out.numel() if channels is None else channels
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() <= channels
    L.check(L.load().uegan_channel_sum(src.ref(), c_off, channels, out.data_ptr(), out.numel(), int(accumulate),
                                       _stream()), "channel_sum")
    _count(1, f"channel_sum c{channels}", src)


def instance_norm_bwd(dout: NHWC, d_c_off: int, z: NHWC, mean_rstd: int, dz: NHWC, ws: torch.Tensor):
    assert ws.dtype == torch.float64 and ws.numel() >= 2 * z.n * z.c
    L.check(L.load().uegan_instance_norm_bwd(dout.ref(), d_c_off, z.ref(), mean_rstd, dz.ref(), ws.data_ptr(),
                                             _stream()), "instance_norm_bwd")
    _count(2, "instance_norm_bwd", z)


def upsample2x_bwd(dout: NHWC, d_c_off: int, dsrc: NHWC):
    L.check(L.load().uegan_upsample2x_bwd(dout.ref(), d_c_off, dsrc.ref(), _stream()), "upsample2x_bwd")
    _count(1, "upsample2x_bwd", dout)


def maxpool2x2_bwd(src: NHWC, dpool: NHWC, dsrc: NHWC):
    L.check(L.load().uegan_maxpool2x2_bwd(src.ref(), dpool.ref(), dsrc.ref(), _stream()), "maxpool2x2_bwd")
    _count(1, "maxpool2x2_bwd", src)


def in_mse_bwd(x: NHWC, y: NHWC, mr_x: int, mr_y: int, weight: float, gscale, deep: Optional[NHWC], dx: NHWC,
               ws: torch.Tensor):
    L.check(L.load().uegan_in_mse_bwd(x.ref(), y.ref(), mr_x, mr_y, float(weight),
                                      gscale.data_ptr() if gscale is not None else None,
                                      deep.ref() if deep is not None else None, dx.ref(), ws.data_ptr(), _stream()),
            "in_mse_bwd")
    _count(2, f"in_mse_bwd{' +deep' if deep is not None else ''}", x)


def unpack_input_grad(dx: NHWC, scale, out: torch.Tensor, skip=None):
    """skip = (dout, res, x) fp32 NCHW: adds the Generator's identity path clamp(res + x) (models.py:72)."""
    sk = [t.data_ptr() for t in skip] if skip is not None else [None, None, None]
    L.check(L.load().uegan_unpack_input_grad(dx.ref(), L.float3(scale), out.data_ptr(), *sk, _stream()),
            "unpack_input_grad")
    _count(1, "unpack_input_grad", dx)


def spectral_bwd(grad: torch.Tensor, w: torch.Tensor, u: torch.Tensor, v: torch.Tensor, sigma: torch.Tensor,
                 ws: torch.Tensor, accum: Optional[torch.Tensor] = None):
    """In place on `grad`, or (accum given) accum += result with `grad` left as the per-pass scratch."""
    rows = w.shape[0]
    L.check(L.load().uegan_spectral_bwd(grad.data_ptr(), w.data_ptr(), u.data_ptr(), v.data_ptr(), sigma.data_ptr(),
                                        rows, w.numel() // rows, ws.data_ptr(),
                                        accum.data_ptr() if accum is not None else None, _stream()), "spectral_bwd")
    _count(2, "spectral_bwd")
