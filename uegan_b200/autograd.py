"""Training path: torch.autograd.Function wrappers whose forward AND backward run on the sm_100a kernels.

torch.autograd is only the OUTER tape (it sequences `d_loss.backward()` / `g_loss.backward()` of trainer.py:96,117
and accumulates parameter `.grad`s); every tensor operation below is a call into libuegan_sm100.so:
  fprop   conv_fprop_kernel            dgrad  the same kernel on the transposed/rotated operand (4 parity launches
  wgrad   conv_wgrad_kernel (tcgen05, MN-major tf32)                                      for stride-2 convs)
plus the HBM-bound backward kernels of csrc/backward.cu.  Reference: autograd of models.py:44-74, 139-155 and
losses.py:22-36, 219-231, 348-377 (`aten::convolution_backward`, reflection_pad2d_backward, native_batch_norm_backward,
upsample_bilinear2d_backward, ...).

Dead parameters: GAM's attention branch (ga*.conv.0/.2), the second half of ga*.fuse.0.weight and ga*.fuse.0.bias only
shift the input of an InstanceNorm by a per-(n,c) constant; their true gradient is exactly zero (the
reference produces rounding noise of ~1e-5 there) and zeros are returned.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import kernels as K

F32 = L.F32


class _Scratch:
    """Gradient scratch buffers, reused across calls (backward passes run one at a time).  With a ScaleBook every buffer
    gets its own power-of-two scale slot (fp16 gradients); `share` makes a buffer use another buffer's slot."""

    def __init__(self, book=None):
        self.store = {}
        self.book = book

    def get(self, name, n, h, w, c, halo=0, dtype=F32, device="cuda", zero=False, share=None):
        """zero=True: allocated zeroed (a halo that only ever holds zeros: the buffer's producers write the interior only)."""
        key = (name, n, h, w, c, halo, dtype, str(device))
        t = self.store.get(key)
        if t is None:
            scale = None
            if self.book is not None and dtype != F32:
                scale = share.scale if share is not None else self.book.slot()
                zero = True
            t = K.NHWC(n, h, w, c, halo, dtype, device, zero=zero, scale=scale)
            if scale is not None and share is None:
                self.book.track_nhwc(t)
            self.store[key] = t
        return t


_scratch = _Scratch()
_FOLD_INPLACE = __import__("os").environ.get("UEGAN_NO_FOLD_INPLACE") != "1"


class _Side:
    """Weight gradients (+ their split-K reductions, bias sums, spectral-norm corrections) are off the critical path of a
    backward pass: only the optimizer reads them.  They run on ONE side stream, forked after the kernel that produced
    their gradient operand and joined at the end of the pass, so the data-gradient chain and its HBM-bound elementwise
    passes overlap with them (in a captured step: parallel branches of the CUDA graph).  Used only with a gradient sink
    (the trainer's flat buckets): every buffer involved is then persistent, so no allocation crosses streams.  All
    weight-gradient work shares the side stream, which also serialises the shared split-K workspace.
    UEGAN_SIDE_STREAM=0 runs everything on the current stream."""
    enabled = __import__("os").environ.get("UEGAN_SIDE_STREAM", "1") != "0"
    streams = {}

    @classmethod
    def get(cls, dev):
        key = str(dev)
        st = cls.streams.get(key)
        if st is None:
            st = cls.streams[key] = torch.cuda.Stream(device=dev)
        return st

    @classmethod
    def run(cls, dev, on, fn):
        if not (on and cls.enabled):
            fn()
            return
        side = cls.get(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fn()

    @classmethod
    def join(cls, dev, on):
        if on and cls.enabled:
            torch.cuda.current_stream(dev).wait_stream(cls.get(dev))


def _zeros_like(p):
    return torch.zeros_like(p, memory_format=torch.contiguous_format)


# =================================================================================================
# Generator
# =================================================================================================
def _g_workspace(G, b, h, w, dev):
    d = G.conv_dim
    dt = G._dtype
    book = K.ScaleBook(dev) if dt != F32 else None

    def T(hh, ww, c, halo=0, zero=False, scaled=True):
        if book is None:
            return K.NHWC(b, hh, ww, c, halo, F32, dev, zero)
        t = K.NHWC(b, hh, ww, c, halo, dt, dev, True, scale=book.slot() if scaled else None)
        if scaled:
            book.track_nhwc(t)
        return t
    f64 = lambda n: torch.empty(n, dtype=torch.float64, device=dev)
    chs = [8 * d, 4 * d, 2 * d, d]
    res = [(h // 8, w // 8), (h // 4, w // 4), (h // 2, w // 2), (h, w)]
    bwd_book = K.ScaleBook(dev) if dt != F32 else None
    return dict(
        book=book, bwd_book=bwd_book, scratch=_Scratch(bwd_book) if dt != F32 else None, dtype=dt,
        x0=T(h, w, 4 if dt == F32 else 8, 3, True, scaled=False), x1=T(h, w, d, 1), x2=T(h // 2, w // 2, 2 * d, 1), x3=T(h // 4, w // 4, 4 * d, 1),
        x4=T(h // 8, w // 8, 8 * d, 1), x5=T(h // 16, w // 16, 16 * d),
        z5=T(h // 16, w // 16, 16 * d), x5n=T(h // 16, w // 16, 16 * d), st5=f64(3 * b * 16 * d),
        u=[T(r[0] // 2, r[1] // 2, c) for r, c in zip(res, chs)],
        cat=[T(r[0], r[1], 2 * c, 1) for r, c in zip(res, chs)],
        z=[T(r[0], r[1], c) for r, c in zip(res, chs)],
        st=[f64(3 * b * c) for c in chs],
        y=[T(r[0], r[1], c) for r, c in zip(res, chs)],
        y4m=T(h, w, d, 1), t=T(h, w, d, 3),
        res=torch.empty(b, 3, h, w, dtype=torch.float32, device=dev),
        mr={},
    )


def _gam_forward(G, name, ga, src, ch, z, dst, off, stats, ws, up=None):
    """up: the low-resolution tensor whose bilinear x2 fills dst[.., 0:ch): the concat is then built in one pass."""
    fuse = ga.fuse[0]
    dt = G._dtype
    wsc = G._wscale(name, fuse)
    wp = G._wcache.get((name, dt), fuse.weight,
                       lambda out=None: K.packed_weight(fuse.weight, src.c, dt, 0, ch, out=out, w_scale=wsc))
    if up is not None:
        K.conv_fprop(src, wp, ch, 1, 1, 0, z, w_scale=wsc)
        ws["mr"][name] = K.cat_build(up, z, stats, dst)
        return
    if dt == F32 and K.fused_stats_ok(src.h, src.w, ch):
        K.conv_fprop(src, wp, ch, 1, 1, 0, z, in_stats=stats)
        K.instance_norm_apply(z, dst, off, stats)
        ws["mr"][name] = stats.data_ptr() + 2 * src.n * ch * 8
    else:
        K.conv_fprop(src, wp, ch, 1, 1, 0, z, w_scale=wsc)
        K.instance_norm(z, dst, off, stats)
        ws["mr"][name] = stats.data_ptr() + 2 * src.n * ch * 8


def _g_layers(G):
    d = G.conv_dim
    ups = [G.upsample1[1], G.upsample2[1], G.upsample3[1], G.upsample4[1]]
    gas = [G.ga4, G.ga3, G.ga2, G.ga1]
    decs = [G.dec1, G.dec2, G.dec3, G.dec4]
    return d, ups, gas, decs


def _settle(book, run, max_iter=48):
    """First use of an fp16 workspace: repeat `run()` + scale update until no per-tensor scale changes any more (every
    pass fixes at least the next layer of the chain; host-synchronising, one-off, happens before any graph capture)."""
    prev = None
    for _ in range(max_iter):
        run()
        book.update()
        cur = book.values()
        if prev is not None and torch.equal(cur, prev):
            return
        prev = cur


def _g_forward_train(G, x, ws):
    if ws["book"] is not None:
        if not ws.get("fwd_settled"):
            ws["fwd_settled"] = True
            _settle(ws["book"], lambda: _g_forward_pass(G, x, ws))
        else:
            ws["book"].update()  # delayed scaling: this pass stores with the magnitudes the previous pass measured
    return _g_forward_pass(G, x, ws)


def _g_forward_pass(G, x, ws):
    d, ups, gas, decs = _g_layers(G)
    act = G._act
    P = ws
    if G._dtype != F32:
        for name, holder in G._conv_table():
            G._wscale(name, holder)
        G._update_weight_scales()

    def conv(src, name, holder, cout, k, stride, dst, act_=L.ACT_NONE, mul=None, premul=None, halo=False):
        """halo: dst's reflection-padding halo is written too (by the epilogue, or by a halo_fill launch after it)."""
        cv = holder.conv
        K.conv_fprop(src, G._w(name, cv, src.c), cout, k, stride, (k - 1) // 2, dst, 0, cv.bias, None, act_, mul=mul,
                     w_scale=G._wscale(name, cv), premul=premul, reflect_halo=halo)

    # opt-in (UEGAN_PREMUL=1): measured r3e it moves the multiply's 0.4 ms into dec4's epilogue-bound launch, no net gain
    f16_premul = G._dtype != F32 and __import__("os").environ.get("UEGAN_PREMUL") == "1"
    K.pack_input(x, P["x0"], L.PAD_REFLECT)
    conv(P["x0"], "enc1", G.enc1, d, 7, 1, P["x1"], act, halo=True)
    conv(P["x1"], "enc2", G.enc2, 2 * d, 3, 2, P["x2"], act, halo=True)
    conv(P["x2"], "enc3", G.enc3, 4 * d, 3, 2, P["x3"], act, halo=True)
    conv(P["x3"], "enc4", G.enc4, 8 * d, 3, 2, P["x4"], act, halo=True)
    conv(P["x4"], "enc5", G.enc5, 16 * d, 3, 2, P["x5"], act)
    _gam_forward(G, "ga5", G.ga5, P["x5"], 16 * d, P["z5"], P["x5n"], 0, P["st5"], ws)
    src = P["x5n"]
    skips = [P["x4"], P["x3"], P["x2"], P["x1"]]
    for i in range(4):
        ch = P["u"][i].c
        conv(src, f"upsample{i+1}", ups[i], ch, 1, 1, P["u"][i])
        if K.cat_build_ok():  # [bilinear x2 of u | IN(conv1x1(skip))] written as whole pixels by one kernel
            _gam_forward(G, f"ga{4-i}", gas[i], skips[i], ch, P["z"][i], P["cat"][i], ch, P["st"][i], ws, up=P["u"][i])
        else:
            K.upsample2x(P["u"][i], P["cat"][i], 0)
            _gam_forward(G, f"ga{4-i}", gas[i], skips[i], ch, P["z"][i], P["cat"][i], ch, P["st"][i], ws)
        K.halo_fill(P["cat"][i])
        if i == 3 and f16_premul:
            # y4.mul(x1), models.py:70, in dec4's epilogue; y4 itself (kept for backward) is its second output
            conv(P["cat"][i], f"dec{i+1}", decs[i], ch, 3, 1, P["y4m"], act, mul=P["x1"], premul=P["y"][i])
        else:
            conv(P["cat"][i], f"dec{i+1}", decs[i], ch, 3, 1, P["y"][i], act)
        src = P["y"][i]
    if not f16_premul:
        K.grad_combine(P["y4m"], d, add_b=P["y"][3], mul=P["x1"])  # y4.mul(x1) (y4 itself is kept for backward)
    K.halo_fill(P["y4m"])
    conv(P["y4m"], "dec5.0", G.dec5[0], d, 3, 1, P["t"], halo=True)
    out = torch.empty_like(x)
    cv = G.dec5[1].conv
    K.conv_planar(P["t"], cv.weight, G._wcache, "dec5.1", 7, 3, cv.bias, None, L.ACT_TANH, out, x, aux_nchw=P["res"],
                  w_scale=G._wscale("dec5.1", cv))
    return out


def _g_backward(G, x, out_grad, ws, need_dx):
    if ws.get("bwd_book") is not None:
        if not ws.get("bwd_settled"):
            ws["bwd_settled"] = True
            _settle(ws["bwd_book"], lambda: _g_backward_pass(G, x, out_grad, ws, need_dx, dry=True))
        else:
            ws["bwd_book"].update()
    return _g_backward_pass(G, x, out_grad, ws, need_dx)


def _g_backward_pass(G, x, out_grad, ws, need_dx, dry=False):
    """dry=True: the data-gradient chain only (settling the gradient scales of a new fp16 workspace): no weight / bias
    gradient is accumulated."""
    d, ups, gas, decs = _g_layers(G)
    P = ws
    b, _, h, w = x.shape
    dev = x.device
    dty = ws.get("dtype", F32)
    f16 = dty != F32
    scr = ws["scratch"] if f16 else _scratch
    S = lambda name, hh, ww, c, halo=0, share=None, zero=False: scr.get(name, b, hh, ww, c, halo, dty, dev, share=share,
                                                                        zero=zero)
    rgb_c = 8 if f16 else 4  # stored channels of a 3-channel tensor: one 16-byte vector
    grads = {}
    cache = G._wcache
    act = G._act
    # gradient sink (uegan_b200.optim.FlatBucket): the weight-gradient kernels accumulate straight into the optimizer's
    # flat bucket and autograd gets None for those parameters; without a sink (plain autograd use) fresh tensors are returned
    sink = getattr(G, "_grad_sink", None)

    def gbuf(pname, param, zero=True):
        if sink is not None:
            return sink[pname], True
        return (_zeros_like(param) if zero else torch.empty_like(param)), False

    side = sink is not None and not dry

    def wgrad(name, conv, xin, dz, k, stride, pad, cin_first=0, cin=None, bias_from=None, zero_halo=False):
        """zero_halo: dz's halo holds zeros (every dgrad operand here), which the sliding-window stack path needs."""
        if dry:
            return
        gw, direct = gbuf(name + ".weight", conv.weight)
        grads[name + ".weight"] = None if direct else gw
        gb = None
        if bias_from is not None and conv.bias is not None:
            gb, direct_b = gbuf(name + ".bias", conv.bias, zero=False)
            grads[name + ".bias"] = None if direct_b else gb

        def run():
            K.conv_wgrad(xin, dz, gw, k, stride, pad, cin_first=cin_first, cin=cin, dz_zero_halo=zero_halo)
            if gb is not None:
                K.channel_sum(bias_from, gb, accumulate=direct_b)
        _Side.run(dev, side, run)

    def dead(pname, param):
        grads[pname] = None if sink is not None else torch.zeros_like(param)

    # ---- dec5.1 + tanh + clamp(res + x)   (models.py:34-35, 72)
    dz5 = S("dz5", h, w, rgb_c, 6)
    K.head_bwd(out_grad, P["res"], x, 2, dz5)
    c51 = G.dec5[1].conv
    if dry:
        pass
    elif K.zwin_ok(3, P["t"], dz5, 7):
        # the horizontal taps are a sliding window over the zero-haloed dz5 itself: only the 7 vertical taps are enumerated
        gw, direct = gbuf("dec5.1.main.1.weight", c51.weight)
        _Side.run(dev, side, lambda: K.conv_wgrad(P["t"], dz5, gw, 7, 1, 3, dz_zero_halo=True))
        grads["dec5.1.main.1.weight"] = None if direct else gw
    elif K.hstack_ok(3, d, 7):
        # horizontal taps unrolled into the gradient's channels: the wgrad keeps only the 7 vertical taps
        e5 = S("dz5e", h, w + 6, 32, 0, share=dz5)
        K.dz_hstack(dz5, 3, 7, e5)
        gw, direct = gbuf("dec5.1.main.1.weight", c51.weight)
        _Side.run(dev, side, lambda: K.conv_wgrad_hstack(P["t"], e5, gw, 7, 3))
        grads["dec5.1.main.1.weight"] = None if direct else gw
    else:
        dz5w = S("dz5w", h, w, 32, 0, share=dz5)  # the same gradient with 32 stored channels (wgrad operand)
        K.head_bwd(out_grad, P["res"], x, 2, dz5w)
        wgrad("dec5.1.main.1", c51, P["t"], dz5w, 7, 1, 3)
    if not dry:
        gb51, direct51 = gbuf("dec5.1.main.1.bias", c51.bias, zero=False)
        # one stored vector reduced, the 3 real channels written
        _Side.run(dev, side, lambda: K.channel_sum(dz5, gb51, 0, rgb_c, accumulate=direct51))
        grads["dec5.1.main.1.bias"] = None if direct51 else gb51
    wsc = lambda name, conv: G._wscale(name, conv)
    dxp = S("dxp_t", h + 6, w + 6, d)
    K.conv_dgrad(dz5, c51.weight, 7, 1, dxp, cache, "dec5.1", w_scale=wsc("dec5.1", c51))
    if _FOLD_INPLACE:
        dt = K.fold_inplace(dxp, 3)  # the padded gradient itself becomes dt (interior + zero halo 3)
    else:
        dt = S("dt", h, w, d, 2)
        K.grad_combine(dt, d, src_a=dxp, pad_a=3)
    # ---- dec5.0 (no activation)
    c50 = G.dec5[0].conv
    wgrad("dec5.0.main.1", c50, P["y4m"], dt, 3, 1, 1, bias_from=dt, zero_halo=True)
    dxp2 = S("dxp_y4m", h + 2, w + 2, d)
    K.conv_dgrad(dt, c50.weight, 3, 1, dxp2, cache, "dec5.0", w_scale=wsc("dec5.0", c50))
    # y4m = y4 * x1 ;  y4 = act(z4)
    dz = S("dz_dec3", h, w, d, 2)
    dx1a = S("dx1a", h, w, d)
    K.grad_combine(dz, d, src_a=dxp2, pad_a=1, mul=P["x1"], mask=P["y"][3], act=act, dst2=dx1a, mul2=P["y"][3])
    # ---- decoder stages 4..1
    skips = [P["x4"], P["x3"], P["x2"], P["x1"]]
    dskip = [None] * 4
    srcs = [P["x5n"], P["y"][0], P["y"][1], P["y"][2]]
    d_x5n = None
    for i in (3, 2, 1, 0):
        ch = P["u"][i].c
        hh, ww = P["y"][i].h, P["y"][i].w
        dec, up, ga = decs[i].conv, ups[i].conv, gas[i].fuse[0]
        wgrad(f"dec{i+1}.main.1", dec, P["cat"][i], dz, 3, 1, 1, bias_from=dz, zero_halo=True)
        dxpc = S(f"dxp_cat{i}", hh + 2, ww + 2, 2 * ch)
        K.conv_dgrad(dz, dec.weight, 3, 1, dxpc, cache, f"dec{i+1}", w_scale=wsc(f"dec{i+1}", dec))
        if _FOLD_INPLACE:
            dcat = K.fold_inplace(dxpc, 1)
        else:
            dcat = S(f"dcat{i}", hh, ww, 2 * ch)
            K.grad_combine(dcat, 2 * ch, src_a=dxpc, pad_a=1)
        # first half: bilinear x2 of the (hoisted) 1x1 conv
        du = S(f"du{i}", hh // 2, ww // 2, ch)
        K.upsample2x_bwd(dcat, 0, du)
        wgrad(f"upsample{i+1}.1.main.1", up, srcs[i], du, 1, 1, 0, bias_from=du)
        if i > 0:
            # the data gradient of the 1x1 up-conv IS the gradient of y_{i-1} = act(z): the LeakyReLU mask rides in the
            # dgrad epilogue and the result lands in the (zero-haloed) operand of the next decoder stage directly
            dz_next = S(f"dz_dec{i-1}", hh // 2, ww // 2, 2 * ch, 2, zero=True)
            K.conv_dgrad(du, up.weight, 1, 1, dz_next, cache, f"upsample{i+1}", w_scale=wsc(f"upsample{i+1}", up),
                         mask=P["y"][i - 1], mask_act=act)
        else:
            dsrc = S(f"dsrc{i}", hh // 2, ww // 2, 2 * ch)
            K.conv_dgrad(du, up.weight, 1, 1, dsrc, cache, f"upsample{i+1}", w_scale=wsc(f"upsample{i+1}", up))
        # second half: InstanceNorm(conv1x1(skip, fuse.weight[:, :ch]))
        dzs = S(f"dzs{i}", hh, ww, ch)
        K.instance_norm_bwd(dcat, ch, P["z"][i], P["mr"][f"ga{4-i}"], dzs,
                            torch.empty(2 * b * ch, dtype=torch.float64, device=dev))
        gname = f"ga{4-i}"
        wgrad(gname + ".fuse.0", ga, skips[i], dzs, 1, 1, 0, cin_first=0, cin=ch)
        dead(gname + ".fuse.0.bias", ga.bias)
        dsk = S(f"dskip{i}", hh, ww, ch)
        K.conv_dgrad(dzs, ga.weight, 1, 1, dsk, cache, gname, cin_first=0, cin=ch, w_scale=wsc(gname, ga))
        dskip[i] = dsk
        if i > 0:
            dz = dz_next
        else:
            d_x5n = dsrc
    # ---- ga5 on x5
    c5 = 16 * d
    h5, w5 = P["x5"].h, P["x5"].w
    dz5g = S("dz5g", h5, w5, c5)
    K.instance_norm_bwd(d_x5n, 0, P["z5"], P["mr"]["ga5"], dz5g, torch.empty(2 * b * c5, dtype=torch.float64, device=dev))
    f5 = G.ga5.fuse[0]
    wgrad("ga5.fuse.0", f5, P["x5"], dz5g, 1, 1, 0, cin_first=0, cin=c5)
    dead("ga5.fuse.0.bias", f5.bias)
    dze = S("dz_e5", h5, w5, c5, 1, zero=True)  # = mask(x5) * dgrad, straight from the epilogue
    K.conv_dgrad(dz5g, f5.weight, 1, 1, dze, cache, "ga5", cin_first=0, cin=c5, w_scale=wsc("ga5", f5),
                 mask=P["x5"], mask_act=act)
    # ---- encoder 5..1
    encs = [G.enc1, G.enc2, G.enc3, G.enc4, G.enc5]
    xs = [P["x0"], P["x1"], P["x2"], P["x3"], P["x4"], P["x5"]]
    for li in (5, 4, 3, 2):  # enc{li}: x_{li-1} -> x_li, k3 s2
        conv = encs[li - 1].conv
        xin = xs[li - 1]
        wgrad(f"enc{li}.main.1", conv, xin, dze, 3, 2, 1, bias_from=dze)
        dxpe = S(f"dxp_x{li-1}", xin.h + 2, xin.w + 2, xin.c)
        K.conv_dgrad(dze, conv.weight, 3, 2, dxpe, cache, f"enc{li}", w_scale=wsc(f"enc{li}", conv))
        halo = 1 if li > 2 else 3  # next dz feeds the dgrad of a k3s2 conv (halo 1); enc1's dz only feeds wgrad
        nxt = S(f"dz_e{li-1}", xin.h, xin.w, xin.c, halo if li > 2 else 0)
        K.grad_combine(nxt, xin.c, src_a=dxpe, pad_a=1, add_b=dskip[5 - li], add_c=dx1a if li == 2 else None,
                       mask=xin, act=act)
        dze = nxt
    wgrad("enc1.main.1", G.enc1.conv, P["x0"], dze, 7, 1, 3, bias_from=dze)
    dx = None
    if need_dx:
        dxp0 = S("dxp_x0", h + 6, w + 6, 16)
        dze3 = S("dz_e1h", h, w, d, 6)
        K.grad_combine(dze3, d, add_b=dze)
        K.conv_dgrad(dze3, G.enc1.conv.weight, 7, 1, dxp0, cache, "enc1", w_scale=wsc("enc1", G.enc1.conv))
        dx0 = S("dx0", h, w, 16)
        K.grad_combine(dx0, 16, src_a=dxp0, pad_a=3)
        dx = torch.empty_like(x)
        # + the identity path of out = clamp(res + x, -1, 1) (models.py:72)
        K.unpack_input_grad(dx0, None, dx, skip=(out_grad, P["res"], x))
    # dead parameters (see module docstring)
    for i in range(1, 6):
        ga = getattr(G, f"ga{i}")
        dead(f"ga{i}.conv.0.weight", ga.conv[0].weight)
        dead(f"ga{i}.conv.2.weight", ga.conv[2].weight)
    _Side.join(dev, side)
    return grads, dx


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        b, _, h, w = x.shape
        if module.conv_dim % 32:
            raise NotImplementedError("training kernels need conv_dim to be a multiple of 32 (wgrad operand rows)")
        x = x.detach().contiguous().float()
        key = (b, h, w, str(x.device), module.precision)
        pool = module._train_pool.setdefault(key, [])
        ws = pool.pop() if pool else _g_workspace(module, b, h, w, x.device)
        out = _g_forward_train(module, x, ws)
        ctx.module, ctx.ws, ctx.key, ctx.x = module, ws, key, x
        return out

    @staticmethod
    def backward(ctx, gout):
        G = ctx.module
        if ctx.ws is None:
            raise RuntimeError("uegan_b200: second backward through the same Generator forward (its activation "
                               "workspace has been recycled); run the forward again")
        grads, dx = _g_backward(G, ctx.x, gout.contiguous().float(), ctx.ws, ctx.needs_input_grad[1])
        G._train_pool[ctx.key].append(ctx.ws)
        ctx.ws = None
        names = [n for n, _ in G.named_parameters()]
        return (None, dx) + tuple(grads[n] for n in names)  # None where the kernels wrote into the gradient sink


def generator_apply(module, x):
    if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] % 16 or x.shape[3] % 16 or min(x.shape[2:]) < 32:
        raise ValueError("Generator needs (B,3,H,W) with H, W multiples of 16 and >= 32")
    if module._act is None:
        raise NotImplementedError("activation function [%s] has no sm_100a epilogue" % module.act_fun)
    if not hasattr(module, "_train_pool"):
        module._train_pool = {}
    return _GeneratorFn.apply(module, x, *[p for _, p in module.named_parameters()])


# =================================================================================================
# Discriminator
# =================================================================================================
def _d_workspace(D, b, h, w, dev):
    ws = D._act_buffers(b, h, w, dev)  # x0, ds[5], book, dtype
    dt = ws["dtype"]
    bwd_book = K.ScaleBook(dev) if dt != F32 else None
    ws.update(sig=[torch.ones(2, dtype=torch.float32, device=dev) for _ in range(5)], u=[], v=[], preds=[],
              bwd_book=bwd_book, scratch=_Scratch(bwd_book) if dt != F32 else None)
    return ws


def _d_forward_train(D, x, ws):
    if ws["book"] is not None:
        if not ws.get("fwd_settled"):
            ws["fwd_settled"] = True
            # The spectral-norm power iteration must advance exactly once, and the scales must settle on the sigma the
            # real pass uses (from a random u, sigma moves by a large factor in the first iteration, which compounds over
            # the five layers): advance u / v first, then settle AND run with u / v held (sigma = u.(W v) is then the very
            # value the train-mode pass would have computed).
            if D.training and D.use_sn:
                _d_advance_sn(D, x.device)
            _settle(ws["book"], lambda: _d_forward_pass(D, x, ws, False))
            return _d_forward_pass(D, x, ws, False)
        ws["book"].update()
    return _d_forward_pass(D, x, ws, D.training)


def _d_advance_sn(D, dev):
    """One power iteration of every spectrally-normalised conv (what a train-mode forward does, models.py:185-188)."""
    wl = [D._weight(i) for i in range(1, 6)]
    sigs = [torch.empty(2, dtype=torch.float32, device=dev) for _ in wl]
    scr = [torch.empty(w_.shape[0] + w_.numel() // w_.shape[0] + 8, dtype=torch.float32, device=dev) for w_ in wl]
    K.spectral_sigma_batch(wl, [D._conv(i).weight_u for i in range(1, 6)], [D._conv(i).weight_v for i in range(1, 6)],
                           True, sigs, scr)


def _d_forward_pass(D, x, ws, training):
    D._register_weight_scales()
    dt = D._dtype
    K.pack_input(x, ws["x0"], L.PAD_REFLECT)
    src, preds = ws["x0"], []
    if D.use_sn:
        # all five layers' power iterations in one launch per phase (the layers are independent); the kernels also leave
        # copies of the u / v this forward used in the workspace (later forwards move the module's buffers on)
        nl = len(D._SPEC)
        wl = [D._weight(i) for i in range(1, nl + 1)]
        if "sn_scratch" not in ws:
            f32 = lambda n: torch.empty(n, dtype=torch.float32, device=x.device)
            ws["sn_scratch"] = [f32(w_.shape[0] + w_.numel() // w_.shape[0] + 8) for w_ in wl]
            ws["u"] = [f32(w_.shape[0]) for w_ in wl]
            ws["v"] = [f32(w_.numel() // w_.shape[0]) for w_ in wl]
        K.spectral_sigma_batch(wl, [D._conv(i).weight_u for i in range(1, nl + 1)],
                               [D._conv(i).weight_v for i in range(1, nl + 1)], training, ws["sig"], ws["sn_scratch"],
                               ws["u"], ws["v"])
    for i, (k, pad) in enumerate(D._SPEC, start=1):
        conv, head, wgt = D._conv(i), D._head(i), D._weight(i)
        alpha = ws["sig"][i - 1][1:2] if D.use_sn else None
        dst = ws["ds"][i - 1]
        wsc = D._wscale(f"d{i}", wgt)
        wp = D._wcache.get((f"d{i}", dt), wgt, lambda out=None: K.packed_weight(wgt, src.c, dt, out=out, w_scale=wsc))
        K.conv_fprop(src, wp, wgt.shape[0], k, 2, pad, dst, 0, conv.bias, alpha, D._act, w_scale=wsc, reflect_halo=True)
        pred = torch.empty(x.shape[0], 1, dst.h, dst.w, dtype=torch.float32, device=x.device)
        K.conv_planar(dst, head.weight, D._wcache, f"p{i}", k, pad, None, None, D._head_act, pred,
                      w_scale=D._wscale(f"p{i}", head.weight))
        preds.append(pred)
        src = dst
    ws["preds"] = preds
    return preds


def _d_backward(D, x, dpreds, ws, need_dx, need_w=True):
    if ws.get("bwd_book") is not None:
        if not ws.get("bwd_settled"):
            ws["bwd_settled"] = True
            _settle(ws["bwd_book"], lambda: _d_backward_pass(D, x, dpreds, ws, need_dx, False))
        else:
            ws["bwd_book"].update()
    return _d_backward_pass(D, x, dpreds, ws, need_dx, need_w)


def _d_backward_pass(D, x, dpreds, ws, need_dx, need_w=True):
    b, _, h, w = x.shape
    dev = x.device
    dty = ws.get("dtype", F32)
    f16 = dty != F32
    scr = ws["scratch"] if f16 else _scratch
    S = lambda name, hh, ww, c, halo=0, share=None: scr.get("D" + name, b, hh, ww, c, halo, dty, dev, share=share)
    rgb_c = 8 if f16 else 4
    cache = D._wcache
    grads = {}
    sink = getattr(D, "_grad_sink", None) if need_w else None
    wsc = lambda i: D._wscale(f"d{i}", D._weight(i))
    wsp = lambda i: D._wscale(f"p{i}", D._head(i).weight)

    def gbuf(pname, param, zero=True):
        if sink is not None:
            return sink[pname], True
        return (_zeros_like(param) if zero else torch.empty_like(param)), False
    side = sink is not None  # weight-gradient work on the side stream (see _Side)
    srcs = [ws["x0"]] + ws["ds"]
    carry = None  # gradient w.r.t. the padded ds_k coming from d_{k+1}
    head_mode = 0 if D._head_act == L.ACT_TANH else 1
    for i in (5, 4, 3, 2, 1):
        k, pad = D._SPEC[i - 1]
        kq = (k + 1) // 2
        ds = ws["ds"][i - 1]
        conv, head, wgt = D._conv(i), D._head(i), D._weight(i)
        dzp = S(f"dzp{i}", ds.h, ds.w, rgb_c, k - 1)
        dp = dpreds[i - 1]
        if dp is None:
            dp = torch.zeros_like(ws["preds"][i - 1])
        dp = dp.contiguous().float()
        K.head_bwd(dp, ws["preds"][i - 1], None, head_mode, dzp)
        if need_w:
            gw, direct = gbuf(f"d{i}_pred.0.1.weight", head.weight)
            if K.zwin_ok(1, ds, dzp, k):
                _Side.run(dev, side, lambda ds=ds, dzp=dzp, gw=gw, k=k, pad=pad:
                          K.conv_wgrad(ds, dzp, gw, k, 1, pad, dz_zero_halo=True))
            elif K.hstack_ok(1, ds.c, k):
                ep = S(f"dzpe{i}", ds.h, ds.w + k - 1, 32, 0, share=dzp)
                K.dz_hstack(dzp, 1, k, ep)
                _Side.run(dev, side, lambda ds=ds, ep=ep, gw=gw, k=k, pad=pad: K.conv_wgrad_hstack(ds, ep, gw, k, pad))
            else:
                dzpw = S(f"dzpw{i}", ds.h, ds.w, 32, 0, share=dzp)
                K.head_bwd(dp, ws["preds"][i - 1], None, head_mode, dzpw)
                _Side.run(dev, side, lambda ds=ds, dzpw=dzpw, gw=gw, k=k, pad=pad: K.conv_wgrad(ds, dzpw, gw, k, 1, pad))
            grads[f"d{i}_pred.0.1.weight"] = None if direct else gw
        dxa = S(f"dxa{i}", ds.h + 2 * pad, ds.w + 2 * pad, ds.c)
        K.conv_dgrad(dzp, head.weight, k, 1, dxa, cache, f"p{i}", w_scale=wsp(i))
        dz = S(f"dz{i}", ds.h, ds.w, ds.c, kq - 1)
        if carry is not None:
            if _FOLD_INPLACE:
                tmp = K.fold_inplace(carry[0], carry[1])
            else:
                tmp = S(f"tmp{i}", ds.h, ds.w, ds.c)
                K.grad_combine(tmp, ds.c, src_a=carry[0], pad_a=carry[1])
            K.grad_combine(dz, ds.c, src_a=dxa, pad_a=pad, add_b=tmp, mask=ds, act=D._act)
        else:
            K.grad_combine(dz, ds.c, src_a=dxa, pad_a=pad, mask=ds, act=D._act)
        # strided SN conv d_i: input srcs[i-1]
        xin = srcs[i - 1]
        alpha = ws["sig"][i - 1][1:2] if D.use_sn else None
        if need_w:
            if D.use_sn:
                # the gradient w.r.t. W_sn of THIS pass (u, v, sigma differ between the passes of a step) in a per-layer
                # scratch, then the rank-one correction, added into the sink by the same kernel
                wname = f"d{i}.0.1.weight_orig"
                tgt, direct = gbuf(wname, wgt, zero=False)
                if direct:
                    gw = _scratch.store.get(("Dsn", i, str(dev)))
                    if gw is None:
                        gw = _scratch.store[("Dsn", i, str(dev))] = torch.empty_like(wgt, memory_format=torch.contiguous_format)
                else:
                    gw = tgt
                dot_ws = _scratch.store.get(("Dsn_dot", i, str(dev)))
                if dot_ws is None:
                    dot_ws = _scratch.store[("Dsn_dot", i, str(dev))] = torch.empty(1, dtype=torch.float64, device=dev)

                def run_w(gw=gw, tgt=tgt, direct=direct, xin=xin, dz=dz, k=k, pad=pad, alpha=alpha, wgt=wgt, i=i, dot_ws=dot_ws):
                    K.zero_(gw)
                    K.conv_wgrad(xin, dz, gw, k, 2, pad, alpha=alpha)
                    K.spectral_bwd(gw, wgt, ws["u"][i - 1], ws["v"][i - 1], ws["sig"][i - 1], dot_ws,
                                   accum=tgt if direct else None)
                _Side.run(dev, side, run_w)
                grads[wname] = None if direct else gw
            else:
                gw, direct = gbuf(f"d{i}.0.1.weight", wgt)
                _Side.run(dev, side, lambda gw=gw, xin=xin, dz=dz, k=k, pad=pad, alpha=alpha:
                          K.conv_wgrad(xin, dz, gw, k, 2, pad, alpha=alpha))
                grads[f"d{i}.0.1.weight"] = None if direct else gw
            gb, direct = gbuf(f"d{i}.0.1.bias", conv.bias, zero=False)
            _Side.run(dev, side, lambda dz=dz, gb=gb, direct=direct: K.channel_sum(dz, gb, accumulate=direct))
            grads[f"d{i}.0.1.bias"] = None if direct else gb
        if i > 1:
            dxb = S(f"dxb{i}", xin.h + 2 * pad, xin.w + 2 * pad, xin.c)
            K.conv_dgrad(dz, wgt, k, 2, dxb, cache, f"d{i}", alpha=alpha, w_scale=wsc(i))
            carry = (dxb, pad)
        elif need_dx:
            dxb = S("dxb1", h + 2 * pad, w + 2 * pad, 16)
            K.conv_dgrad(dz, wgt, k, 2, dxb, cache, "d1", alpha=alpha, w_scale=wsc(1))
            dx0 = S("dx0", h, w, 16)
            K.grad_combine(dx0, 16, src_a=dxb, pad_a=pad)
            dx = torch.empty_like(x)
            K.unpack_input_grad(dx0, None, dx)
            _Side.join(dev, side)
            return grads, dx
    _Side.join(dev, side)
    return grads, None


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        b, _, h, w = x.shape
        if module.conv_dim % 32:
            raise NotImplementedError("training kernels need conv_dim to be a multiple of 32 (wgrad operand rows)")
        x = x.detach().contiguous().float()
        key = (b, h, w, str(x.device), module.precision)
        pool = module._train_pool.setdefault(key, [])
        ws = pool.pop() if pool else _d_workspace(module, b, h, w, x.device)
        preds = _d_forward_train(module, x, ws)
        ctx.module, ctx.ws, ctx.key, ctx.x = module, ws, key, x
        return tuple(preds)

    @staticmethod
    def backward(ctx, *dpreds):
        D = ctx.module
        if ctx.ws is None:
            raise RuntimeError("uegan_b200: second backward through the same Discriminator forward (its activation "
                               "workspace has been recycled); run the forward again")
        need_w = any(ctx.needs_input_grad[2:])
        grads, dx = _d_backward(D, ctx.x, dpreds, ctx.ws, ctx.needs_input_grad[1], need_w)
        D._train_pool[ctx.key].append(ctx.ws)
        ctx.ws = None
        named = list(D.named_parameters())
        out = [grads.get(n) for n, _ in named]
        if need_w and dx is None and all(g is None for g in out):
            # every gradient went straight into the sink.  A backward node that hands autograd NOTHING invalidates an
            # ongoing CUDA-graph capture (measured r2: capture fails iff all outputs are None); one zero gradient for the
            # smallest parameter keeps a leaf on the device queue.  `p.grad += 0` is exact.
            j = min(range(len(named)), key=lambda i: named[i][1].numel())
            out[j] = torch.zeros_like(named[j][1])
        return (None, dx) + tuple(out)


def discriminator_apply(module, x):
    if x.dim() != 4 or x.shape[1] != 3 or min(x.shape[2:]) < 96:
        raise ValueError("Discriminator needs (B,3,H,W) with H, W >= 96")
    if module._act is None:
        raise NotImplementedError("activation function [%s] has no sm_100a epilogue" % module.act_fun)
    if not hasattr(module, "_train_pool"):
        module._train_pool = {}
    return list(_DiscriminatorFn.apply(module, x, *[p for _, p in module.named_parameters()]))


# =================================================================================================
# losses
# =================================================================================================
class _GanLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mode, for_d, n, group, *maps):
        real = [t.detach().contiguous().float() for t in maps[:n]]
        fake = [t.detach().contiguous().float() for t in maps[n:]]
        ws = torch.empty(48, dtype=torch.float64, device=real[0].device)
        loss = torch.empty((), dtype=torch.float32, device=real[0].device)
        ctx.world = K.gan_loss_fwd(mode, for_d, real, fake, ws, loss, group)
        ctx.mode, ctx.for_d, ctx.n, ctx.real, ctx.fake, ctx.ws = mode, for_d, n, real, fake, ws
        return loss

    @staticmethod
    def backward(ctx, gout):
        n = ctx.n
        need = ctx.needs_input_grad[4:]
        d_real = [torch.zeros_like(t) if need[i] else None for i, t in enumerate(ctx.real)]
        d_fake = [torch.zeros_like(t) if need[n + i] else None for i, t in enumerate(ctx.fake)]
        g = gout.detach().reshape(1).float().contiguous()
        K.gan_loss_bwd(ctx.mode, ctx.for_d, ctx.real, ctx.fake, ctx.ws, d_real, d_fake, gscale_dev=g, world=ctx.world)
        return (None, None, None, None) + tuple(d_real) + tuple(d_fake)


def gan_loss_apply(mode, for_d, real, fake, group=None):
    return _GanLossFn.apply(mode, for_d, len(real), group, *real, *fake)


class _MsRecFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rec_type, scales, pred, gt):
        p, g = pred.detach().contiguous().float(), gt.detach().contiguous().float()
        accum = torch.empty(3, dtype=torch.float64, device=p.device)
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        K.msrec_loss(p, g, rec_type, scales, accum, loss)
        ctx.rec_type, ctx.scales, ctx.p, ctx.g = rec_type, scales, p, g
        return loss

    @staticmethod
    def backward(ctx, gout):
        grad = torch.empty_like(ctx.p)
        accum = torch.empty(3, dtype=torch.float64, device=grad.device)
        loss = torch.empty(1, dtype=torch.float32, device=grad.device)
        K.msrec_loss(ctx.p, ctx.g, ctx.rec_type, ctx.scales, accum, loss, grad, 1.0,
                     gout.detach().reshape(1).float().contiguous())
        return None, None, grad, None


def msrec_apply(module, pred, gt):
    return _MsRecFn.apply(module.rec_type, module.scales, pred, gt)


class _PerceptualFn(torch.autograd.Function):
    """PerceptualLoss forward + backward w.r.t. x (the frozen tower has no weight gradients, losses.py:117-118)."""

    @staticmethod
    def forward(ctx, module, x, y):
        x = x.detach().contiguous().float()
        y = y.detach().contiguous().float()
        vgg = module.vgg
        joint = module.joint_taps()
        taps_x, Px = vgg.run(x, "x", module.eps, stats=not joint)
        taps_y, _ = vgg.run(y, "y", module.eps, stats=not joint)
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        ctx.joint = None
        if joint:
            # statistics of both towers' taps, the MSE and the backward sums from ONE pass per tap
            ctx.joint = module.tap_terms(taps_x, taps_y, Px, loss)
        else:
            if module._accum is None or module._accum.device != x.device:
                module._accum = torch.zeros(1, dtype=torch.float64, device=x.device)
            for wgt, (tx, mx), (ty, my) in zip(module.weights, taps_x, taps_y):
                K.in_mse_fwd(tx, ty, mx, my, wgt, module._accum, loss)
        ctx.module, ctx.taps_x, ctx.taps_y, ctx.Px, ctx.shape = module, taps_x, taps_y, Px, x.shape
        return loss

    @staticmethod
    def backward(ctx, gout):
        from .losses import _VGG_LAYERS, _TAP_IDX, IMAGENET_STD
        module, vgg, P = ctx.module, ctx.module.vgg, ctx.Px
        acts = P["acts"]
        b, _, h, w = ctx.shape
        dev = acts[0].buf.device
        g = gout.detach().reshape(1).float().contiguous()
        # fp16 gradient chain with a fixed loss scale: per-element gradients of a mean over ~1e8 features are ~1e-9,
        # far below fp16's range; S makes the deepest tap's gradient O(1) (the others are <= 2^-11 of that), and is
        # divided out when the image gradient is unpacked.  fp16 keeps 10 mantissa bits (bf16's 8 cost 12 % rel-L2).
        t5 = ctx.taps_x[4][0]
        # (the stored tower -- power-of-two factors folded into the weights, losses.VGG19_relu -- computes the same loss as
        # a function of the image, with O(1..32) activations at every layer whatever the checkpoint's own magnitudes are;
        # its gradient w.r.t. the image IS the true one, so no factor rides along.)
        scale = float(t5.n * t5.c * t5.h * t5.w) / (2.0 * module.weights[4])
        # gradient buffers are zero-initialised once: nobody writes their halo, which is the dgrad's zero padding
        if "grads" not in P:
            P["grads"] = [K.NHWC(a.n, a.h, a.w, max(a.c, 16), 1, L.F16, dev, zero=True) for a in acts]
            P["gws"] = torch.empty(2 * b * 512, dtype=torch.float64, device=dev)
        G = P["grads"]
        tap_of = {}
        ti = 0
        for li, spec in enumerate(_VGG_LAYERS):
            if spec != "M" and spec[0] in _TAP_IDX:
                tap_of[li + 1] = ti
                ti += 1
        have = [False] * len(acts)  # have[j]: G[j] holds dL/d(acts[j]) (for conv outputs: already ReLU-masked)
        for li in range(len(_VGG_LAYERS) - 1, -1, -1):
            spec = _VGG_LAYERS[li]
            out_idx = li + 1
            if spec == "M":
                # acts[out_idx] = pool(acts[li]); route to the arg-max and apply the ReLU mask of acts[li]
                K.maxpool2x2_bwd(acts[li], G[out_idx], G[li])
                have[li] = True
                continue
            idx, cin, cout = spec
            if out_idx in tap_of:
                t = tap_of[out_idx]
                (tx, mx), (ty, my) = ctx.taps_x[t], ctx.taps_y[t]
                deep = G[out_idx] if have[out_idx] else None
                dzt = _scratch.get(f"vgg_dz{out_idx}", tx.n, tx.h, tx.w, tx.c, 1, L.F16, dev, zero=True)
                if ctx.joint is not None:
                    jx, jy, jsums = ctx.joint[t]
                    K.in_mse_bwd_apply(tx, ty, jx, jy, module.weights[t] * scale, g, deep, dzt, jsums)
                else:
                    K.in_mse_bwd(tx, ty, mx, my, module.weights[t] * scale, g, deep, dzt, P["gws"])
                dz = dzt
            else:
                dz = G[out_idx]  # masked by the producer (dgrad epilogue mask or max-pool backward)
            src = acts[li]
            key = ("vgg_dg", idx)
            hit = vgg._wcache.get(key)
            if hit is None:
                hit = (None, K.packed_weight_dgrad(vgg.layer(idx)[0], dz.c, L.F16, 1, 0, 0))
                vgg._wcache[key] = hit
            # the producer of `src`: a conv (ReLU mask needed, unless it is a tap, masked later) or a pool / the input
            prev_is_conv = li > 0 and _VGG_LAYERS[li - 1] != "M"
            prev_is_tap = li in tap_of
            mask = src if (prev_is_conv and not prev_is_tap) else None
            cout_dg = cin if cin != 3 else 16
            K.conv_generic(dz, hit[1], cout_dg, 3, 1, 1, G[li], 0, None, None, L.ACT_NONE, mask, L.ACT_RELU)
            have[li] = True
        dx = torch.empty(b, 3, h, w, dtype=torch.float32, device=dev)
        K.unpack_input_grad(G[0], [1.0 / (s * scale) for s in IMAGENET_STD], dx)
        return None, dx, None


def perceptual_apply(module, x, y):
    return _PerceptualFn.apply(module, x, y)
