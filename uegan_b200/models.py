"""Drop-in `models` module: `Generator` / `Discriminator` with the reference constructors, parameter names
(state_dict keys) and forward contracts (/root/reference/models.py:10-74, 104-182), executing on the
hand-written sm_100a kernels behind include/uegan_sm100.h.

The torch.nn modules built here are PARAMETER CONTAINERS ONLY (so that `.parameters()`, `.apply(init_fn)`,
`.state_dict()` / `.load_state_dict()` and checkpoints behave exactly like the reference's);
they are never called.  `forward` runs the native plan; there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib as L
from . import kernels as K

_ACTS = {"LeakyReLU": L.ACT_LRELU, "ReLU": L.ACT_RELU, "none": L.ACT_NONE}


def _act_module(name):
    # get_act_fun, models.py:249-264 (containers only; Swish/SELU have no native epilogue yet)
    if name == "LeakyReLU":
        return nn.LeakyReLU(0.2, inplace=True)
    if name == "ReLU":
        return nn.ReLU(inplace=True)
    if name == "none":
        return nn.Sequential()
    if name in ("Swish", "SELU"):
        raise NotImplementedError("activation function [%s] has no sm_100a epilogue in uegan_b200" % name)
    raise NotImplementedError("activation function [%s] is not found" % name)


def _check_norm(name):
    # get_norm_fun, models.py:272-281
    if name == "none":
        return
    if name in ("BatchNorm", "InstanceNorm"):
        raise NotImplementedError("normalization function [%s] has no sm_100a path in uegan_b200" % name)
    raise NotImplementedError("normalization function [%s] is not found" % name)


class Identity(nn.Module):
    pass


def _holder(cin, cout, k, stride=1, bias=True, extra=()):
    """ReflectionPad2d + Conv2d (+ norm placeholder + activation) holder with the reference's child indices."""
    mods = [nn.ReflectionPad2d((k - 1) // 2), nn.Conv2d(cin, cout, k, stride=stride, padding=0, bias=bias)]
    mods.extend(extra)
    return nn.Sequential(*mods)


class _MainHolder(nn.Module):
    """ConvBlock / SNConv shell: parameters live under `.main.1` (models.py:77-101)."""

    def __init__(self, cin, cout, k, stride=1, extra=()):
        super().__init__()
        self.main = _holder(cin, cout, k, stride, True, extra)

    @property
    def conv(self):
        return self.main[1]


class _GAMHolder(nn.Module):
    """GAM parameters (models.py:215-228): conv.0, conv.2 (no bias), fuse.0 (bias)."""

    def __init__(self, c, reduction=8):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(2 * c, c // reduction, 1, bias=False), nn.ReLU(inplace=True),
                                  nn.Conv2d(c // reduction, c, 1, bias=False))
        self.fuse = nn.Sequential(nn.Conv2d(2 * c, c, 1, bias=True))


class _Interp(nn.Module):
    pass


class _ParamCache:
    """Re-packs a parameter into its kernel operand only when the parameter changed (optimizer step / load).
    The operand buffer of a key is allocated once and re-packed IN PLACE (`fn(out=buf)`): consumers captured in a
    CUDA graph keep reading the same address, and a re-pack node later in the graph refreshes it for the next replay."""

    def __init__(self):
        self.store = {}
        self.epoch = 0

    def bump(self):
        """The parameters were rewritten by a kernel torch does not see (uegan_adam_step): every operand is stale."""
        self.epoch += 1

    def get(self, key, param, fn):
        tag = (param.data_ptr(), param._version, self.epoch)
        hit = self.store.get(key)
        if hit is None:
            hit = (tag, fn())
        elif hit[0] != tag:
            hit = (tag, fn(out=hit[1]))
        else:
            return hit[1]
        self.store[key] = hit
        return hit[1]


class _ScaledNet:
    """Operand precision of a network's convolutions and, for fp16, the device-resident power-of-two scales of its master
    weights.  precision = "f16" (default; UEGAN_GD_DTYPE overrides): fp16 storage / kind::f16 MMAs with per-tensor
    power-of-two scales maintained on the device (uegan_scale_update): the same 11 significant bits as tf32 at twice the
    tensor rate and half the bytes per pass.  precision = "tf32": fp32 storage, kind::tf32 MMAs (the reference's own default
    GPU arithmetic).  Both meet the same parity gates (tests/test_gpu_generator*.py, test_gpu_pinned_chain.py)."""

    def _init_precision(self):
        self.precision = __import__("os").environ.get("UEGAN_GD_DTYPE", "f16")
        self._wbook = None  # ScaleBook of the master weights (f16 mode)

    @property
    def _dtype(self):
        if self.precision not in ("tf32", "f16"):
            raise ValueError("precision must be 'tf32' or 'f16'")
        return L.F16 if self.precision == "f16" else L.F32

    def _wscale(self, name, conv):
        """Device scalar the packed fp16 weight of `conv` is multiplied by (None in tf32 mode)."""
        if self._dtype == L.F32:
            return None
        w = conv if isinstance(conv, torch.Tensor) else conv.weight
        wb = self._wbook
        if wb is None or wb.buf.device != w.device:
            wb = self._wbook = K.ScaleBook(w.device)
            wb.items = {}
        it = wb.items.get(name)
        if it is None or it[0] is not w:
            it = wb.items[name] = (w, it[1] if it is not None else wb.slot())
        return it[1]

    def _update_weight_scales(self):
        """One launch: every weight's scale from its CURRENT fp32 master values (exact, not delayed); skipped while the
        parameters are unchanged (inference).  The table is rebuilt when a parameter moved (FlatBucket, .to())."""
        wb = self._wbook
        if wb is None or not wb.items:
            return
        ptrs = tuple(w.data_ptr() for w, _ in wb.items.values())
        if getattr(wb, "ptrs", None) != ptrs:
            wb.clear_tracked()
            for w, sl in wb.items.values():
                wb.track(w.data_ptr(), w.numel(), L.F32, sl, true_values=True)
            wb.ptrs, wb.tag = ptrs, None
        tag = tuple(w._version for w, _ in wb.items.values()) + (self._wcache.epoch,)
        if wb.tag != tag:
            wb.update()
            wb.tag = tag



class Generator(nn.Module, _ScaledNet):
    """Generator network (reference: models.py:10-74).  forward: (B,3,H,W) fp32 in [-1,1], H,W % 16 == 0."""

    def __init__(self, conv_dim, norm_fun, act_fun, use_sn):
        super().__init__()
        _check_norm(norm_fun)
        if use_sn:
            raise NotImplementedError("spectral norm inside the Generator has no sm_100a path in uegan_b200 "
                                      "(reference default g_use_sn=False, config.py:23)")
        self.conv_dim, self.act_fun = conv_dim, act_fun
        self._act = _ACTS.get(act_fun)
        d = conv_dim

        def block(cin, cout, k, stride=1):
            return _MainHolder(cin, cout, k, stride, (Identity(), _act_module(act_fun)))

        self.enc1 = block(3, d, 7)
        self.enc2 = block(d, 2 * d, 3, 2)
        self.enc3 = block(2 * d, 4 * d, 3, 2)
        self.enc4 = block(4 * d, 8 * d, 3, 2)
        self.enc5 = block(8 * d, 16 * d, 3, 2)
        self.upsample1 = nn.Sequential(_Interp(), _MainHolder(16 * d, 8 * d, 1))
        self.upsample2 = nn.Sequential(_Interp(), _MainHolder(8 * d, 4 * d, 1))
        self.upsample3 = nn.Sequential(_Interp(), _MainHolder(4 * d, 2 * d, 1))
        self.upsample4 = nn.Sequential(_Interp(), _MainHolder(2 * d, d, 1))
        self.dec1 = block(16 * d, 8 * d, 3)
        self.dec2 = block(8 * d, 4 * d, 3)
        self.dec3 = block(4 * d, 2 * d, 3)
        self.dec4 = block(2 * d, d, 3)
        self.dec5 = nn.Sequential(_MainHolder(d, d, 3), _MainHolder(d, 3, 7), nn.Tanh())
        self.ga5 = _GAMHolder(16 * d)
        self.ga4 = _GAMHolder(8 * d)
        self.ga3 = _GAMHolder(4 * d)
        self.ga2 = _GAMHolder(2 * d)
        self.ga1 = _GAMHolder(d)
        self._plans = {}
        self._wcache = _ParamCache()
        self._init_precision()

    # ---------------------------------------------------------------- native forward
    def _w(self, name, conv, cin_stored, cin_first=0, cin=None):
        dt = self._dtype
        ws = self._wscale(name, conv)
        return self._wcache.get((name, dt), conv.weight,
                                lambda out=None: K.packed_weight(conv.weight, cin_stored, dt, cin_first, cin, out=out,
                                                                 w_scale=ws))

    def _plan(self, b, h, w, device):
        key = (b, h, w, str(device), self.precision)
        pl = self._plans.get(key)
        if pl is None:
            d = self.conv_dim
            dt = self._dtype
            book = K.ScaleBook(device) if dt != L.F32 else None

            def T(hh, ww, c, halo=0, zero=False, scaled=True):
                if book is None:
                    return K.NHWC(b, hh, ww, c, halo, L.F32, device, zero)
                t = K.NHWC(b, hh, ww, c, halo, dt, device, True, scale=book.slot() if scaled else None)
                if scaled:
                    book.track_nhwc(t)
                return t
            pl = dict(
                book=book,
                x0=T(h, w, 4 if dt == L.F32 else 8, 3, zero=True, scaled=False),
                x1=T(h, w, d, 1), x2=T(h // 2, w // 2, 2 * d, 1), x3=T(h // 4, w // 4, 4 * d, 1),
                x4=T(h // 8, w // 8, 8 * d, 1), x5=T(h // 16, w // 16, 16 * d),
                z5=T(h // 16, w // 16, 16 * d), x5n=T(h // 16, w // 16, 16 * d),
                u1=T(h // 16, w // 16, 8 * d), cat1=T(h // 8, w // 8, 16 * d, 1), z4=T(h // 8, w // 8, 8 * d),
                y1=T(h // 8, w // 8, 8 * d),
                u2=T(h // 8, w // 8, 4 * d), cat2=T(h // 4, w // 4, 8 * d, 1), z3=T(h // 4, w // 4, 4 * d),
                y2=T(h // 4, w // 4, 4 * d),
                u3=T(h // 4, w // 4, 2 * d), cat3=T(h // 2, w // 2, 4 * d, 1), z2=T(h // 2, w // 2, 2 * d),
                y3=T(h // 2, w // 2, 2 * d),
                u4=T(h // 2, w // 2, d), cat4=T(h, w, 2 * d, 1), z1=T(h, w, d),
                y4m=T(h, w, d, 1), t=T(h, w, d, 3),
                stats=torch.empty(3 * b * 16 * d, dtype=torch.float64, device=device),
            )
            self._plans[key] = pl
        return pl

    def forward(self, x):
        if not x.is_cuda:
            raise L.UeganError("uegan_b200.models.Generator runs on CUDA (sm_100a) only; no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import generator_apply  # training path (fprop + dgrad + wgrad kernels)
            return generator_apply(self, x)
        return self.forward_native(x)

    @torch.no_grad()
    def forward_native(self, x, keep=None, packed=False):
        """packed=True: the plan's x0 operand already holds x (uegan_b200.io.pack_u8 wrote both from uint8)."""
        assert x.dim() == 4 and x.shape[1] == 3, "expected (B,3,H,W)"
        b, _, h, w = x.shape
        if h % 16 or w % 16 or h < 32 or w < 32:
            raise ValueError("Generator needs H, W multiples of 16 and >= 32 (models.py:55-67 skip concat)")
        if self._act is None:
            raise NotImplementedError("activation function [%s] has no sm_100a epilogue" % self.act_fun)
        x = x.contiguous().float()
        P = self._plan(b, h, w, x.device)
        book = P["book"]
        if book is not None and not P.get("calibrated"):
            # fp16 storage: settle the per-tensor scales on this batch (each pass fixes at least the next layer; the
            # activations of the reference's orthogonal(0.02) init need scales up to 2^40).  One-off per shape.
            prev = None
            for it in range(48):
                self._forward_pass(x, P, packed and it == 0)
                book.update()
                cur = book.values()
                if prev is not None and torch.equal(cur, prev):
                    break
                prev = cur
            P["calibrated"] = True
            packed = True  # x0 is in place
        elif book is not None:
            book.update()  # delayed scaling: this pass uses the magnitudes the previous pass stored
        out = self._forward_pass(x, P, packed)
        if keep is not None:
            keep.update(P)
        return out

    def _forward_pass(self, x, P, packed):
        d, act = self.conv_dim, self._act
        c = lambda holder: holder.conv
        if self._dtype != L.F32:
            # register every weight's scale slot, then ONE launch refreshes them all from the fp32 masters
            for name, holder in self._conv_table():
                self._wscale(name, holder)
            self._update_weight_scales()

        # inference: the reflection halo by a halo_fill launch after the conv (same-box A/B r4i: 4.94 vs 5.03 ms per batch
        # of 32 with the mirrored stores in the epilogue, whose per-tile chain is what bounds these layers; the training
        # step, where it is time-neutral and saves 35 launches, uses the epilogue: autograd.py)
        halo_launch = os.environ.get("UEGAN_EPILOGUE_HALO") != "1"

        def conv(src, name, holder, cout, k, stride, dst, act_=L.ACT_NONE, off=0, mul=None, bias=True, halo=False):
            # halo: dst's reflection-padding halo is written too (by the epilogue, or by a halo_fill launch after it)
            cv = c(holder)
            K.conv_fprop(src, self._w(name, cv, src.c), cout, k, stride, (k - 1) // 2, dst, off,
                         cv.bias if bias else None, None, act_, mul, w_scale=self._wscale(name, cv), reflect_halo=halo,
                         halo_launch=halo_launch)

        if not packed:
            K.pack_input(x, P["x0"], L.PAD_REFLECT)
        conv(P["x0"], "enc1", self.enc1, d, 7, 1, P["x1"], act, halo=True)
        conv(P["x1"], "enc2", self.enc2, 2 * d, 3, 2, P["x2"], act, halo=True)
        conv(P["x2"], "enc3", self.enc3, 4 * d, 3, 2, P["x3"], act, halo=True)
        conv(P["x3"], "enc4", self.enc4, 8 * d, 3, 2, P["x4"], act, halo=True)
        conv(P["x4"], "enc5", self.enc5, 16 * d, 3, 2, P["x5"], act)

        def gam(name, ga, src, ch, z, dst, off, up=None):
            # GAM(x) == IN(conv1x1(x, fuse.weight[:, :C])): the attention branch and fuse bias are constant per
            # (n, c) and cancel in the InstanceNorm that follows (models.py:230-237; SURVEY.md 8a rewrite 1).
            # up: the low-resolution tensor whose bilinear x2 is the other half of dst: one kernel then writes whole pixels
            fuse = ga.fuse[0]
            dt = self._dtype
            ws_ = self._wscale(name, fuse)
            wp = self._wcache.get((name, dt), fuse.weight,
                                  lambda out=None: K.packed_weight(fuse.weight, src.c, dt, 0, ch, out=out, w_scale=ws_))
            if up is not None:
                K.conv_fprop(src, wp, ch, 1, 1, 0, z, w_scale=ws_)
                K.cat_build(up, z, P["stats"], dst)
                return
            if dt == L.F32 and K.fused_stats_ok(src.h, src.w, ch):  # statistics ride in the conv epilogue
                K.conv_fprop(src, wp, ch, 1, 1, 0, z, in_stats=P["stats"])
                K.instance_norm_apply(z, dst, off, P["stats"])
            else:
                K.conv_fprop(src, wp, ch, 1, 1, 0, z, w_scale=ws_)
                K.instance_norm(z, dst, off, P["stats"])

        gam("ga5", self.ga5, P["x5"], 16 * d, P["z5"], P["x5n"], 0)
        stages = [("upsample1", self.upsample1, "ga4", self.ga4, "dec1", self.dec1, "x5n", "x4", "u1", "cat1", "z4", "y1", 8 * d),
                  ("upsample2", self.upsample2, "ga3", self.ga3, "dec2", self.dec2, "y1", "x3", "u2", "cat2", "z3", "y2", 4 * d),
                  ("upsample3", self.upsample3, "ga2", self.ga2, "dec3", self.dec3, "y2", "x2", "u3", "cat3", "z2", "y3", 2 * d),
                  ("upsample4", self.upsample4, "ga1", self.ga1, "dec4", self.dec4, "y3", "x1", "u4", "cat4", "z1", "y4m", d)]
        for un, up, gn, ga, dn, dec, src, skip, u, cat, z, y, ch in stages:
            # conv1x1 then bilinear x2 == bilinear x2 then conv1x1 (both linear, lerp weights sum to 1): run the
            # 1x1 at low resolution (models.py:23-26; SURVEY.md 8a rewrite 2)
            conv(P[src], un, up[1], ch, 1, 1, P[u])
            if K.cat_build_ok():
                gam(gn, ga, P[skip], ch, P[z], P[cat], ch, up=P[u])
            else:
                K.upsample2x(P[u], P[cat], 0)
                gam(gn, ga, P[skip], ch, P[z], P[cat], ch)
            K.halo_fill(P[cat])
            last = dn == "dec4"
            conv(P[cat], dn, dec, ch, 3, 1, P[y], act, 0, P["x1"] if last else None, halo=last)  # dec4 fuses y4.mul(x1)
        conv(P["y4m"], "dec5.0", self.dec5[0], d, 3, 1, P["t"], halo=True)
        out = torch.empty_like(x)
        cv = c(self.dec5[1])
        K.conv_planar(P["t"], cv.weight, self._wcache, "dec5.1", 7, 3, cv.bias, None, L.ACT_TANH, out, x,
                      w_scale=self._wscale("dec5.1", cv))
        return out

    def _conv_table(self):
        """(cache name, module holding .weight) of every convolution the native plan runs."""
        t = [("enc1", self.enc1.conv), ("enc2", self.enc2.conv), ("enc3", self.enc3.conv), ("enc4", self.enc4.conv),
             ("enc5", self.enc5.conv), ("dec1", self.dec1.conv), ("dec2", self.dec2.conv), ("dec3", self.dec3.conv),
             ("dec4", self.dec4.conv), ("dec5.0", self.dec5[0].conv), ("dec5.1", self.dec5[1].conv)]
        t += [(f"upsample{i}", getattr(self, f"upsample{i}")[1].conv) for i in range(1, 5)]
        t += [(f"ga{i}", getattr(self, f"ga{i}").fuse[0]) for i in range(1, 6)]
        return t


class Discriminator(nn.Module, _ScaledNet):
    """Multi-scale PatchGAN discriminator (reference: models.py:104-182): five spectrally-normalised strided convs,
    each followed by a Cout=1 prediction head; forward returns the list of five (B,1,H/2^k,W/2^k) maps."""

    _SPEC = [(7, 3), (7, 3), (7, 3), (5, 2), (5, 2)]  # (kernel, pad) of d{k} and d{k}_pred, models.py:109-126

    def __init__(self, conv_dim, norm_fun, act_fun, use_sn, adv_loss_type):
        super().__init__()
        _check_norm(norm_fun)
        if adv_loss_type in ("ls", "rals"):
            self._head_act = L.ACT_SIGMOID
        elif adv_loss_type in ("hinge", "rahinge"):
            self._head_act = L.ACT_TANH
        else:
            raise NotImplementedError("Adversarial loss [{}] is not found".format(adv_loss_type))
        self.conv_dim, self.act_fun, self.use_sn = conv_dim, act_fun, bool(use_sn)
        self._act = _ACTS.get(act_fun)
        d = conv_dim
        chans = [3, d, 2 * d, 4 * d, 8 * d, 16 * d]
        for i, (k, _) in enumerate(self._SPEC, start=1):
            conv = nn.Conv2d(chans[i - 1], chans[i], k, stride=2, padding=0, bias=True)
            if use_sn:
                conv = nn.utils.spectral_norm(conv)  # container semantics only: weight_orig / weight_u / weight_v
            body = nn.Sequential(nn.ReflectionPad2d((k - 1) // 2), conv, Identity(), _act_module(act_fun))
            setattr(self, f"d{i}", nn.Sequential(body))
            head_act = nn.Sigmoid() if self._head_act == L.ACT_SIGMOID else nn.Tanh()
            head = nn.Sequential(nn.ReflectionPad2d((k - 1) // 2), nn.Conv2d(chans[i], 1, k, bias=False), head_act)
            setattr(self, f"d{i}_pred", nn.Sequential(head))
        self._plans = {}
        self._wcache = _ParamCache()
        self._init_precision()

    def _conv(self, i):
        return getattr(self, f"d{i}")[0][1]

    def _head(self, i):
        return getattr(self, f"d{i}_pred")[0][1]

    def _weight(self, i):
        c = self._conv(i)
        return c.weight_orig if self.use_sn else c.weight

    def _register_weight_scales(self):
        if self._dtype != L.F32:
            for i in range(1, 6):
                self._wscale(f"d{i}", self._weight(i))
                self._wscale(f"p{i}", self._head(i).weight)
            self._update_weight_scales()

    def _act_buffers(self, b, h, w, device):
        """x0 + the five activation tensors (+ their ScaleBook in f16 mode) of one forward pass."""
        d = self.conv_dim
        dt = self._dtype
        chans = [d, 2 * d, 4 * d, 8 * d, 16 * d]
        book = K.ScaleBook(device) if dt != L.F32 else None
        out = dict(book=book, dtype=dt, x0=K.NHWC(b, h, w, 4 if dt == L.F32 else 8, 3, dt, device, zero=True), ds=[])
        hh, ww = h, w
        for i in range(5):
            hh, ww = (hh + 1) // 2, (ww + 1) // 2
            halo = 3 if i < 3 else 2
            if book is None:
                t = K.NHWC(b, hh, ww, chans[i], halo, L.F32, device)
            else:
                t = K.NHWC(b, hh, ww, chans[i], halo, dt, device, True, scale=book.slot())
                book.track_nhwc(t)
            out["ds"].append(t)
        return out

    def _plan(self, b, h, w, device):
        key = (b, h, w, str(device), self.precision)
        pl = self._plans.get(key)
        if pl is None:
            pl = self._act_buffers(b, h, w, device)
            pl.update(sig=[], ws=[])
            for i in range(5):
                wgt = self._weight(i + 1)
                pl["sig"].append(torch.ones(2, dtype=torch.float32, device=device))
                pl["ws"].append(torch.empty(wgt.shape[0] + wgt.numel() // wgt.shape[0] + 8, dtype=torch.float32,
                                            device=device))
            self._plans[key] = pl
        return pl

    def forward(self, x):
        if not x.is_cuda:
            raise L.UeganError("uegan_b200.models.Discriminator runs on CUDA (sm_100a) only; no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import discriminator_apply
            return discriminator_apply(self, x)
        return self.forward_native(x)

    @torch.no_grad()
    def forward_native(self, x, keep=None):
        assert x.dim() == 4 and x.shape[1] == 3, "expected (B,3,H,W)"
        b, _, h, w = x.shape
        if min(h, w) < 96:
            raise ValueError("Discriminator needs H, W >= 96 (ReflectionPad2d(2) on the 1/32-scale map, models.py:125)")
        if self._act is None:
            raise NotImplementedError("activation function [%s] has no sm_100a epilogue" % self.act_fun)
        x = x.contiguous().float()
        P = self._plan(b, h, w, x.device)
        book = P["book"]
        if book is not None and not P.get("calibrated"):
            # settle the per-tensor scales (see Generator.forward_native) on the sigma the real pass uses: advance the
            # spectral-norm power iteration ONCE (if in train mode), then settle and run with u / v held (see
            # autograd._d_forward_train)
            P["calibrated"] = True
            was = self.training
            if was and self.use_sn:
                from .autograd import _d_advance_sn
                _d_advance_sn(self, x.device)
            self.training = False
            try:
                prev = None
                for _ in range(48):
                    self._forward_pass(x, P)
                    book.update()
                    cur = book.values()
                    if prev is not None and torch.equal(cur, prev):
                        break
                    prev = cur
                preds = self._forward_pass(x, P)
            finally:
                self.training = was
            if keep is not None:
                keep.update(P)
            return preds
        elif book is not None:
            book.update()
        preds = self._forward_pass(x, P)
        if keep is not None:
            keep.update(P)
        return preds

    def _forward_pass(self, x, P):
        b = x.shape[0]
        self._register_weight_scales()
        dt = self._dtype
        K.pack_input(x, P["x0"], L.PAD_REFLECT)
        src, preds = P["x0"], []
        if self.use_sn:
            # one power iteration per train-mode forward, in place on the weight_u / weight_v buffers; the five layers are
            # independent, so each phase is one launch for all of them
            nl = len(self._SPEC)
            K.spectral_sigma_batch([self._weight(i) for i in range(1, nl + 1)],
                                   [self._conv(i).weight_u for i in range(1, nl + 1)],
                                   [self._conv(i).weight_v for i in range(1, nl + 1)], self.training, P["sig"], P["ws"])
        for i, (k, pad) in enumerate(self._SPEC, start=1):
            conv, head, wgt = self._conv(i), self._head(i), self._weight(i)
            alpha = P["sig"][i - 1][1:2] if self.use_sn else None
            dst = P["ds"][i - 1]
            wsc = self._wscale(f"d{i}", wgt)
            wp = self._wcache.get((f"d{i}", dt), wgt,
                                  lambda out=None: K.packed_weight(wgt, src.c, dt, out=out, w_scale=wsc))
            K.conv_fprop(src, wp, wgt.shape[0], k, 2, pad, dst, 0, conv.bias, alpha, self._act, w_scale=wsc, reflect_halo=True)
            pred = torch.empty(b, 1, dst.h, dst.w, dtype=torch.float32, device=x.device)
            K.conv_planar(dst, head.weight, self._wcache, f"p{i}", k, pad, None, None, self._head_act, pred,
                          w_scale=self._wscale(f"p{i}", head.weight))
            preds.append(pred)
            src = dst
        return preds
