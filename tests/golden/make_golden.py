"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) in this container.

Usage (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py

For every case it builds the reference nn.Modules (models.Generator / models.Discriminator /
losses.PerceptualLoss / GANLoss / MultiscaleRecLoss), loads the deterministic synthetic weights from
oracle.uegan_oracle.make_*_params through load_state_dict (same keys), runs the
reference forward (and one full trainer.py:75-119 step via autograd + torch.optim.Adam) on seeded
inputs and stores the outputs in tests/golden/*.npz (fp32; big maps stored as fp16-lossless subsamples
are avoided: only small shapes are used so files stay small).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import uegan_oracle as O  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    import models as ref_models  # noqa
    # losses.py imports torchvision only; fine
    import losses as ref_losses  # noqa
    return ref_models, ref_losses


def seed_vgg_cache(vp):
    """losses.py:43 calls vgg19(pretrained=True); seed the hub cache with the synthetic tower."""
    import torchvision
    home = os.path.join("/tmp", "uegan_torch_home")
    os.environ["TORCH_HOME"] = home
    ck = os.path.join(home, "hub", "checkpoints")
    os.makedirs(ck, exist_ok=True)
    net = torchvision.models.vgg19(weights=None)
    sd = net.state_dict()
    for k, v in vp.items():
        sd[k] = v.clone()
    torch.save(sd, os.path.join(ck, "vgg19-dcbb9e9d.pth"))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref_models, ref_losses = import_reference()
    vp = O.make_vgg_params()
    seed_vgg_cache(vp)
    for regime in ("o1", "tiny"):
        gp = O.make_generator_params(32, 0, regime)
        dp = O.make_discriminator_params(32, 1, regime)
        G = ref_models.Generator(32, "none", "LeakyReLU", False)
        D = ref_models.Discriminator(32, "none", "LeakyReLU", True, "rahinge")
        G.load_state_dict(gp, strict=True)
        D.load_state_dict(dp, strict=True)
        out = {}
        # ---- config 1 of BASELINE.json: 2x3x128x128 G + D forward
        x = O.make_images((2, 3, 128, 128), 10)
        G.eval()
        with torch.no_grad():
            gout = G(x)
            # pre-clamp residual through the reference modules
            x1 = G.enc1(x); x2 = G.enc2(x1); x3 = G.enc3(x2); x4 = G.enc4(x3); x5 = G.ga5(G.enc5(x4))
            y1 = G.dec1(torch.cat([G.upsample1(x5), G.ga4(x4)], 1))
            y2 = G.dec2(torch.cat([G.upsample2(y1), G.ga3(x3)], 1))
            y3 = G.dec3(torch.cat([G.upsample3(y2), G.ga2(x2)], 1))
            y4 = G.dec4(torch.cat([G.upsample4(y3), G.ga1(x1)], 1))
            res = G.dec5(y4.mul(x1))
        out["g128_out"] = gout.numpy()
        out["g128_res"] = res.numpy()
        D.train()
        with torch.no_grad():
            preds = D(x)
        for i, pmap in enumerate(preds):
            out[f"d128_pred{i+1}"] = pmap.numpy()
        for k in range(1, 6):
            out[f"d128_u{k}"] = D.state_dict()[f"d{k}.0.1.weight_u"].numpy().copy()
            out[f"d128_v{k}"] = D.state_dict()[f"d{k}.0.1.weight_v"].numpy().copy()
        D.eval()
        with torch.no_grad():
            preds = D(x)
        for i, pmap in enumerate(preds):
            out[f"d128_eval_pred{i+1}"] = pmap.numpy()
        # ---- non-square, H,W multiples of 16 (config 5 shapes in miniature)
        x = O.make_images((1, 3, 96, 160), 11)
        with torch.no_grad():
            out["g96x160_out"] = G(x).numpy()
        # ---- losses
        if regime == "o1":
            P = ref_losses.PerceptualLoss()
            a = O.make_images((2, 3, 64, 64), 12)
            b = O.make_images((2, 3, 64, 64), 13)
            with torch.no_grad():
                out["percep64"] = np.float32(P((a + 1) / 2, (b + 1) / 2).item())
                taps = P.vgg(((a + 1) / 2 - P.mean) / P.std)
                for k in ("relu1_1", "relu2_1", "relu3_1", "relu4_1", "relu5_1"):
                    out["vgg64_" + k + "_mean"] = np.float32(taps[k].mean().item())
                    out["vgg64_" + k + "_absmax"] = np.float32(taps[k].abs().max().item())
                out["vgg64_relu5_1"] = taps["relu5_1"].numpy()
                out["vgg64_relu3_1_n0c0"] = taps["relu3_1"][0, :4].numpy()
            gl = ref_losses.GANLoss("rahinge")
            rp = [torch.tanh(O.make_images((2, 1, s, s), 20 + i)) for i, s in enumerate((64, 32, 16, 8, 4))]
            fp = [torch.tanh(O.make_images((2, 1, s, s), 30 + i)) for i, s in enumerate((64, 32, 16, 8, 4))]
            out["rahinge_d"] = np.float32(gl(rp, fp, None, None, for_discriminator=True).item())
            out["rahinge_g"] = np.float32(gl(rp, fp, None, None, for_discriminator=False).item())
            gl2 = ref_losses.GANLoss("rals")
            out["rals_d"] = np.float32(gl2(rp, fp, None, None, for_discriminator=True).item())
            out["rals_g"] = np.float32(gl2(rp, fp, None, None, for_discriminator=False).item())
            ms = ref_losses.MultiscaleRecLoss(3, "l1", True)
            out["msl1"] = np.float32(ms(a, b).item())
            # ---- one full training step: loop body of trainer.py:75-119, pool_size=0, default lambdas
            G.load_state_dict(gp); D.load_state_dict(dp)
            G.train(); D.train()
            g_opt = torch.optim.Adam(G.parameters(), lr=1e-4, betas=[0.5, 0.999], weight_decay=0.0001)
            d_opt = torch.optim.Adam(D.parameters(), lr=4e-4, betas=[0.5, 0.999], weight_decay=0.0001)
            raw = O.make_images((2, 3, 128, 128), 40)
            exp = O.make_images((2, 3, 128, 128), 41)
            for step in range(2):
                fake = G(raw)
                d_opt.zero_grad()
                rpred = D(exp); fpred = D(fake.detach())
                d_loss = gl(rpred, fpred, None, None, for_discriminator=True)
                ipred = D(raw)
                d_loss = d_loss + gl(rpred, ipred, None, None, for_discriminator=True)
                d_loss.backward(); d_opt.step()
                g_opt.zero_grad()
                rpred = D(exp); fpred = D(fake)
                g_adv = 0.10 * gl(rpred, fpred, None, None, for_discriminator=False)
                g_per = 1.0 * P((fake + 1.) / 2., (raw + 1.) / 2.)
                idt = G(exp)
                g_idt = 0.10 * ms(idt, exp)
                g_loss = g_adv + g_per + g_idt
                g_loss.backward()
                if step == 0:
                    out["step_grad_enc1_w"] = G.enc1.main[1].weight.grad.numpy().copy()
                    out["step_grad_dec5_1_w"] = G.dec5[1].main[1].weight.grad.numpy().copy()
                    out["step_grad_ga3_fuse_w_norm"] = np.float32(G.ga3.fuse[0].weight.grad.norm().item())
                g_opt.step()
                out[f"step{step}_losses"] = np.array([d_loss.item(), g_adv.item(), g_per.item(), g_idt.item(),
                                                     g_loss.item()], dtype=np.float64)
            gsd, dsd = G.state_dict(), D.state_dict()
            out["step_post_enc1_w"] = gsd["enc1.main.1.weight"].numpy()
            out["step_post_dec5_1_w"] = gsd["dec5.1.main.1.weight"].numpy()
            out["step_post_dec5_1_b"] = gsd["dec5.1.main.1.bias"].numpy()
            out["step_post_d1_w"] = dsd["d1.0.1.weight_orig"].numpy()
            out["step_post_d1_u"] = dsd["d1.0.1.weight_u"].numpy()
            out["step_post_d5_pred_w"] = dsd["d5_pred.0.1.weight"].numpy()
            out["step_post_g_sum"] = np.float64(sum(float(v.double().sum()) for v in gsd.values()))
            out["step_post_d_sum"] = np.float64(sum(float(v.double().sum()) for v in dsd.values()))
        np.savez_compressed(os.path.join(OUT, f"golden_{regime}.npz"), **out)
        print(regime, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
