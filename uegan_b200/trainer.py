"""Drop-in `trainer` module: `Trainer(loaders, args).train()` with the reference's step semantics
(/root/reference/trainer.py:19-145, 313-354) on the native models / losses.

The loop body of the reference (trainer.py:75-119) is factored into `train_step(real_raw, real_exp)`; everything the
reference does around it that is not on the hot path (tensorboard, PSNR/SSIM/NIMA validation, sample dumps) is not
re-implemented here -- use the reference's own trainer with `dropin/models.py` + `dropin/losses.py` for those.

Data parallel: one process per GPU, batch sharded.  Each network's gradients live in ONE flat fp32 bucket in
symmetric (peer-mapped) memory; `uegan_adam_step_peers` sums the ranks' buckets in rank order while it applies Adam (D before
`d_optimizer.step()`, G before `g_optimizer.step()`), and the 16 / 32 doubles per GAN-loss evaluation that keep the
relativistic means global go through `uegan_peer_sum_f64` -- no NCCL call on the step, so it is captured as one CUDA graph at
any world size.  NCCL all-reduces of the same buckets are the fallback (`peer_reduce=False`).
"""
from __future__ import annotations

import os
import random
import time

import torch
import torch.nn as nn

from . import kernels as K
from .losses import GANLoss, MultiscaleRecLoss, PerceptualLoss
from .models import Discriminator, Generator


class ImagePool:
    """History buffer of generated images (utils.py:23-50); pool_size=0 returns its input."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.num_imgs, self.images = 0, []

    def query(self, images):
        if self.pool_size == 0:
            return images
        out = []
        for image in images:
            image = torch.unsqueeze(image.data, 0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                out.append(image)
            elif random.uniform(0, 1) > 0.5:
                idx = random.randint(0, self.pool_size - 1)
                out.append(self.images[idx].clone())
                self.images[idx] = image
            else:
                out.append(image)
        return torch.cat(out, 0)


class _FlatGrads:
    """All gradients of a module in ONE flat fp32 bucket (p.grad are views), so a step needs one all-reduce (RMSprop
    branch; the Adam branch uses uegan_b200.optim.FlatBucket / FlatAdam)."""

    def __init__(self, module, group=None):
        self.group = group
        self.params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_grad(self):
        self.flat.zero_()

    def all_reduce(self):
        if self.group is not None:
            import torch.distributed as dist
            dist.all_reduce(self.flat, group=self.group)


def init_weights(net, init_type="orthogonal", gain=0.02):
    """trainer.py:357-390 (class-name matching on 'Conv')."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                nn.init.normal_(m.weight.data, 0.0, gain)
            elif init_type == "xavier":
                nn.init.xavier_normal_(m.weight.data, gain=gain)
            elif init_type == "xavier_uniform":
                nn.init.xavier_uniform_(m.weight.data, gain=1.0)
            elif init_type == "kaiming":
                nn.init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "kaiming_uniform":
                nn.init.kaiming_uniform_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                nn.init.orthogonal_(m.weight.data, gain=gain)
            elif init_type == "none":
                m.reset_parameters()
            else:
                raise NotImplementedError("Initialization method [{}] is not implemented".format(init_type))
            if hasattr(m, "bias") and m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)
    net.apply(init_func)


class Trainer(object):
    def __init__(self, loaders, args, process_group=None, vgg_state_dict=None):
        self.loaders, self.args = loaders, args
        if not torch.cuda.is_available():
            raise RuntimeError("uegan_b200.trainer.Trainer needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.group = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.model_save_path = os.path.join(args.save_root_dir, args.version, args.model_save_path)
        self.vgg_state_dict = vgg_state_dict
        self.build_model()

    # ------------------------------------------------------------------ trainer.py:313-354
    def build_model(self):
        a = self.args
        self.G = Generator(a.g_conv_dim, a.g_norm_fun, a.g_act_fun, a.g_use_sn).to(self.device)
        self.D = Discriminator(a.d_conv_dim, a.d_norm_fun, a.d_act_fun, a.d_use_sn, a.adv_loss_type).to(self.device)
        if a.init_type:
            init_weights(self.G, a.init_type, 0.02)
            init_weights(self.D, a.init_type, 0.02)
        if self.group is not None:  # identical replicas: broadcast rank 0's weights and spectral-norm buffers once
            import torch.distributed as dist
            for t in list(self.G.state_dict().values()) + list(self.D.state_dict().values()):
                dist.broadcast(t, src=0, group=self.group)
        # Data-parallel reductions over peer memory (NVLink / NVSwitch): gradient buckets and the GAN-loss partial sums
        # live in symmetric memory and are summed inside our own kernels; NCCL remains the fallback
        # (args.peer_reduce=False, or no symmetric-memory support).
        self.comm = None
        if self.group is not None and getattr(a, "peer_reduce", True):
            try:
                from .peer import PeerComm
                self.comm = PeerComm(self.group)
            except Exception as e:  # noqa: BLE001
                import warnings
                warnings.warn(f"uegan_b200: peer-memory reductions unavailable ({type(e).__name__}: {e}); using NCCL")
        if a.optimizer_type == "adam":
            from .optim import FlatAdam, FlatBucket
            galloc = self.comm.alloc if self.comm is not None else None
            self.g_grads, self.d_grads = FlatBucket(self.G, galloc), FlatBucket(self.D, galloc)

            def opt(bucket, lr, net):
                before = after = peers = None
                if self.comm is not None:
                    peers, before, after = self.comm.peer_ptrs(bucket.grad), self.comm.barrier, self.comm.barrier
                elif self.group is not None:
                    import torch.distributed as dist
                    before = lambda: dist.all_reduce(bucket.grad, group=self.group)
                return FlatAdam(bucket, lr=lr, betas=[a.beta1, a.beta2], weight_decay=0.0001, peers=peers,
                                before_step=before, after_step=after, on_update=net._wcache.bump)
            self.g_optimizer, self.d_optimizer = opt(self.g_grads, a.g_lr, self.G), opt(self.d_grads, a.d_lr, self.D)
        elif a.optimizer_type == "rmsprop":
            # non-default branch (trainer.py:340-342): torch's RMSprop on flat gradient buckets, NCCL all-reduce
            self.g_grads, self.d_grads = _FlatGrads(self.G, self.group), _FlatGrads(self.D, self.group)
            self.g_optimizer = torch.optim.RMSprop(params=self.G.parameters(), lr=a.g_lr, alpha=a.alpha)
            self.d_optimizer = torch.optim.RMSprop(params=self.D.parameters(), lr=a.d_lr, alpha=a.alpha)
        else:
            raise NotImplementedError("=== Optimizer [{}] is not found ===".format(a.optimizer_type))
        if a.lr_decay:
            rule = lambda epoch: 1.0 - max(0, epoch + 1 - a.lr_num_epochs_decay) / a.lr_decay_ratio
            self.lr_scheduler_g = torch.optim.lr_scheduler.LambdaLR(self.g_optimizer, lr_lambda=rule)
            self.lr_scheduler_d = torch.optim.lr_scheduler.LambdaLR(self.d_optimizer, lr_lambda=rule)
        self.fake_exp_pool = ImagePool(a.pool_size)
        self.criterionPercep = PerceptualLoss(self.vgg_state_dict).to(self.device)
        self.criterionIdt = MultiscaleRecLoss(scale=3, rec_loss_type=a.idt_loss_type, multiscale=True)
        self.criterionGAN = GANLoss(a.adv_loss_type, tensor=torch.cuda.FloatTensor)
        self.criterionGAN.process_group = self.comm if self.comm is not None else self.group

    # ------------------------------------------------------------------ trainer.py:75-119
    def train_step(self, real_raw, real_exp, sync_scalars=True):
        """One iteration on this rank's shard.  Returns the five loss values (python floats when sync_scalars, as
        the reference's `.item()` calls; 0-d CUDA tensors otherwise, which keeps the step free of host syncs)."""
        a, gan = self.args, self.criterionGAN
        self.G.train(); self.D.train()
        self.real_raw, self.real_exp = real_raw, real_exp
        local = 1.0 / self.world  # mean-type losses are per-rank means; gradients are SUMMED across ranks
        # G(real_raw) (trainer.py:80) and G(real_exp) (the identity pass, trainer.py:106) see the same weights -- G is only
        # updated at the end of the step -- and every layer of G is per-sample (InstanceNorm, GAM pooling): one forward /
        # backward over the 2B images gives the same two results with half of G's kernel launches (UEGAN_BATCH_G=0: two
        # passes, in the reference's order)
        batch_g = os.environ.get("UEGAN_BATCH_G", "1") != "0" and real_raw.shape == real_exp.shape
        if batch_g:
            nb = real_raw.shape[0]
            if getattr(self, "_gxy", None) is not None and real_raw is self._gx and real_exp is self._gy:
                both_in = self._gxy  # the static input buffer of the captured step holds the two batches back to back
            else:
                both_in = torch.cat([real_raw, real_exp])
            both = self.G(both_in)
            self.fake_exp, real_exp_idt = both[:nb], both[nb:]
        else:
            self.fake_exp = self.G(real_raw)
        self.fake_exp_store = self.fake_exp_pool.query(self.fake_exp)
        # ---- update D
        self.d_grads.zero_grad()
        real_exp_preds = self.D(real_exp)
        fake_exp_preds = self.D(self.fake_exp_store.detach())
        d_loss = gan(real_exp_preds, fake_exp_preds, None, None, for_discriminator=True)
        if a.adv_input:
            input_preds = self.D(real_raw)
            d_loss = d_loss + gan(real_exp_preds, input_preds, None, None, for_discriminator=True)
        d_loss.backward()
        if isinstance(self.d_grads, _FlatGrads):
            self.d_grads.all_reduce()
        self.d_optimizer.step()  # FlatAdam: gradient reduction (peer memory or NCCL) + Adam, trainer.py:97
        # ---- update G
        self.g_grads.zero_grad()
        # The reference lets g_loss.backward() also fill D's parameter gradients and then discards them at the next
        # d_optimizer.zero_grad() (trainer.py:89).  Same result without the wasted weight-gradient
        # GEMMs: D's parameters are frozen for these two forwards (the gradient w.r.t. fake_exp still flows).
        d_params = [p for p in self.D.parameters() if p.requires_grad]
        for p in d_params:
            p.requires_grad_(False)
        try:
            real_exp_preds = self.D(real_exp)
            fake_exp_preds = self.D(self.fake_exp)
        finally:
            for p in d_params:
                p.requires_grad_(True)
        g_adv_loss = a.lambda_adv * gan(real_exp_preds, fake_exp_preds, None, None, for_discriminator=False)
        g_percep_loss = a.lambda_percep * self.criterionPercep((self.fake_exp + 1.) / 2., (real_raw + 1.) / 2.)
        self.real_exp_idt = real_exp_idt if batch_g else self.G(real_exp)
        g_idt_loss = a.lambda_idt * self.criterionIdt(self.real_exp_idt, real_exp)
        g_loss = g_adv_loss + local * (g_percep_loss + g_idt_loss)
        g_loss.backward()
        if isinstance(self.g_grads, _FlatGrads):
            self.g_grads.all_reduce()
        self.g_optimizer.step()
        vals = dict(d_loss=d_loss, g_adv_loss=g_adv_loss, g_percep_loss=g_percep_loss, g_idt_loss=g_idt_loss,
                    g_loss=g_adv_loss + g_percep_loss + g_idt_loss)
        if sync_scalars:
            vals = {k: float(v.detach()) for k, v in vals.items()}
            for k, v in vals.items():
                setattr(self, k, v)
        return vals

    # ------------------------------------------------------------------ whole step as a CUDA graph
    def capture(self, real_raw, real_exp, warmup=3):
        """Captures train_step (2 G + 5 D + 2 VGG forwards, all backwards, both fused reduce+Adam steps) into one CUDA
        graph.  ~1000 kernel launches per step otherwise cost more host time than the GPU needs to run them.  With the
        peer-memory reductions there is no NCCL call inside the step, so the capture works at any world size.
        Requires pool_size == 0 (ImagePool is host logic).
        Afterwards `replay(real_raw, real_exp)` copies the batch into the static input buffers and launches the graph;
        the five losses come back as 0-d CUDA tensors (no host sync)."""
        assert self.args.pool_size == 0
        if self.group is not None and self.comm is None:
            raise RuntimeError("CUDA-graph capture at world size > 1 needs the peer-memory reductions (NCCL fallback active)")
        if real_raw.shape == real_exp.shape:
            nb = real_raw.shape[0]
            self._gxy = torch.cat([real_raw, real_exp])  # one buffer: the batched G pass reads it without a copy
            self._gx, self._gy = self._gxy[:nb], self._gxy[nb:]
        else:
            self._gxy = None
            self._gx, self._gy = real_raw.clone(), real_exp.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # allocates every workspace / scratch buffer and settles the weight-pack caches
                self.train_step(self._gx, self._gy, sync_scalars=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        l0 = K.launches()
        with torch.cuda.graph(self._graph):
            self._gout = self.train_step(self._gx, self._gy, sync_scalars=False)
        self.graph_launches = K.launches() - l0  # kernels of libuegan_sm100.so inside one replay
        return self._graph

    def prefetch(self, real_raw_host, real_exp_host):
        """Starts the host -> device copy of the NEXT batch (pinned host tensors) on a copy stream, into staging buffers;
        the following `replay(None, None)` consumes it.  Called right after a replay, the copy overlaps that step's graph
        (the input pipeline of a DataLoader with pin_memory, one batch ahead)."""
        if getattr(self, "_sx", None) is None:
            self._sx, self._sy = torch.empty_like(self._gx), torch.empty_like(self._gy)
            self._copy_stream = torch.cuda.Stream()
            self._copy_done, self._stage_free = torch.cuda.Event(), torch.cuda.Event()
            self._stage_free.record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free)  # the previous replay has copied the staging buffers out
            self._sx.copy_(real_raw_host, non_blocking=True)
            self._sy.copy_(real_exp_host, non_blocking=True)
            self._copy_done.record()
        self._staged = True

    def replay(self, real_raw, real_exp, sync_scalars=False):
        """real_raw / real_exp = None: use the batch staged by `prefetch`."""
        for o in (self.g_optimizer, self.d_optimizer):
            if hasattr(o, "sync_lr"):
                o.sync_lr()  # a scheduler may have changed the learning rate: host -> device scalar, outside the graph
        if real_raw is None:
            assert getattr(self, "_staged", False), "replay(None, None) needs a prefetch() first"
            torch.cuda.current_stream().wait_event(self._copy_done)
            self._gx.copy_(self._sx, non_blocking=True)
            self._gy.copy_(self._sy, non_blocking=True)
            self._stage_free.record()
            self._staged = False
        else:
            self._gx.copy_(real_raw, non_blocking=True)
            self._gy.copy_(real_exp, non_blocking=True)
        self._graph.replay()
        K._count(self.graph_launches)
        if sync_scalars:
            # one device -> host read for the five scalars (the reference's five .item() calls, trainer.py:98-119)
            keys = list(self._gout.keys())
            vals = torch.stack([self._gout[k].detach().reshape(()).float() for k in keys]).tolist()
            return dict(zip(keys, vals))
        return self._gout

    # ------------------------------------------------------------------ trainer.py:40-145 (hot loop only)
    def train(self):
        a = self.args
        loader = self.loaders.ref
        steps_per_epoch = len(loader)
        total_steps = int(a.total_epochs * steps_per_epoch)
        start_step = 0
        if a.pretrained_model:
            start_step = int(a.pretrained_model * steps_per_epoch)
            self.load_pretrained_model(a.pretrained_model)
        save_step = max(1, int(a.model_save_epoch * steps_per_epoch))
        it = iter(loader)
        t0 = time.time()
        for step in range(start_step, total_steps):
            try:
                x, y, _ = next(it)
            except StopIteration:
                it = iter(loader)
                x, y, _ = next(it)
            real_exp, real_raw = x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True)
            vals = self.train_step(real_raw, real_exp)
            if (step + 1) % a.info_step == 0:
                print("Elapse:{:>.8s}, D_Step:{:>6d}/{}, G_Step:{:>6d}/{}, D_loss:{:>.4f}, G_loss:{:>.4f}, "
                      "G_percep_loss:{:>.4f}, G_adv_loss:{:>.4f}, G_idt_loss:{:>.4f}".format(
                          str(time.time() - t0), step + 1, total_steps, step + 1, total_steps, vals["d_loss"],
                          vals["g_loss"], vals["g_percep_loss"], vals["g_adv_loss"], vals["g_idt_loss"]))
            if (step + 1) % save_step == 0:
                self.save_checkpoint((step + 1) / steps_per_epoch)  # float epoch, trainer.py:163,205
            if a.lr_decay and step % steps_per_epoch == 0:
                epoch = step // steps_per_epoch
                self.lr_scheduler_g.step(epoch=epoch)
                self.lr_scheduler_d.step(epoch=epoch)

    # ------------------------------------------------------------------ trainer.py:186-210, 402-423
    def _ckpt_path(self, epoch):
        return os.path.join(self.model_save_path, "{}_{}_{}.pth".format(self.args.version, self.args.adv_loss_type, epoch))

    def save_checkpoint(self, epoch):
        os.makedirs(self.model_save_path, exist_ok=True)
        ck = {"G_net": self.G.state_dict(), "D_net": self.D.state_dict(), "epoch": epoch,
              "g_optimizer": self.g_optimizer.state_dict(), "d_optimizer": self.d_optimizer.state_dict()}
        if self.args.lr_decay:
            ck["lr_scheduler_g"] = self.lr_scheduler_g.state_dict()
            ck["lr_scheduler_d"] = self.lr_scheduler_d.state_dict()
        torch.save(ck, self._ckpt_path(epoch))

    def load_pretrained_model(self, resume_epochs):
        ck = torch.load(self._ckpt_path(resume_epochs), map_location=self.device)
        self.G.load_state_dict(ck["G_net"])
        self.D.load_state_dict(ck["D_net"])
        self.g_optimizer.load_state_dict(ck["g_optimizer"])
        self.d_optimizer.load_state_dict(ck["d_optimizer"])
        if self.args.lr_decay and "lr_scheduler_g" in ck:
            self.lr_scheduler_g.load_state_dict(ck["lr_scheduler_g"])
            self.lr_scheduler_d.load_state_dict(ck["lr_scheduler_d"])
