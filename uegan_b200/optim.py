"""Flat parameter / gradient buckets and the fused Adam of the native Trainer (reference: trainer.py:337-338,
`torch.optim.Adam(params, lr, betas=[beta1, beta2], weight_decay=0.0001)`).

`FlatBucket` re-homes every parameter of a network into ONE contiguous fp32 buffer (`p.data` become views, so
`state_dict()` / `load_state_dict()` / checkpoints are unchanged) and gives every parameter a `.grad` view into one
gradient bucket; the weight-gradient kernels accumulate straight into those views (`module._grad_sink`).
`FlatAdam` is a `torch.optim.Adam` subclass -- same constructor, `param_groups`, `state_dict()` layout, so reference
checkpoints load and LambdaLR drives it -- whose `step()` is ONE launch of `uegan_adam_step` over the bucket, or of
`uegan_adam_step_peers`, which sums the ranks' gradient buckets out of peer memory (NVLink / NVSwitch) while it applies
the update: the data-parallel all-reduce and the optimizer are one kernel, no NCCL call on the step.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional

import torch

from . import _lib as L
from . import kernels as K


class FlatBucket:
    def __init__(self, module: torch.nn.Module, grad_alloc: Optional[Callable[[int], torch.Tensor]] = None):
        self.module = module
        self.names = [n for n, p in module.named_parameters() if p.requires_grad]
        self.params = [p for _, p in module.named_parameters() if p.requires_grad]
        dev = self.params[0].device
        # every parameter starts on a 16-byte boundary (float4 accesses in the Adam kernel, TMA-free but vector friendly)
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = grad_alloc(n) if grad_alloc is not None else torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad.zero_()
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[off:off + p.numel()].view_as(p)
        module._grad_sink = {name: p.grad for name, p in zip(self.names, self.params)}

    def zero_grad(self):
        self.grad.zero_()


class FlatAdam(torch.optim.Adam):
    """torch.optim.Adam over a FlatBucket.  `peers` (device addresses of all ranks' gradient buckets, rank order) selects
    the fused all-reduce + Adam kernel; `before_step` / `after_step` are the device-side barriers that order it against
    the other ranks' backward passes (uegan_b200.peer.PeerComm)."""

    def __init__(self, bucket: FlatBucket, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, peers: Optional[List[int]] = None,
                 before_step: Optional[Callable[[], None]] = None, after_step: Optional[Callable[[], None]] = None,
                 on_update: Optional[Callable[[], None]] = None):
        super().__init__(bucket.params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.bucket = bucket
        dev = bucket.flat.device
        self.exp_avg = torch.zeros_like(bucket.flat)
        self.exp_avg_sq = torch.zeros_like(bucket.flat)
        self.dev_state = torch.zeros(4, dtype=torch.float32, device=dev)  # {step, lr / bc1, 1 / sqrt(bc2)}
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        self.peers = list(peers) if peers else None
        self.before_step, self.after_step, self.on_update = before_step, after_step, on_update
        self._bind_state(step=0.0)

    # ---- torch.optim state in the reference's layout: views of the flat moments + a step scalar per parameter
    def _bind_state(self, step: float):
        b = self.bucket
        for p, off in zip(b.params, b.offsets):
            self.state[p] = {"step": torch.tensor(float(step), dtype=torch.float32),
                             "exp_avg": self.exp_avg[off:off + p.numel()].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off:off + p.numel()].view_as(p)}

    def sync_lr(self):
        """Host -> device copy of the learning rate when a scheduler changed it (outside any CUDA-graph capture)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        g = self.param_groups[0]
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        b = self.bucket
        lib = L.load()
        b1, b2 = g["betas"]
        if self.before_step is not None:
            self.before_step()
        if self.peers:
            arr = (C.c_void_p * len(self.peers))(*self.peers)
            L.check(lib.uegan_adam_step_peers(b.flat.data_ptr(), arr, len(self.peers), self.exp_avg.data_ptr(),
                                              self.exp_avg_sq.data_ptr(), b.numel, self.dev_state.data_ptr(),
                                              self.lr_dev.data_ptr(), b1, b2, g["eps"], g["weight_decay"], K._stream()),
                    "adam_step_peers")
        else:
            L.check(lib.uegan_adam_step(b.flat.data_ptr(), b.grad.data_ptr(), self.exp_avg.data_ptr(),
                                        self.exp_avg_sq.data_ptr(), b.numel, self.dev_state.data_ptr(),
                                        self.lr_dev.data_ptr(), b1, b2, g["eps"], g["weight_decay"], K._stream()),
                    "adam_step")
        K._count(2, "adam_step")
        if self.after_step is not None:
            self.after_step()
        if self.on_update is not None:
            self.on_update()  # the kernels wrote the parameters behind torch's back: invalidate the packed-weight caches

    def state_dict(self):
        step = float(self.dev_state[0].item())
        for st in self.state.values():
            st["step"] = torch.tensor(step, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)  # fills self.state[p] with the checkpoint's tensors (cast to p's device)
        b = self.bucket
        step = 0.0
        with torch.no_grad():
            for p, off in zip(b.params, b.offsets):
                st = self.state.get(p)
                if not st:
                    continue
                self.exp_avg[off:off + p.numel()].view_as(p).copy_(st["exp_avg"])
                self.exp_avg_sq[off:off + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
                step = max(step, float(st["step"]))
            self.dev_state.zero_()
            self.dev_state[0] = step
        self._bind_state(step)
        self._lr_host = None
        self.sync_lr()
