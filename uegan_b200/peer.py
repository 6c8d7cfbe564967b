"""Peer-memory plumbing of the data-parallel step: gradient buckets and the GAN-loss partial sums live in
torch symmetric memory (one allocation per rank, mapped into every rank's address space over NVLink / NVSwitch), so the
reductions are plain loads inside OUR kernels (uegan_adam_step_peers, uegan_peer_sum_f64) ordered by device-side barriers
-- no NCCL collective on the step, and the whole step is CUDA-graph capturable at any world size.

torch.distributed is used for what it is here: process-group rendezvous (exchange of the memory handles, once)."""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib as L
from . import kernels as K


class PeerComm:
    STAGE_SLOTS = 8   # rotating slots of 64 doubles for the small reductions (reused only after >= 1 barrier, see reduce())
    SLOT = 64

    def __init__(self, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._symm = symm_mem
        self._handles = []
        self.stage = symm_mem.empty(self.STAGE_SLOTS * self.SLOT, dtype=torch.float64, device="cuda")
        self.stage.zero_()
        self._stage_hdl = symm_mem.rendezvous(self.stage, group)
        self._stage_ptrs = [int(p) for p in self._stage_hdl.buffer_ptrs]
        self._slot = 0
        self._barrier_channel = 0
        torch.cuda.synchronize()
        dist.barrier(group)

    def alloc(self, numel: int, dtype=torch.float32) -> torch.Tensor:
        """A zeroed symmetric buffer; `peer_ptrs(t)` gives its address on every rank (rank order)."""
        t = self._symm.empty(numel, dtype=dtype, device="cuda")
        t.zero_()
        hdl = self._symm.rendezvous(t, self.group)
        self._handles.append((t, hdl))
        return t

    def peer_ptrs(self, t: torch.Tensor) -> List[int]:
        for buf, hdl in self._handles:
            if buf.data_ptr() == t.data_ptr():
                return [int(p) for p in hdl.buffer_ptrs]
        raise KeyError("tensor was not allocated by PeerComm.alloc")

    def barrier(self):
        """Device-side barrier across the ranks on the current stream (orders peer-memory reads after peers' writes)."""
        self._stage_hdl.barrier(channel=self._barrier_channel)

    def reduce(self, ws: torch.Tensor, lo: int, hi: int):
        """ws[lo:hi] (float64, local) <- sum over ranks of ws[lo:hi], in rank order (bit-identical on every rank)."""
        n = hi - lo
        assert n <= self.SLOT and ws.dtype == torch.float64
        slot = self._slot
        self._slot = (self._slot + 1) % self.STAGE_SLOTS
        off = slot * self.SLOT
        # publish this rank's partials, wait for everybody's, sum them.  A slot is written again only STAGE_SLOTS reductions
        # later; every rank has passed at least one barrier in between, i.e. has finished reading it.
        self.stage[off:off + n].copy_(ws[lo:hi])
        self.barrier()
        arr = (C.c_void_p * self.world)(*[p + off * 8 for p in self._stage_ptrs])
        L.check(L.load().uegan_peer_sum_f64(ws.data_ptr() + lo * 8, arr, self.world, n, K._stream()), "peer_sum_f64")
        K._count(1, "peer_sum_f64")
