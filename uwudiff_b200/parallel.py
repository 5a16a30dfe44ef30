"""Data-parallel gradient exchange: the only collective on the path.

The reference gets it implicitly from Lightning `strategy: ddp` (commented in configs/demo_training_lycoris.yaml:6-8;
per-rank seeding test_scripts/test_train.py:68-69): torch DDP all-reduces 25 MB gradient buckets during backward.  Here,
one process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch): all trainable gradients live in ONE flat fp32
buffer laid out in module order; the hand-scheduled backward reports each finished top-level block
(`unet.after_backward`), and every bucket whose gradients are complete is averaged with `all_reduce` on a side stream
while the earlier blocks are still in backward.  The optimizer waits on the side stream, not the host.
In LyCORIS mode only the ≈52.4 M adapter parameters are exchanged (≈210 MB fp32 per step).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class GradientBuckets:
    def __init__(self, trainer, process_group=None, n_buckets: int = 0):
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.backend = dist.get_backend(process_group)
        self.unet = trainer.unet
        ly = trainer.lycoris_model
        self._sync_replicas(trainer)
        if ly is not None:
            self.flat = ly.flat_grads
            params = list(ly.parameters())
        else:
            params = [p for p in self.unet.parameters() if p.requires_grad]
            total = sum((p.numel() + 3) // 4 * 4 for p in params)
            self.flat = torch.zeros((total,), device=params[0].device, dtype=torch.float32)
            off = 0
            for p in params:
                p.grad = self.flat[off:off + p.numel()].view(p.shape)
                off += (p.numel() + 3) // 4 * 4
        base = self.flat.data_ptr()
        offs = {id(p): (p.grad.data_ptr() - base) // 4 for p in params}
        # first flat offset owned by each top-level block, in forward (= layout) order
        self._block_start = {}
        tops = self.unet.ddp_blocks()  # the units the hand-scheduled backward reports through `after_backward`
        for top in tops:
            mine = []
            for m in top.modules():
                ad = getattr(m, "_uwu_adapter", None)
                if ad is not None:
                    mine += [offs[id(p)] for p in ad.parameters() if id(p) in offs]
                mine += [offs[id(p)] for p in m.parameters(recurse=False) if id(p) in offs]
            if mine:
                self._block_start[id(top)] = min(mine)
        n = self.flat.numel()
        if n_buckets <= 0:
            # ~64 MiB per exchange: the LyCORIS buffer (210 MB) goes out in 4 pieces, the full fine-tuning buffers (DiT 2.7 GB,
            # SD-1.5 3.4 GB) in 32, so that only a small last bucket is left to run after the backward pass has finished
            n_buckets = max(4, min(32, (n * 4 + (64 << 20) - 1) // (64 << 20)))
        n_buckets = max(1, min(n_buckets, n // 1024 or 1))
        step = (n + n_buckets - 1) // n_buckets
        step = (step + 255) // 256 * 256
        self.bounds = [(s, min(n, s + step)) for s in range(0, n, step)]
        self.cuda = self.flat.is_cuda
        self.side = torch.cuda.Stream() if self.cuda else None
        self._ready = n
        self._launched: List[bool] = []
        self._works = []
        self.unet.after_backward = self._on_blocks_done
        self.reduced_elems = 0
        self.enabled = True  # False while a gradient-accumulation micro-batch other than the last is in backward

    def _sync_replicas(self, trainer):
        """What torch DDP does at construction and the reference relies on (per-rank `seed + rank`,
        test_scripts/test_train.py:68-69): parameters AND buffers of rank 0 — frozen base, adapters, loss head — are broadcast
        so every replica starts from the same state whatever its local seed was; the bf16 operand caches are dropped."""
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        ly = trainer.lycoris_model
        dev = (ly.flat_params if ly is not None else next(self.unet.parameters())).device

        def bcast(t):
            # host-resident buffers (the adapters' `alpha` scalars, scheduler tables) travel through the device the process
            # group's backend serves (NCCL has no CPU tensors)
            if t.device != dev and dev.type == "cuda":
                tmp = t.to(dev)
                dist.broadcast(tmp, src=src, group=self.group)
                t.copy_(tmp)
            else:
                dist.broadcast(t, src=src, group=self.group)

        with torch.no_grad():
            if ly is not None:
                bcast(ly.flat_params)  # every adapter tensor is a view into it
                for b in ly.buffers():
                    bcast(b)
            mods = [self.unet] + ([trainer.loss] if isinstance(getattr(trainer, "loss", None), torch.nn.Module) else [])
            for mod in mods:
                for t in list(mod.parameters()) + list(mod.buffers()):
                    if t.is_floating_point() or t.dtype in (torch.int64, torch.int32):
                        bcast(t.data)
        if hasattr(self.unet, "refresh_weights"):
            self.unet.refresh_weights()

    def begin_step(self):
        self._ready = self.flat.numel()
        self._launched = [False] * len(self.bounds)
        self._works = []
        self.reduced_elems = 0

    def _on_blocks_done(self, mods):
        if not self.enabled:
            return
        for m in mods:
            s = self._block_start.get(id(m))
            if s is not None:
                self._ready = min(self._ready, s)
        self._launch_ready(self._ready)

    def _launch_ready(self, ready: int):
        for i in reversed(range(len(self.bounds))):
            lo, hi = self.bounds[i]
            if self._launched[i] or lo < ready:
                continue
            self._launched[i] = True
            view = self.flat[lo:hi]
            self.reduced_elems += hi - lo
            if self.cuda:
                ev = torch.cuda.Event()
                ev.record()
                self.side.wait_event(ev)
                with torch.cuda.stream(self.side):
                    self._reduce(view)
            else:
                self._reduce(view)

    def _reduce(self, view):
        if self.backend == "nccl":
            dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            view.div_(self.world)

    def finish(self):
        """All remaining buckets, then make the compute stream wait for the exchange (no host sync)."""
        self._launch_ready(0)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.side)
