"""`FusedAdamW` — torch.optim.AdamW semantics (the optimizer every reference config names,
configs/demo_training_lycoris.yaml:50; constructed at src/duwu/trainer/trainer.py:52-74) executed as ONE multi-tensor
kernel over all parameters, with Lightning's `gradient_clip_val` (configs/demo_training_lycoris.yaml:13) folded in:
the global L2 norm and the clip coefficient are computed on the device and consumed by the update kernel without a host
synchronisation.  State-dict layout (`state[p] = {step, exp_avg, exp_avg_sq}`) matches torch.optim.AdamW.
"""
from __future__ import annotations

import math
from typing import Iterable, Optional

import torch

from ._lib import check, lib

_CHUNK = 1 << 16


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 amsgrad: bool = False, *, max_grad_norm: Optional[float] = None, maximize: bool = False,
                 foreach: Optional[bool] = None, capturable: bool = False, differentiable: bool = False,
                 fused: Optional[bool] = None):
        # torch.optim.AdamW's keyword set; the ones that only select an ATen code path are accepted, the ones that change the
        # update rule are refused instead of ignored
        if amsgrad or maximize or differentiable:
            raise NotImplementedError("FusedAdamW: amsgrad / maximize / differentiable are not built")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self._tables = None
        self._step = 0
        self.last_norm = None  # device tensor {norm, clip coefficient} of the latest step

    # ---- device tables (built once; parameters, gradients and moments must keep their storage) ----------------
    def _build(self):
        groups = []
        for gi, g in enumerate(self.param_groups):
            ps = [p for p in g["params"] if p.requires_grad]
            if not ps:
                continue
            dev = ps[0].device
            for p in ps:
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise ValueError("FusedAdamW needs contiguous fp32 parameters (master weights)")
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
            i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
            chunk_t, chunk_i = [], []
            for ti, p in enumerate(ps):
                n = (p.numel() + _CHUNK - 1) // _CHUNK
                chunk_t += [ti] * n
                chunk_i += list(range(n))
            groups.append(dict(
                gi=gi, params=ps, device=dev,
                p=i64([p.data_ptr() for p in ps]), g=i64([p.grad.data_ptr() for p in ps]),
                m=i64([self.state[p]["exp_avg"].data_ptr() for p in ps]),
                v=i64([self.state[p]["exp_avg_sq"].data_ptr() for p in ps]),
                numel=i64([p.numel() for p in ps]),
                ct=torch.tensor(chunk_t, dtype=torch.int32, device=dev), ci=torch.tensor(chunk_i, dtype=torch.int32, device=dev),
                n_chunks=len(chunk_t), grad_ptrs=[p.grad.data_ptr() for p in ps],
                partial=torch.empty((len(chunk_t),), dtype=torch.float32, device=dev),
            ))
        self._tables = groups
        self._norm_out = torch.ones((2,), dtype=torch.float32, device=groups[0]["device"]) if groups else None
        # the clip coefficient is global (Lightning clips the norm over ALL optimizer parameters): one norm table over
        # every group
        self._norm_table = groups[0] if len(groups) == 1 else None
        if len(groups) > 1:
            dev = groups[0]["device"]
            ps = [p for g in groups for p in g["params"]]
            chunk_t, chunk_i = [], []
            for ti, p in enumerate(ps):
                n = (p.numel() + _CHUNK - 1) // _CHUNK
                chunk_t += [ti] * n
                chunk_i += list(range(n))
            self._norm_table = dict(
                g=torch.tensor([p.grad.data_ptr() for p in ps], dtype=torch.int64, device=dev),
                numel=torch.tensor([p.numel() for p in ps], dtype=torch.int64, device=dev),
                ct=torch.tensor(chunk_t, dtype=torch.int32, device=dev), ci=torch.tensor(chunk_i, dtype=torch.int32, device=dev),
                n_chunks=len(chunk_t), partial=torch.empty((len(chunk_t),), dtype=torch.float32, device=dev))
        # a rebuild (gradient storage replaced) must not restart the bias-correction counter
        self._step = max([getattr(self, "_step", 0)] + [int(self.state[p]["step"]) for g in groups for p in g["params"]])

    def _check_storage(self):
        for t in self._tables:
            for p, gp in zip(t["params"], t["grad_ptrs"]):
                if p.grad is None or p.grad.data_ptr() != gp:
                    self._tables = None
                    return

    def zero_grad(self, set_to_none: bool = False):
        # gradients keep their storage (flat buckets for the all-reduce): always zero in place
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    p.grad.zero_()

    def hyper_values(self, group: int = 0):
        """{lr, 1 - beta1^t, sqrt(1 - beta2^t)} of the NEXT step for a param group (what a graph-captured step reads from
        device memory instead of taking by value)."""
        g = self.param_groups[group]
        # the C ABI takes the betas as fp32 and forms the corrections in double from those (uwu_mt_adamw): same values here
        b1, b2 = (torch.tensor(b, dtype=torch.float32).item() for b in g["betas"])
        t = self._step + 1
        return [float(g["lr"]), 1.0 - math.pow(b1, t), math.sqrt(1.0 - math.pow(b2, t))]  # same libm calls as the C side

    @torch.no_grad()
    def step(self, closure=None, hyper_dev=None):
        """`hyper_dev`: optional list (one per param group that owns tables) of device fp32[3] tensors holding
        `hyper_values()`; the kernels then read lr / bias corrections from device memory (CUDA-graph capture)."""
        loss = closure() if closure is not None else None
        if self._tables is not None:
            self._check_storage()
        if self._tables is None:
            self._build()
        if not self._tables:
            return loss
        stream = torch.cuda.current_stream().cuda_stream
        L = lib()
        clip = None
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            t = self._norm_table
            check(L.uwu_mt_gradnorm(t["g"].data_ptr(), t["numel"].data_ptr(), t["ct"].data_ptr(), t["ci"].data_ptr(),
                                    t["n_chunks"], _CHUNK, float(self.max_grad_norm), t["partial"].data_ptr(),
                                    self._norm_out.data_ptr(), stream), "uwu_mt_gradnorm")
            clip = self._norm_out
            self.last_norm = self._norm_out
        self._step += 1
        for ti, t in enumerate(self._tables):
            g = self.param_groups[t["gi"]]
            b1, b2 = g["betas"]
            hd = hyper_dev[ti].data_ptr() if hyper_dev is not None else None
            check(L.uwu_mt_adamw(t["p"].data_ptr(), t["g"].data_ptr(), t["m"].data_ptr(), t["v"].data_ptr(),
                                 t["numel"].data_ptr(), t["ct"].data_ptr(), t["ci"].data_ptr(), t["n_chunks"], _CHUNK,
                                 float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                                 self._step, clip.data_ptr() if clip is not None else None, hd, stream), "uwu_mt_adamw")
        return loss

    def state_dict(self):
        # per-parameter `step` tensors are materialised lazily (one shared counter drives the kernel)
        for t in self._tables or []:
            for p in t["params"]:
                self.state[p]["step"] = torch.tensor(float(self._step))
        return super().state_dict()

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        self._tables = None
        self._step = 0  # re-read from the loaded per-parameter `step` entries

    def launches_per_step(self) -> int:
        n = len(self._tables or [])
        return n + (2 if self.max_grad_norm else 0)
