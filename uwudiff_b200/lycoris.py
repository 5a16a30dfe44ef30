"""LyCORIS wrapper protocol for the kernel-backed UNet: `LycorisNetwork.apply_preset`, `create_lycoris`, `.apply_to()`,
`.parameters()`, `.state_dict()`, `.restore()`, `.merge_to()` — the calls the reference trainer makes
(src/duwu/trainer/trainer.py:13,148-169,184-187,211-215) with the preset file configs/lycoris/sdxl-diffusers.toml.

The reference gets these from the un-vendored package `lycoris-lora>=3.0.1.dev10`; its bookkeeping is reproduced here
(adapter names `lycoris_<module path with '.' -> '_'>`, `factorization`, LoRA r/alpha, LoKr `full_matrix` factor shapes,
norm deltas, parameter names `lokr_w1/lokr_w2`, `lora_down.weight/lora_up.weight`, `w_norm/b_norm`, `alpha`).
Instead of patching `module.forward` with `F.linear(x, W + kron(w1, w2))`, `apply_to()` attaches the adapter to the
kernel-backed module: the delta is folded into the GEMM's bf16 operand on the device (uwu_fold_lokr / uwu_fold_lora) and
the adapter gradients are contracted out of the wgrad GEMM (uwu_lokr_grad / uwu_lora_grad).  All adapter parameters live
in ONE flat fp32 buffer (and their gradients in another) so the optimizer and the data-parallel all-reduce are single
passes over contiguous memory.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn as nn

from . import ops


def factorization(dimension: int, factor: int = -1):
    """(m, n), m <= n, m * n == dimension; `factor` preferred as m when it divides, else the closest divisor pair
    whose smaller member does not exceed `factor` (lycoris.functional.general.factorization, restated)."""
    if factor > 0 and (dimension % factor) == 0:
        m, n = factor, dimension // factor
        if m > n:
            n, m = m, n
        return m, n
    if factor < 0:
        factor = dimension
    m, n = 1, dimension
    length = m + n
    while m < n:
        new_m = m + 1
        while dimension % new_m != 0:
            new_m += 1
        new_n = dimension // new_m
        if new_m + new_n > length or new_m > factor:
            break
        m, n = new_m, new_n
    if m > n:
        n, m = m, n
    return m, n


class _Adapter(nn.Module):
    multiplier: float = 1.0

    def trainable(self) -> bool:
        return any(p.requires_grad for p in self.parameters())


class _W(nn.Module):
    def __init__(self, r, i):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(r, i))


class LoraLinear(_Adapter):
    def __init__(self, name: str, org: nn.Linear, multiplier: float, dim: int, alpha: float):
        super().__init__()
        self.lora_name, self.multiplier, self.dim = name, multiplier, dim
        self.lora_down = _W(dim, org.in_features)
        self.lora_up = _W(org.out_features, dim)
        self.register_buffer("alpha", torch.tensor(float(alpha)))
        self.scale = alpha / dim
        nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
        nn.init.constant_(self.lora_up.weight, 0)

    def fold_into(self, W, dst):
        ops.fold_lora(W, self.lora_up.weight, self.lora_down.weight, self.scale * self.multiplier, dst)

    def grads_from(self, G):
        from .unet import _grad_of

        ops.lora_grad(G, self.lora_up.weight, self.lora_down.weight, self.scale * self.multiplier,
                      _grad_of(self.lora_up.weight), _grad_of(self.lora_down.weight))

    def delta(self):
        return (self.lora_up.weight @ self.lora_down.weight) * self.scale


class LokrLinear(_Adapter):
    def __init__(self, name: str, org: nn.Linear, multiplier: float, dim: int, alpha: float, factor: int):
        super().__init__()
        self.lora_name, self.multiplier = name, multiplier
        in_m, in_n = factorization(org.in_features, factor)
        out_l, out_k = factorization(org.out_features, factor)
        self.shape = ((out_l, out_k), (in_m, in_n))
        self.lokr_w1 = nn.Parameter(torch.empty(out_l, in_m))
        self.lokr_w2 = nn.Parameter(torch.empty(out_k, in_n))
        self.register_buffer("alpha", torch.tensor(float(dim)))  # full_matrix: alpha := lora_dim -> scale 1
        self.scale = 1.0
        nn.init.constant_(self.lokr_w2, 0)
        nn.init.kaiming_uniform_(self.lokr_w1, a=math.sqrt(5))

    def fold_into(self, W, dst):
        ops.fold_lokr(W, self.lokr_w1, self.lokr_w2, dst, self.scale * self.multiplier)

    def grads_from(self, G):
        from .unet import _grad_of

        ops.lokr_grad(G, self.lokr_w1, self.lokr_w2, _grad_of(self.lokr_w1), _grad_of(self.lokr_w2),
                      self.scale * self.multiplier)

    def delta(self):
        return torch.kron(self.lokr_w1, self.lokr_w2.contiguous()) * self.scale


class NormDelta(_Adapter):
    def __init__(self, name: str, org: nn.Module, multiplier: float):
        super().__init__()
        self.lora_name, self.multiplier = name, multiplier
        dim = org.weight.shape[0]
        self.w_norm = nn.Parameter(torch.zeros(dim))
        self.b_norm = nn.Parameter(torch.zeros(dim))


class LycorisNetwork(nn.Module):
    ENABLE_CONV = True
    TARGET_MODULE: List[str] = ["Transformer2DModel", "ResnetBlock2D", "Downsample2D", "Upsample2D"]
    TARGET_NAME: List[str] = []
    MODULE_ALGO_MAP: Dict[str, dict] = {}
    LORA_PREFIX = "lycoris"

    @classmethod
    def apply_preset(cls, preset: dict):
        if "enable_conv" in preset:
            cls.ENABLE_CONV = preset["enable_conv"]
        if "target_module" in preset:
            cls.TARGET_MODULE = list(preset["target_module"])
        if "target_name" in preset:
            cls.TARGET_NAME = list(preset["target_name"])
        if "module_algo_map" in preset:
            cls.MODULE_ALGO_MAP = dict(preset["module_algo_map"])

    def __init__(self, module: nn.Module, multiplier: float = 1.0, linear_dim: int = 4, linear_alpha: float = 1.0,
                 algo: str = "lora", train_norm: bool = False, **kwargs):
        super().__init__()
        self.multiplier = multiplier
        self.loras: List[_Adapter] = []
        self._orgs: List[nn.Module] = []
        names = set()
        if self.ENABLE_CONV:
            raise NotImplementedError("uwudiff_b200.lycoris: conv adapters (enable_conv = true) are not built; the "
                                      "shipped preset configs/lycoris/sdxl-diffusers.toml sets enable_conv = false")

        def single(name, mod, algo, cfg):
            if isinstance(mod, nn.Linear) and linear_dim > 0:
                if algo == "lokr":
                    if not cfg.get("full_matrix", False):
                        raise NotImplementedError("uwudiff_b200.lycoris: LoKr without full_matrix is not built")
                    return LokrLinear(name, mod, multiplier, linear_dim, linear_alpha, int(cfg.get("factor", -1)))
                if algo == "lora":
                    return LoraLinear(name, mod, multiplier, linear_dim, linear_alpha)
                raise NotImplementedError(f"uwudiff_b200.lycoris: algo '{algo}' is not built (lora, lokr)")
            if isinstance(mod, (nn.GroupNorm, nn.LayerNorm)) and train_norm:
                return NormDelta(name, mod, multiplier)
            return None

        def walk(prefix, root, algo, cfg):
            for name, mod in root.named_modules():
                cls_name = mod.__class__.__name__
                if cls_name in self.MODULE_ALGO_MAP and mod is not root:
                    nxt = self.MODULE_ALGO_MAP[cls_name]
                    walk(f"{prefix}.{name}" if name else prefix, mod, nxt.get("algo", algo), nxt)
                lname = (f"{prefix}.{name}" if name else prefix).replace(".", "_")
                if lname in names:
                    continue
                lora = single(lname, mod, algo, cfg)
                if lora is not None:
                    names.add(lname)
                    self.loras.append(lora)
                    self._orgs.append(mod)

        for name, mod in module.named_modules():
            if mod.__class__.__name__ in self.TARGET_MODULE:
                walk(f"{self.LORA_PREFIX}.{name}", mod, algo, kwargs)
        for lora in self.loras:
            self.add_module(lora.lora_name, lora)
        dev = next(module.parameters()).device
        self._flatten(dev)

    # all adapter parameters (and gradients) are views into two flat fp32 buffers
    def _flatten(self, device):
        params = list(self.parameters())
        total = sum(_pad4(p.numel()) for p in params)
        flat = torch.zeros((total,), device=device, dtype=torch.float32)
        gflat = torch.zeros((total,), device=device, dtype=torch.float32)
        off = 0
        for p in params:
            n = p.numel()
            v = flat[off:off + n].view(p.shape)
            v.copy_(p.data)
            p.data = v
            p.grad = gflat[off:off + n].view(p.shape)
            off += _pad4(n)
        object.__setattr__(self, "flat_params", flat)
        object.__setattr__(self, "flat_grads", gflat)

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        ps = list(self.parameters())
        if ps and ps[0].device != self.flat_params.device:
            self._flatten(ps[0].device)
        return r

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grads.zero_()

    def apply_to(self):
        for lora, org in zip(self.loras, self._orgs):
            object.__setattr__(org, "_uwu_adapter", lora)
            if hasattr(org, "drop_cache"):
                org.drop_cache()

    def restore(self):
        for org in self._orgs:
            object.__setattr__(org, "_uwu_adapter", None)
            if hasattr(org, "drop_cache"):
                org.drop_cache()

    @torch.no_grad()
    def merge_to(self, weight: float = 1.0):
        for lora, org in zip(self.loras, self._orgs):
            if isinstance(lora, NormDelta):
                org.weight += lora.w_norm * weight
                org.bias += lora.b_norm * weight
            else:
                org.weight += lora.delta().to(org.weight.dtype) * weight
            if hasattr(org, "drop_cache"):
                org.drop_cache()


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def create_lycoris(module: nn.Module, multiplier: float = 1.0, linear_dim: int = 4, linear_alpha: float = 1.0,
                   algo: str = "lora", train_norm: bool = False, **kwargs) -> LycorisNetwork:
    for k in ("conv_dim", "conv_alpha", "use_tucker"):
        kwargs.pop(k, None)  # conv settings are moot with enable_conv = false
    return LycorisNetwork(module, multiplier=multiplier, linear_dim=linear_dim, linear_alpha=linear_alpha, algo=algo,
                          train_norm=train_norm, **kwargs)
