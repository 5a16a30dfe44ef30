"""CPU restatement of the LyCORIS wrapper as the duwu trainer uses it — TEST INFRASTRUCTURE ONLY.

Reference call sites: src/duwu/trainer/trainer.py:13,148-169 (`LycorisNetwork.apply_preset(preset)`,
`create_lycoris(unet, **config)`, `.apply_to()`, `.parameters()`), :184-187 (`restore`, `merge_to`), :211-215 (state dict
dump); preset file configs/lycoris/sdxl-diffusers.toml:1-27.  The arithmetic lives in the un-vendored dependency
`lycoris-lora>=3.0.1.dev10` (pyproject.toml:33), absent from /root/reference and from this image, so its published
algorithm is restated here (PARITY UNPINNED): `factorization`, LoRA (`lora_down`/`lora_up`, alpha/r), LoKr with
`full_matrix` (`lokr_w1` kaiming, `lokr_w2` zeros, scale 1, dW = kron(w1, w2)), norm deltas (`w_norm`, `b_norm`), the
preset walk (target_module classes -> children; module_algo_map by ancestor class) and `lycoris_<path>` naming.
Forward is the non-bypass form: F.linear(x, W + dW * multiplier, b).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F


def factorization(dimension: int, factor: int = -1):
    """lycoris.functional.general.factorization (restated): (m, n) with m <= n, m * n == dimension."""
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


class LoraLinear(nn.Module):
    def __init__(self, name: str, org: nn.Linear, multiplier: float, dim: int, alpha: float):
        super().__init__()
        self.lora_name, self.multiplier, self.dim = name, multiplier, dim
        self.lora_down = nn.Linear(org.in_features, dim, bias=False)
        self.lora_up = nn.Linear(dim, org.out_features, bias=False)
        self.register_buffer("alpha", torch.tensor(float(alpha)))
        self.scale = alpha / dim
        nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
        nn.init.constant_(self.lora_up.weight, 0)
        self.org = [org]

    def delta(self):
        return (self.lora_up.weight @ self.lora_down.weight) * self.scale

    def forward(self, x):
        org = self.org[0]
        return F.linear(x, org.weight + self.delta().to(org.weight.dtype) * self.multiplier, org.bias)


class LohaLinear(nn.Module):
    """LyCORIS `loha` (Hadamard product of two low-rank products) **[restated]**: dW = (w1a @ w1b) * (w2a @ w2b) * alpha/r;
    init: w1_b ~ N(0, 1), w1_a ~ N(0, 0.1), w2_b ~ N(0, 1), w2_a = 0 (so dW starts at 0)."""

    def __init__(self, name: str, org: nn.Linear, multiplier: float, dim: int, alpha: float):
        super().__init__()
        self.lora_name, self.multiplier, self.dim = name, multiplier, dim
        self.hada_w1_a = nn.Parameter(torch.empty(org.out_features, dim))
        self.hada_w1_b = nn.Parameter(torch.empty(dim, org.in_features))
        self.hada_w2_a = nn.Parameter(torch.empty(org.out_features, dim))
        self.hada_w2_b = nn.Parameter(torch.empty(dim, org.in_features))
        self.register_buffer("alpha", torch.tensor(float(alpha)))
        self.scale = alpha / dim
        nn.init.normal_(self.hada_w1_b, std=1)
        nn.init.normal_(self.hada_w1_a, std=0.1)
        nn.init.normal_(self.hada_w2_b, std=1)
        nn.init.constant_(self.hada_w2_a, 0)
        self.org = [org]

    def delta(self):
        return (self.hada_w1_a @ self.hada_w1_b) * (self.hada_w2_a @ self.hada_w2_b) * self.scale

    def forward(self, x):
        org = self.org[0]
        return F.linear(x, org.weight + self.delta().to(org.weight.dtype) * self.multiplier, org.bias)


class LokrLinear(nn.Module):
    """full_matrix LoKr: both Kronecker factors dense (no low-rank split), scale forced to 1."""

    def __init__(self, name: str, org: nn.Linear, multiplier: float, dim: int, alpha: float, factor: int):
        super().__init__()
        self.lora_name, self.multiplier = name, multiplier
        in_m, in_n = factorization(org.in_features, factor)
        out_l, out_k = factorization(org.out_features, factor)
        self.shape = ((out_l, out_k), (in_m, in_n))
        self.lokr_w1 = nn.Parameter(torch.empty(out_l, in_m))
        self.lokr_w2 = nn.Parameter(torch.empty(out_k, in_n))
        self.register_buffer("alpha", torch.tensor(float(dim)))  # use_w1 and use_w2 -> alpha = lora_dim -> scale 1
        self.scale = 1.0
        nn.init.constant_(self.lokr_w2, 0)
        nn.init.kaiming_uniform_(self.lokr_w1, a=math.sqrt(5))
        self.org = [org]

    def delta(self):
        return torch.kron(self.lokr_w1, self.lokr_w2.contiguous()) * self.scale

    def forward(self, x):
        org = self.org[0]
        return F.linear(x, org.weight + self.delta().to(org.weight.dtype) * self.multiplier, org.bias)


class NormDelta(nn.Module):
    def __init__(self, name: str, org: nn.Module, multiplier: float):
        super().__init__()
        self.lora_name, self.multiplier = name, multiplier
        dim = org.weight.shape[0]
        self.w_norm = nn.Parameter(torch.zeros(dim))
        self.b_norm = nn.Parameter(torch.zeros(dim))
        self.org = [org]

    def forward(self, x):
        org = self.org[0]
        w = org.weight + self.w_norm * self.multiplier
        b = org.bias + self.b_norm * self.multiplier
        if isinstance(org, nn.GroupNorm):
            return F.group_norm(x, org.num_groups, w, b, org.eps)
        return F.layer_norm(x, org.normalized_shape, w, b, org.eps)


class LycorisNetwork(nn.Module):
    """Restated `lycoris.LycorisNetwork` (preset class attributes + wrapper)."""

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
        self.loras: List[nn.Module] = []
        names = set()

        def single(name, mod, algo, cfg):
            if isinstance(mod, nn.Linear) and linear_dim > 0:
                if algo == "lokr":
                    return LokrLinear(name, mod, multiplier, linear_dim, linear_alpha, int(cfg.get("factor", -1)))
                if algo == "lora":
                    return LoraLinear(name, mod, multiplier, linear_dim, linear_alpha)
                if algo == "loha":
                    return LohaLinear(name, mod, multiplier, linear_dim, linear_alpha)
                raise NotImplementedError(algo)
            if isinstance(mod, (nn.GroupNorm, nn.LayerNorm)) and train_norm:
                return NormDelta(name, mod, multiplier)
            return None  # 3x3 convs: enable_conv = false in the shipped preset

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

        for name, mod in module.named_modules():
            if mod.__class__.__name__ in self.TARGET_MODULE:
                walk(f"{self.LORA_PREFIX}.{name}", mod, algo, kwargs)
        for lora in self.loras:
            self.add_module(lora.lora_name, lora)

    def apply_to(self):
        for lora in self.loras:
            org = lora.org[0]
            lora._org_forward = org.forward
            org.forward = lora.forward

    def restore(self):
        for lora in self.loras:
            lora.org[0].forward = lora._org_forward

    @torch.no_grad()
    def merge_to(self, weight: float = 1.0):
        for lora in self.loras:
            org = lora.org[0]
            if isinstance(lora, NormDelta):
                org.weight += lora.w_norm * weight
                org.bias += lora.b_norm * weight
            else:
                org.weight += lora.delta().to(org.weight.dtype) * weight


def create_lycoris(module: nn.Module, multiplier: float = 1.0, linear_dim: int = 4, linear_alpha: float = 1.0,
                   algo: str = "lora", train_norm: bool = False, **kwargs) -> LycorisNetwork:
    kwargs.pop("conv_dim", None)
    kwargs.pop("conv_alpha", None)
    kwargs.pop("use_tucker", None)
    return LycorisNetwork(module, multiplier=multiplier, linear_dim=linear_dim, linear_alpha=linear_alpha, algo=algo,
                          train_norm=train_norm, **kwargs)
