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
import os
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


_LOKR_FUSED = os.environ.get("UWU_LOKR_FUSED", "1") != "0"
# Mirrored factored route for the FeedForward down projections (UWU_LOKR_MIRROR=0 falls back to the G = dY^T X route): same-box
# A/B -3 ms of a 347 ms step (profiles/r02_bench_g_*).
_LOKR_MIRROR = os.environ.get("UWU_LOKR_MIRROR", "1") != "0"


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
        if dim > 16:
            raise NotImplementedError("uwudiff_b200.lycoris: LoRA rank > 16 is not built (the gradient kernels keep the rank in registers)")
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


class LohaLinear(_Adapter):
    """lycoris `loha`: dW = (hada_w1_a @ hada_w1_b) o (hada_w2_a @ hada_w2_b) * alpha / r, folded into the bf16 GEMM operand;
    factor gradients from G = dY^T X (uwu_loha_grad)."""

    def __init__(self, name: str, org: nn.Linear, multiplier: float, dim: int, alpha: float):
        super().__init__()
        if dim > 16:
            raise NotImplementedError("uwudiff_b200.lycoris: LoHa rank > 16 is not built")
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

    def fold_into(self, W, dst):
        ops.fold_loha(W, self.hada_w1_a, self.hada_w1_b, self.hada_w2_a, self.hada_w2_b, self.scale * self.multiplier, dst)

    def grads_from(self, G):
        from .unet import _grad_of

        ops.loha_grad(G, self.hada_w1_a, self.hada_w1_b, self.hada_w2_a, self.hada_w2_b, self.scale * self.multiplier,
                      _grad_of(self.hada_w1_a), _grad_of(self.hada_w1_b), _grad_of(self.hada_w2_a), _grad_of(self.hada_w2_b))

    def delta(self):
        return (self.hada_w1_a @ self.hada_w1_b) * (self.hada_w2_a @ self.hada_w2_b) * self.scale


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
        self._w2_bf16 = None  # view into LycorisNetwork.flat_params_bf16 (refreshed once per backward)
        nn.init.constant_(self.lokr_w2, 0)
        nn.init.kaiming_uniform_(self.lokr_w1, a=math.sqrt(5))

    def fold_into(self, W, dst):
        ops.fold_lokr(W, self.lokr_w1, self.lokr_w2, dst, self.scale * self.multiplier)

    def g_buffer(self, N: int, K: int):
        """Persistent fp32 [N, K] buffer for this adapter's G = dY^T X: the contraction into (dw1, dw2) is deferred and runs
        batched with the other adapters of the same top-level block (LycorisNetwork.flush_grads).  None = not deferring."""
        net = getattr(self, "_net", None)
        if net is None or not net.defer_grads or not self.lokr_w1.is_cuda:
            return None
        g = getattr(self, "_g", None)
        if g is None or g.shape != (N, K) or g.device != self.lokr_w1.device:
            g = torch.empty((N, K), device=self.lokr_w1.device, dtype=torch.float32)
            object.__setattr__(self, "_g", g)
        return g

    def grads_from(self, G, persistent: bool = False):
        from .unet import _grad_of

        net = getattr(self, "_net", None)
        if persistent and net is not None and net.defer_grads:
            _grad_of(self.lokr_w1), _grad_of(self.lokr_w2)
            net._pending.append((self, G))
            return
        ops.lokr_grad(G, self.lokr_w1, self.lokr_w2, _grad_of(self.lokr_w1), _grad_of(self.lokr_w2),
                      self.scale * self.multiplier)

    def delta(self):
        return torch.kron(self.lokr_w1, self.lokr_w2.contiguous()) * self.scale

    def factored_ok(self, M: int) -> bool:
        """Use the factored gradient (no G = dY^T X) when the factor shapes fit its kernels and it is cheaper than the
        full token-reduction GEMM: 2/in_m of the FLOPs, but four passes over [M, *]-sized data."""
        (ol, ok), (im, inn) = self.shape
        if self.fused_ok(M) or self.mirror_ok(M):
            return True
        # measured on B200 (profiles/r01_lokr_factored_stages.log): the factored route wins for the FeedForward adapters
        # (w2 2048x256: 263 vs 454 us, w2 1024x128: 404 vs 673 us) and loses for the 64x64 attention factors (135 vs 82 us),
        # whose four stages are latency / issue bound rather than FLOP bound
        return (self._w2_bf16 is not None and M >= 4096 and ol <= 32 and im <= 32 and im >= 4 and inn % 16 == 0 and ok % 8 == 0
                and inn <= 256 and ok * inn >= 32768)

    def fused_ok(self, M: int) -> bool:
        """64x64 w2 (the attention adapters): ONE pass over x and dY on tcgen05 (uwu_lokr_fused_grad) instead of the
        token-reduction GEMM G = dY^T X + contraction."""
        (ol, ok), (im, inn) = self.shape
        # (the ~12 us of per-launch set-up — operand build, pipeline fill, final reduction — only pay off from a few thousand
        # tokens on: the cross-attention K / V adapters see 1232 text tokens and stay on the G route, 35 us for the pair)
        return _LOKR_FUSED and M >= 4096 and self.lokr_w1.is_cuda and ops.lokr_fused_supported(ol, ok, im, inn)

    def mirror_ok(self, M: int) -> bool:
        """out_k < in_n (the FeedForward down projection: w2 256 x 1024): the factored route with the w1-mixing applied to dY
        instead of X, so that every intermediate is [M, im * ok] wide instead of [M, ol * in_n] (lokr_factored_grads_mirror)."""
        (ol, ok), (im, inn) = self.shape
        # measured (profiles/r02_lokr_mirror.log): w2 256 x 1024 at 16384 tokens 170 vs 184 us on the G route; w2 128 x 512 at
        # 65536 tokens LOSES (299 vs 207 us: one row tile, no CTA pairs, five 128-wide GEMMs), hence ok == 256
        return (_LOKR_MIRROR and self._w2_bf16 is not None and M >= 4096 and ol <= 32 and im <= 32 and ol >= 4 and ok == 256
                and inn % 64 == 0 and ok < inn)

    def grads_factored(self, dy, x, M):
        """dy: bf16 [M, ol*ok] (row stride dy.stride(0)), x: bf16 [M, im*inn]; accumulates into lokr_w1.grad / lokr_w2.grad."""
        from .unet import _grad_of

        if not self.fused_ok(M) and self.mirror_ok(M):
            if not dy.is_contiguous():  # (a column slice of a fused projection: the dY-side kernels want whole rows)
                dy = ops.copy2d(dy, torch.empty((M, dy.shape[1]), device=dy.device, dtype=dy.dtype))
            lokr_factored_grads_mirror(dy, x, M, self.lokr_w1.detach(), self._w2_bf16, _grad_of(self.lokr_w1), _grad_of(self.lokr_w2),
                                       self.scale * self.multiplier)
            return

        if self.fused_ok(M):
            ops.lokr_fused_grad(dy, x, M, self.lokr_w1.detach(), self.lokr_w2.detach(), _grad_of(self.lokr_w1), _grad_of(self.lokr_w2),
                                self.scale * self.multiplier)
            return
        lokr_factored_grads(dy, x, M, self.lokr_w1, self._w2_bf16, _grad_of(self.lokr_w1), _grad_of(self.lokr_w2),
                            self.scale * self.multiplier)


def lokr_factored_grads(dy, x, M, w1, w2_bf16, dw1, dw2, mult: float = 1.0):
    """Adapter gradients of y = x (W + kron(w1, w2))^T without G = dY^T X:
         Z[m,l,n] = sum_i w1[l,i] X[m,i,n];  dw2 += sum_{m,l} dY[m,l,:]^T Z[m,l,:];
         V[m,l,n] = sum_k dY[m,l,k] w2[k,n]; dw1[l,i] += sum_{m,n} V[m,l,n] X[m,i,n]."""
    from ._lib import A_COL, A_ROW, B_KN

    (ol, im), (ok, inn) = w1.shape, w2_bf16.shape
    zw = ops._workspace((M * ol * inn + 1) // 2, dy.device, "lokr_z").view(torch.bfloat16)[: M * ol * inn].view(M, ol * inn)
    ops.lokr_z(x, w1, M, inn, zw)
    # token-reduction GEMM whose reduction runs over (l, m): segment l reads dY[:, l*ok:(l+1)*ok] and Z[:, l*inn:(l+1)*inn]
    ops.gemm(dy, zw, ok, inn, M, a_layout=A_COL, lda=dy.stride(0), b_layout=B_KN, ldb=ol * inn, out=dw2, accumulate=True,
             alpha=mult, stream_k=1, k_segs=ol, a_seg_off=ok, b_seg_off=inn)
    # V_l = dY_l w2: one shared right operand for all l (grouped N)
    vw = ops._workspace((M * ol * inn + 1) // 2, dy.device, "lokr_v").view(torch.bfloat16)[: M * ol * inn].view(M, ol * inn)
    bn = next(b for b in (256, 128, 64, 32, 16) if inn % b == 0)
    ops.gemm(dy, w2_bf16, M, ol * inn, ok, a_layout=A_ROW, lda=dy.stride(0), b_layout=B_KN, ldb=inn, out=vw, grp_n=inn,
             a_grp_koff=ok, block_n=bn)
    ops.lokr_dw1(vw, x, M, ol, im, inn, dw1, mult)


def lokr_factored_grads_mirror(dy, x, M, w1, w2_bf16, dw1, dw2, mult: float = 1.0):
    """The same gradients with the w1-mixing applied on the OUTPUT side (cheaper when out_k < in_n, e.g. w2 256 x 1024):
         U[m,j,:] = sum_i w1[i,j] dY[m,i,:];          dw2 += sum_{m,j} U[m,j,:]^T X[m,j,:];
         T[m,j,:] = X[m,j,:] w2^T  ([M, im*ok]);      dw1[i,j] += sum_m <dY[m,i,:], T[m,j,:]>.
    Every intermediate is [M, im * ok] (42 MB at 16384 x 1280) where the X-side route writes [M, ol * in_n] (168 MB) twice."""
    from ._lib import A_COL, B_KN

    (ol, im), (ok, inn) = w1.shape, w2_bf16.shape
    uw = ops._workspace((M * im * ok + 1) // 2, dy.device, "lokr_z").view(torch.bfloat16)[: M * im * ok].view(M, im * ok)
    ops.lokr_z(dy, w1, M, ok, uw, transposed=True)
    # token-reduction GEMM whose reduction runs over (j, m): segment j reads U[:, j*ok:(j+1)*ok] and X[:, j*inn:(j+1)*inn]
    ops.gemm(uw, x, ok, inn, M, a_layout=A_COL, lda=im * ok, b_layout=B_KN, ldb=x.stride(0), out=dw2, accumulate=True,
             alpha=mult, stream_k=1, k_segs=im, a_seg_off=ok, b_seg_off=inn)
    tw = ops._workspace((M * im * ok + 1) // 2, dy.device, "lokr_v").view(torch.bfloat16)[: M * im * ok].view(M, im * ok)
    # T_j = X_j w2^T for all j in one grouped launch: group j reads A columns [j*inn, +inn) and the SAME right operand, w2
    # [ok, inn] K-major as stored
    bn = next(b for b in (256, 128, 64, 32, 16) if ok % b == 0)
    ops.gemm(x, w2_bf16, M, im * ok, inn, lda=x.stride(0), out=tw, grp_n=ok, a_grp_koff=inn, block_n=bn)
    ops.lokr_dw1(dy, tw, M, ol, im, ok, dw1, mult)


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
        object.__setattr__(self, "_root", module)
        self.loras: List[_Adapter] = []
        self._orgs: List[nn.Module] = []
        # deferred LoKr contractions: (adapter, G) pairs waiting for flush_grads(); UWU_LOKR_DEFER=0 restores per-layer launches
        self.defer_grads = os.environ.get("UWU_LOKR_DEFER", "1") != "0"
        object.__setattr__(self, "_pending", [])
        object.__setattr__(self, "_grad_tables", {})
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
                if algo == "loha":
                    return LohaLinear(name, mod, multiplier, linear_dim, linear_alpha)
                raise NotImplementedError(f"uwudiff_b200.lycoris: algo '{algo}' is not built (lora, lokr, loha)")
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
        total = sum(_pad8(p.numel()) for p in params)
        total = (total + 1023) // 1024 * 1024
        flat = torch.zeros((total,), device=device, dtype=torch.float32)
        gflat = torch.zeros((total,), device=device, dtype=torch.float32)
        flat16 = torch.zeros((total,), device=device, dtype=torch.bfloat16) if flat.is_cuda else None
        off = 0
        offs = {}
        for p in params:
            n = p.numel()
            v = flat[off:off + n].view(p.shape)
            v.copy_(p.data)
            p.data = v
            p.grad = gflat[off:off + n].view(p.shape)
            offs[id(p)] = off
            off += _pad8(n)
        object.__setattr__(self, "flat_params", flat)
        object.__setattr__(self, "flat_grads", gflat)
        object.__setattr__(self, "flat_params_bf16", flat16)
        object.__setattr__(self, "_fold", None)
        for lora in self.loras:
            if isinstance(lora, LokrLinear):
                o2 = offs[id(lora.lokr_w2)]
                lora._w2_bf16 = None if flat16 is None else flat16[o2:o2 + lora.lokr_w2.numel()].view(lora.lokr_w2.shape)

    def fold_all(self):
        """One launch folding every adapter into the operand its module's GEMM / norm kernel reads (uwu_fold_batch)."""
        import ctypes as C

        from . import _lib
        from .unet import FOLD, Attention

        if not self.flat_params.is_cuda:
            return
        t = getattr(self, "_fold", None)
        if t is None or t["gen"] != FOLD.gen:
            for m in self._root.modules():
                if isinstance(m, Attention):
                    m.ensure_fused()
            ents, keep = [], []
            chunk = 1 << 14

            def add(kind, W, a, b, dst, N, K, p0=0, p1=0, p2=0, scale=1.0):
                assert W.is_contiguous() and dst.is_contiguous() and (kind == 3 or K % 4 == 0) and N * K < 2 ** 31
                e = _lib.FoldEntry()
                e.W, e.a, e.b, e.dst = W.data_ptr(), a.data_ptr(), (b.data_ptr() if b is not None else None), dst.data_ptr()
                e.kind, e.N, e.K, e.p0, e.p1, e.p2, e.scale = kind, N, K, p0, p1, p2, scale
                ents.append(e)
                keep.extend([W, a, b, dst])

            for lora, org in zip(self.loras, self._orgs):
                if isinstance(lora, NormDelta):
                    g, b = org.fold_dst()
                    Cn = org.weight.numel()
                    add(3, org.weight, lora.w_norm, None, g, 1, Cn, scale=lora.multiplier)
                    add(3, org.bias, lora.b_norm, None, b, 1, Cn, scale=lora.multiplier)
                elif isinstance(lora, LokrLinear):
                    (ol, ok), (im, inn) = lora.shape
                    add(1, org.weight, lora.lokr_w1, lora.lokr_w2, org.fold_dst(), org.out_features, org.in_features, ok, inn, im,
                        lora.scale * lora.multiplier)
                elif isinstance(lora, LohaLinear):
                    base = lora.hada_w1_a.data_ptr()
                    off_a, off_b = lora.hada_w2_a.data_ptr() - base, lora.hada_w2_b.data_ptr() - base
                    assert off_a % 4 == 0 and off_b % 4 == 0 and abs(off_a) < 2 ** 33 and abs(off_b) < 2 ** 33
                    add(4, org.weight, lora.hada_w1_a, lora.hada_w1_b, org.fold_dst(), org.out_features, org.in_features,
                        lora.dim, off_a // 4, off_b // 4, lora.scale * lora.multiplier)
                    keep.extend([lora.hada_w2_a, lora.hada_w2_b])
                else:
                    add(2, org.weight, lora.lora_up.weight, lora.lora_down.weight, org.fold_dst(), org.out_features,
                        org.in_features, lora.dim, scale=lora.scale * lora.multiplier)
            chunk_entry = []
            for i, e in enumerate(ents):
                e.chunk0 = len(chunk_entry)
                chunk_entry += [i] * ((e.N * e.K + chunk - 1) // chunk)
            arr = (_lib.FoldEntry * len(ents))(*ents)
            dev = self.flat_params.device
            table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
            ce = torch.tensor(chunk_entry, dtype=torch.int32, device=dev)
            t = dict(gen=FOLD.gen, table=table, ce=ce, n=len(chunk_entry), chunk=chunk, keep=keep)
            object.__setattr__(self, "_fold", t)
        _lib.check(_lib.lib().uwu_fold_batch(t["table"].data_ptr(), t["ce"].data_ptr(), t["n"], t["chunk"],
                                             torch.cuda.current_stream().cuda_stream), "uwu_fold_batch")
        for org in self._orgs:
            org._fold_epoch = FOLD.epoch

    def mark_folded(self) -> bool:
        """The operands still hold the fold of the previous forward (the adapters have not changed since: later micro-batches
        of a gradient-accumulation window).  False if there is no previous fold to rely on."""
        from .unet import FOLD

        t = getattr(self, "_fold", None)
        if t is None or t["gen"] != FOLD.gen:
            return False
        for org in self._orgs:
            org._fold_epoch = FOLD.epoch
        return True

    def flush_grads(self):
        """Contract every pending G of the LoKr adapters into (dw1, dw2) in ONE launch (uwu_lokr_grad_batch).  Called by the
        denoiser's backward whenever a top-level block is finished (before its gradients are handed to the DDP buckets)."""
        from . import _lib

        pend = self._pending
        if not pend:
            return
        key = tuple((id(ad), G.data_ptr(), ad.lokr_w1.grad.data_ptr(), ad.lokr_w2.grad.data_ptr(), ad.lokr_w1.data_ptr())
                    for ad, G in pend)
        t = self._grad_tables.get(key)
        if t is None:
            n = len(pend)
            sms = torch.cuda.get_device_properties(pend[0][1].device).multi_processor_count
            target = max(8, (8 * sms) // n)
            ents, block0 = [], 0
            for ad, G in pend:
                (ol, ok), (im, inn) = ad.shape
                e = _lib.LokrGradEntry()
                e.G, e.w1, e.w2 = G.data_ptr(), ad.lokr_w1.data_ptr(), ad.lokr_w2.data_ptr()
                e.dw1, e.dw2 = ad.lokr_w1.grad.data_ptr(), ad.lokr_w2.grad.data_ptr()
                e.ldg = G.stride(0)
                e.out_l, e.out_k, e.in_m, e.in_n = ol, ok, im, inn
                e.multiplier = ad.scale * ad.multiplier
                e.vec = int(inn % 4 == 0 and G.stride(0) % 4 == 0 and G.data_ptr() % 16 == 0 and ad.lokr_w2.data_ptr() % 16 == 0)
                e.target, e.block0 = target, block0
                nb = int(_lib.lib().uwu_lokr_grad_plan_blocks(ol, ok, im, inn, e.vec, target))
                assert nb > 0
                block0 += nb
                ents.append(e)
            arr = (_lib.LokrGradEntry * n)(*ents)
            # staged through PINNED memory that lives as long as the table: under CUDA-graph capture the upload becomes a
            # memcpy node that re-reads the host buffer on every replay (a pageable temporary would be freed by then)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory()
            table = torch.empty_like(host, device=pend[0][1].device)
            table.copy_(host, non_blocking=True)
            t = (table, n, block0, host)
            if len(self._grad_tables) > 64 and not torch.cuda.is_current_stream_capturing():
                self._grad_tables.clear()
            self._grad_tables[key] = t
        _lib.check(_lib.lib().uwu_lokr_grad_batch(t[0].data_ptr(), t[1], t[2], torch.cuda.current_stream().cuda_stream),
                   "uwu_lokr_grad_batch")
        pend.clear()

    def refresh_bf16(self):
        """bf16 copy of all adapter parameters (operands of the factored-gradient GEMMs): one kernel per backward."""
        if self.flat_params_bf16 is not None:
            ops.copy2d(self.flat_params.view(-1, 1024), self.flat_params_bf16.view(-1, 1024))

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        ps = list(self.parameters())
        if ps and ps[0].device != self.flat_params.device:
            self._flatten(ps[0].device)
        return r

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grads.zero_()

    def apply_to(self):
        object.__setattr__(self._root, "_uwu_lycoris", self)
        for lora, org in zip(self.loras, self._orgs):
            object.__setattr__(org, "_uwu_adapter", lora)
            object.__setattr__(lora, "_net", self)
            if hasattr(org, "drop_cache"):
                org.drop_cache()

    def restore(self):
        object.__setattr__(self._root, "_uwu_lycoris", None)
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


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def create_lycoris(module: nn.Module, multiplier: float = 1.0, linear_dim: int = 4, linear_alpha: float = 1.0,
                   algo: str = "lora", train_norm: bool = False, **kwargs) -> LycorisNetwork:
    for k in ("conv_dim", "conv_alpha", "use_tucker"):
        kwargs.pop(k, None)  # conv settings are moot with enable_conv = false
    return LycorisNetwork(module, multiplier=multiplier, linear_dim=linear_dim, linear_alpha=linear_alpha, algo=algo,
                          train_norm=train_norm, **kwargs)
