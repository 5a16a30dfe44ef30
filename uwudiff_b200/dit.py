"""DiT (class-conditional diffusion transformer with adaLN-Zero blocks) on the sm_100a kernels — BASELINE.json configs[3].

The reference tree ships no DiT model; its only in-tree trace of the architecture is the `ada_norm_zero`
branch of the patched transformer block (/root/reference/src/duwu/modules/rope_unet.py:306-309, :344-345, :395-398,
:406-407).  This module therefore mirrors the public DiT definition (parameter names of the original implementation:
`x_embedder.proj`, `t_embedder.mlp.{0,2}`, `y_embedder.embedding_table`, `blocks.{i}.{attn.qkv, attn.proj, mlp.fc1, mlp.fc2,
adaLN_modulation.1}`, `final_layer.{linear, adaLN_modulation.1}`) behind the same denoiser protocol the reference loss uses:
`model(sample, timestep, added_cond_kwargs={"class_labels": y})[0]` (src/duwu/loss/diffusion.py:172-176).

Like the UNet, the whole network is one `torch.autograd.Function`: forward and backward are hand-scheduled kernel
sequences over token-major bf16 activations [B*T, D] with fp32 master weights.

  per block (forward):  mod = Linear(SiLU(c)) fp32 [B, 6D]      tcgen05 GEMM, fp32 epilogue
                        n1  = LN(x) * (1 + scale) + shift        adaln_fwd
                        qkv = n1 Wqkv^T + b ;  o = flash(q,k,v)  GEMM ; attention (d = 72 -> mma.sync path, d <= 64 tcgen05)
                        x1  = x + gate * (o Wp^T + b)            GEMM ; gate_residual_fwd
                        n2  = LN(x1) * (1 + scale) + shift       adaln_fwd
                        x2  = x1 + gate * fc2(gelu_tanh(fc1(n2)))  GEMM ; elementwise ; GEMM ; gate_residual_fwd
"""
from __future__ import annotations

import types
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .ops import B_KN
from .unet import Linear, _grad_of

BF16 = torch.bfloat16

DIT_XL_2_CONFIG = dict(input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16, mlp_ratio=4.0,
                       num_classes=1000, learn_sigma=True, frequency_embedding_size=256)
KNOWN_DIT_CONFIGS = {"DiT-XL/2": DIT_XL_2_CONFIG, "facebook/DiT-XL-2-256": DIT_XL_2_CONFIG}


def sincos_pos_embed_2d(dim: int, grid: int) -> torch.Tensor:
    """Fixed 2-D sin-cos table [grid*grid, dim]: first half of the channels from the w coordinate, second from h."""
    def axis(d, pos):
        omega = 1.0 / 10000 ** (torch.arange(d // 2, dtype=torch.float64) / (d / 2.0))
        ang = pos.reshape(-1)[:, None] * omega[None, :]
        return torch.cat([ang.sin(), ang.cos()], dim=1)

    idx = torch.arange(grid * grid)
    w, h = (idx % grid).double(), (idx // grid).double()
    return torch.cat([axis(dim // 2, w), axis(dim // 2, h)], dim=1).float()


class PatchProj(Linear):
    """`x_embedder.proj`: Conv2d(C, D, p, stride p) as a GEMM over patchified rows (K = C*p*p zero-padded to 64)."""

    def __init__(self, in_channels: int, hidden: int, patch: int):
        k_real = in_channels * patch * patch
        super().__init__(max(64, (k_real + 63) // 64 * 64), hidden)
        self.k_real = k_real
        w = torch.empty((hidden, in_channels, patch, patch))
        nn.init.xavier_uniform_(w.view(hidden, -1))
        self.weight = nn.Parameter(w)

    def w16(self, dst=None):
        if not self._needs_refresh():
            return self._cache
        if getattr(self, "_cache", None) is None:
            self._cache = torch.zeros((self.out_features, self.in_features), device=self.weight.device, dtype=BF16)
        ops.copy2d(self.weight.view(self.out_features, self.k_real), self._cache[:, : self.k_real])
        return self._cache

    def param_grads(self, dy, x, M, **_):
        N, K = self.out_features, self.in_features
        if self.weight.requires_grad:
            G = ops._workspace(N * K, dy.device, "wgrad")[: N * K].view(N, K)
            ops.gemm(dy, x, N, K, M, a_layout=ops.A_COL, lda=dy.stride(0), b_layout=B_KN, ldb=x.stride(0), out=G)
            ops.conv_wgrad_unpack(G, N, self.k_real, K, 1, _grad_of(self.weight))
        if self.bias is not None and self.bias.requires_grad:
            ops.colsum(dy, out=_grad_of(self.bias), accumulate=True)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden: int, freq: int):
        super().__init__()
        self.mlp = nn.ModuleList([Linear(freq, hidden), nn.SiLU(), Linear(hidden, hidden)])
        self.freq = freq

    def fwd(self, tf, B, residual=None):
        h = self.mlp[0].fwd(tf, B)
        a = ops.elementwise(h, None, ops.EW_SILU)
        self._sv = (tf, h, a)
        return self.mlp[2].fwd(a, B, residual=residual)

    def bwd(self, dc, B):
        tf, h, a = self._sv
        self._sv = None
        da = self.mlp[2].bwd(dc, a, B)
        dh = ops.elementwise(da, h, ops.EW_SILU_BWD)
        self.mlp[0].bwd(dh, tf, B, need_dx=False)


class LabelEmbedder(nn.Module):
    def __init__(self, num_classes: int, hidden: int):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + 1, hidden)
        self.num_classes = num_classes


class Attention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.heads, self.dim_head = heads, dim // heads
        self.qkv = Linear(dim, 3 * dim)
        self.proj = Linear(dim, dim)


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = Linear(dim, hidden)
        self.fc2 = Linear(hidden, dim)


class DiTBlock(nn.Module):
    def __init__(self, hidden: int, heads: int, mlp_ratio: float):
        super().__init__()
        self.attn = Attention(hidden, heads)
        self.mlp = Mlp(hidden, int(hidden * mlp_ratio))
        self.adaLN_modulation = nn.ModuleList([nn.SiLU(), Linear(hidden, 6 * hidden)])
        self.D = hidden

    def fwd(self, x, sc, st):
        D, T, B, M = self.D, st.T, st.B, st.M
        mod = self.adaLN_modulation[1].fwd(sc, B, out_dtype=torch.float32)            # [B, 6D] fp32
        n1, s1 = ops.adaln_fwd(x, mod, 0, D, T)
        qkv = self.attn.qkv.fwd(n1, M)
        o, lse = ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, self.attn.heads, T, T, head_dim=self.attn.dim_head)
        a = self.attn.proj.fwd(o, M)
        x1 = ops.gate_residual_fwd(x, a, mod, 2 * D, T)
        n2, s2 = ops.adaln_fwd(x1, mod, 3 * D, 4 * D, T)
        h = self.mlp.fc1.fwd(n2, M)
        g = ops.elementwise(h, None, ops.EW_GELU_TANH)
        m = self.mlp.fc2.fwd(g, M)
        x2 = ops.gate_residual_fwd(x1, m, mod, 5 * D, T)
        self._sv = (x, mod, n1, s1, qkv, o, lse, a, x1, n2, s2, h, g, m)
        return x2

    def bwd(self, dx2, sc, st):
        """dx2: gradient of the block output; returns the gradient of the block input and accumulates d(SiLU(c))."""
        D, T, B, M = self.D, st.T, st.B, st.M
        x, mod, n1, s1, qkv, o, lse, a, x1, n2, s2, h, g, m = self._sv
        self._sv = None
        dmod = torch.empty((B, 6 * D), device=dx2.device, dtype=BF16)
        dm = ops.gate_residual_bwd(dx2, m, mod, 5 * D, T, dmod, 5 * D)
        dg = self.mlp.fc2.bwd(dm, g, M)
        dh = ops.elementwise(dg, h, ops.EW_GELU_TANH_BWD)
        dn2 = self.mlp.fc1.bwd(dh, n2, M)
        dx1 = ops.adaln_bwd(x1, dn2, mod, 4 * D, s2, T, dmod, 3 * D, 4 * D, dres=dx2)
        da = ops.gate_residual_bwd(dx1, a, mod, 2 * D, T, dmod, 2 * D)
        do = self.attn.proj.bwd(da, o, M)
        dqkv = torch.empty((M, 3 * D), device=dx2.device, dtype=BF16)
        ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, B, self.attn.heads, T, T,
                     head_dim=self.attn.dim_head, dq=dqkv[:, :D], dk=dqkv[:, D:2 * D], dv=dqkv[:, 2 * D:])
        dn1 = self.attn.qkv.bwd(dqkv, n1, M)
        dx = ops.adaln_bwd(x, dn1, mod, D, s1, T, dmod, 0, D, dres=dx1)
        st.add_dsc(self.adaLN_modulation[1], dmod, sc)
        return dx


class FinalLayer(nn.Module):
    def __init__(self, hidden: int, patch: int, out_channels: int):
        super().__init__()
        self.linear = Linear(hidden, patch * patch * out_channels)
        self.adaLN_modulation = nn.ModuleList([nn.SiLU(), Linear(hidden, 2 * hidden)])
        self.D = hidden

    def fwd(self, x, sc, st):
        mod = self.adaLN_modulation[1].fwd(sc, st.B, out_dtype=torch.float32)
        n, s = ops.adaln_fwd(x, mod, 0, self.D, st.T)
        self._sv = (x, mod, n, s)
        return self.linear.fwd(n, st.M)

    def bwd(self, dy, sc, st):
        x, mod, n, s = self._sv
        self._sv = None
        dn = self.linear.bwd(dy, n, st.M)
        dmod = torch.empty((st.B, 2 * self.D), device=dy.device, dtype=BF16)
        dx = ops.adaln_bwd(x, dn, mod, self.D, s, st.T, dmod, 0, self.D)
        st.add_dsc(self.adaLN_modulation[1], dmod, sc)
        return dx


class _State:
    """Per-forward bookkeeping shared by the blocks (sizes, the accumulated gradient of SiLU(c))."""

    def __init__(self):
        self.dsc = None

    def add_dsc(self, lin: Linear, dmod, sc):
        """adaLN Linear backward: parameter gradients + d(SiLU(c)) accumulated in fp32 over all blocks."""
        lin.param_grads(dmod, sc, self.B)
        if self.dsc is None:
            self.dsc = torch.zeros((self.B, lin.in_features), device=dmod.device, dtype=torch.float32)
        ops.gemm(dmod, lin._cache, self.B, lin.in_features, lin.out_features, b_layout=B_KN, ldb=lin.in_features, out=self.dsc,
                 accumulate=True, stream_k=0)


class _DiTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hook, model, sample, timestep, labels, fused_temb):
        ctx.model = model
        return model._forward_impl(sample, timestep, labels, fused_temb)

    @staticmethod
    def backward(ctx, gout):
        ctx.model._backward_impl(gout)
        return (torch.zeros((), device=gout.device),) + (None,) * 5


class DiT(nn.Module):
    def __init__(self, **cfg):
        super().__init__()
        c = dict(DIT_XL_2_CONFIG)
        c.update({k: v for k, v in cfg.items() if not k.startswith("_")})
        self.config = types.SimpleNamespace(**c)
        D, p = c["hidden_size"], c["patch_size"]
        if D % 8 != 0 or D % c["num_heads"] != 0 or (D // c["num_heads"]) % 8 != 0 or D // c["num_heads"] > 160:
            raise NotImplementedError(f"uwudiff_b200: DiT hidden {D} / heads {c['num_heads']} unsupported "
                                      "(head dim must be a multiple of 8, <= 160)")
        if c["frequency_embedding_size"] % 8 != 0:
            raise NotImplementedError("uwudiff_b200: frequency_embedding_size must be a multiple of 8")
        self.out_channels = c["in_channels"] * (2 if c["learn_sigma"] else 1)
        self.x_embedder = nn.Module()
        self.x_embedder.proj = PatchProj(c["in_channels"], D, p)
        self.t_embedder = TimestepEmbedder(D, c["frequency_embedding_size"])
        self.y_embedder = LabelEmbedder(c["num_classes"], D)
        self.blocks = nn.ModuleList([DiTBlock(D, c["num_heads"], c["mlp_ratio"]) for _ in range(c["depth"])])
        self.final_layer = FinalLayer(D, p, self.out_channels)
        grid = c["input_size"] // p
        self.register_buffer("pos_embed", sincos_pos_embed_2d(D, grid)[None], persistent=False)
        self._hook = None
        self._pos_cache = None
        self.after_backward = None
        self.gradient_checkpointing = False

    # ---- reference-facing API (same surface as the UNet wrapper, src/duwu/modules/unet_patch.py:13-57) ------------
    @classmethod
    def load_config(cls, config, **_):
        if isinstance(config, str):
            if config not in KNOWN_DIT_CONFIGS:
                raise OSError(f"DiT config '{config}' unknown (known: {sorted(KNOWN_DIT_CONFIGS)}); the HF hub is unreachable")
            return dict(KNOWN_DIT_CONFIGS[config])
        return dict(config)

    @classmethod
    def from_config(cls, config, **kwargs):
        model = cls(**cls.load_config(config, **kwargs))
        model.init_weight()
        return model

    def init_weight(self):
        for m in self.modules():
            if isinstance(m, nn.Linear) and not isinstance(m, PatchProj):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)
        w = self.x_embedder.proj.weight
        nn.init.xavier_uniform_(w.data.view(w.shape[0], -1))
        nn.init.zeros_(self.x_embedder.proj.bias)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for blk in list(self.blocks) + [self.final_layer]:
            nn.init.zeros_(blk.adaLN_modulation[1].weight)
            nn.init.zeros_(blk.adaLN_modulation[1].bias)
        nn.init.zeros_(self.final_layer.linear.weight)
        nn.init.zeros_(self.final_layer.linear.bias)
        self.refresh_weights()

    def ddp_blocks(self):
        """Units reported through `after_backward` (uwudiff_b200/parallel.py), in parameter layout order."""
        return [self.x_embedder, self.t_embedder, self.y_embedder] + list(self.blocks) + [self.final_layer]

    def enable_gradient_checkpointing(self):
        self.gradient_checkpointing = True  # recorded; 180 GB of HBM3e hold the saved activations of the named configs

    def refresh_weights(self):
        for m in self.modules():
            if hasattr(m, "drop_cache"):
                m.drop_cache()
        self._pos_cache = None

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.refresh_weights()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.refresh_weights()
        return r

    def forward(self, sample, timestep, class_labels: Optional[torch.Tensor] = None, encoder_hidden_states=None,
                encoder_attention_mask=None, added_cond_kwargs=None, cross_attention_kwargs=None, _fused_temb=None,
                return_dict: bool = False, **_):
        if not sample.is_cuda:
            raise RuntimeError("uwudiff_b200.DiT runs on the sm_100a kernels only: inputs must be CUDA tensors (no CPU fallback)")
        if class_labels is None:
            class_labels = (added_cond_kwargs or {}).get("class_labels")
        if class_labels is None:
            raise ValueError("DiT needs class_labels (keyword or added_cond_kwargs['class_labels'])")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if self._hook is None:
                self._hook = torch.zeros((), device=sample.device, requires_grad=True)
            out = _DiTFunction.apply(self._hook, self, sample, timestep, class_labels, _fused_temb)
        else:
            out = self._forward_impl(sample, timestep, class_labels, _fused_temb)
            self._drop_saved()
        return (out,)

    def _drop_saved(self):
        for m in self.modules():
            if hasattr(m, "_sv"):
                m._sv = None
        self._fsv = None

    # ---- forward / backward kernel schedules -----------------------------------------------------------------------
    def _pos_rows(self, B, T, dev):
        """pos_embed broadcast to the [B*T, D] token rows (bf16), used as the residual of the patch-embedding GEMM."""
        if self._pos_cache is None or self._pos_cache.shape[0] != B * T or self._pos_cache.device != dev:
            self._pos_cache = self.pos_embed[0].to(dev).to(BF16).repeat(B, 1).contiguous()
        return self._pos_cache

    def _forward_impl(self, sample, timestep, labels, fused_temb):
        c = self.config
        B, Cin, H, W = sample.shape
        p, D = c.patch_size, c.hidden_size
        if H % p or W % p or (H // p) * (W // p) != self.pos_embed.shape[1]:
            raise ValueError(f"DiT input {H}x{W} does not match input_size {c.input_size}")
        st = _State()
        st.B, st.T = B, (H // p) * (W // p)
        st.M = B * st.T
        dev = sample.device
        labels = labels.to(dev).long().contiguous()
        # conditioning c = t_emb + y_emb, shared SiLU(c)
        if fused_temb is not None:
            tf = fused_temb
        else:
            tf = ops.sincos_embed(timestep.to(dev).reshape(-1).expand(B), c.frequency_embedding_size, True)
        yemb = ops.embed_gather(self.y_embedder.embedding_table.weight, labels)
        cvec = self.t_embedder.fwd(tf, B, residual=yemb)
        sc = ops.elementwise(cvec, None, ops.EW_SILU)
        # tokens
        xe = self.x_embedder.proj
        tok = ops.patchify(sample, p, 0, Cin, xe.in_features)
        x = xe.fwd(tok, st.M, residual=self._pos_rows(B, st.T, dev))
        for blk in self.blocks:
            x = blk.fwd(x, sc, st)
        y = self.final_layer.fwd(x, sc, st)
        out = ops.unpatchify(y, B, Cin, H, W, p, 1, self.out_channels)
        self._fsv = (st, tok, cvec, sc, labels, (B, Cin, H, W))
        return out

    def _backward_impl(self, gout):
        st, tok, cvec, sc, labels, (B, Cin, H, W) = self._fsv
        self._fsv = None
        p = self.config.patch_size
        done = self.after_backward or (lambda mods: None)
        fl = self.final_layer.linear
        dy = ops.patchify(gout, p, 1, self.out_channels, fl.out_features)   # sigma channels of learn_sigma get zero gradient
        dx = self.final_layer.bwd(dy, sc, st)
        done([self.final_layer])
        for blk in reversed(self.blocks):
            dx = blk.bwd(dx, sc, st)
            done([blk])
        self.x_embedder.proj.bwd(dx, tok, st.M, need_dx=False)
        if st.dsc is not None:
            dsc = ops.copy2d(st.dsc, torch.empty(st.dsc.shape, device=st.dsc.device, dtype=BF16))
            dc = ops.elementwise(dsc, cvec, ops.EW_SILU_BWD)
            if self.y_embedder.embedding_table.weight.requires_grad:
                ops.embed_scatter_add(dc, labels, _grad_of(self.y_embedder.embedding_table.weight))
            self.t_embedder.bwd(dc, B)
        done([self.x_embedder, self.t_embedder, self.y_embedder])


def dit_forward_flops(cfg: dict, tokens: Optional[int] = None) -> dict:
    """Algorithmic forward FLOPs per sample (GEMM 2MNK, attention 4 L^2 d per head; norms / elementwise excluded)."""
    D, L = cfg["hidden_size"], cfg["depth"]
    T = tokens or (cfg["input_size"] // cfg["patch_size"]) ** 2
    hid = int(D * cfg["mlp_ratio"])
    p2c = cfg["patch_size"] ** 2
    cout = cfg["in_channels"] * (2 if cfg["learn_sigma"] else 1)
    lin = L * 2.0 * T * (3 * D * D + D * D + 2 * D * hid)
    attn = L * 4.0 * T * T * D
    adaln = L * 2.0 * D * 6 * D + 2.0 * D * 2 * D
    emb = 2.0 * T * p2c * cfg["in_channels"] * D + 2.0 * T * D * p2c * cout + 2.0 * (cfg["frequency_embedding_size"] * D + D * D)
    return {"linear": lin, "attn": attn, "adaln": adaln, "embed": emb, "total": lin + attn + adaln + emb}
