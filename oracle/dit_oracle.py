"""TEST INFRASTRUCTURE ONLY — plain-PyTorch fp32 restatement of the DiT denoiser (BASELINE.json configs[3]: "DiT-XL/2
class-conditional, 4x32x32 latents, adaLN-zero").  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this file; the product path (uwudiff_b200/dit.py) never does.

PARITY UNPINNED: the reference tree (/root/reference) contains no DiT model and no golden vectors for one.
What it does contain is the adaLN-Zero block algebra inside its patched diffusers `BasicTransformerBlock.forward`
(/root/reference/src/duwu/modules/rope_unet.py:306-309 `norm1(hidden_states, timestep, class_labels)` -> shift/scale/gate,
:344-345 `gate_msa.unsqueeze(1) * attn_output`, :395-398 `norm2(..) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]`,
:406-407 `gate_mlp.unsqueeze(1) * ff_output`).  Everything else restates the published DiT algorithm (Peebles & Xie, "Scalable
Diffusion Models with Transformers", 2023; constants in SURVEY.md Appendix A.2/A.3):

  * patch embedding: Conv2d(C, D, p, stride p) over the latent, tokens in row-major (h, w) order, + fixed 2-D sin-cos pos-emb;
  * timestep embedding: cat[cos, sin] of t * exp(-ln(1e4) i / 128), i < 128 -> Linear(256, D) -> SiLU -> Linear(D, D);
  * label embedding: table (num_classes + 1) x D (the extra row is the classifier-free "null" class);
  * block: (shift, scale, gate) x 2 = Linear(D, 6D)(SiLU(c)).chunk(6);  x += gate_msa * Attn(LN(x) * (1 + scale_msa) + shift_msa);
           x += gate_mlp * MLP(LN(x) * (1 + scale_mlp) + shift_mlp);  LN without affine, eps 1e-6; MLP = Linear -> GELU(tanh) -> Linear;
  * final layer: shift, scale = Linear(D, 2D)(SiLU(c)).chunk(2); Linear(D, p*p*C_out)(LN(x) * (1 + scale) + shift); unpatchify
    `nhwpqc -> nchpwq`; with learn_sigma C_out = 2C and the first C channels are the epsilon prediction.
"""
from __future__ import annotations

import math
import types
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

DIT_XL_2_CONFIG = dict(input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16, mlp_ratio=4.0,
                       num_classes=1000, learn_sigma=True, frequency_embedding_size=256)


def sincos_pos_embed_2d(dim: int, grid: int) -> torch.Tensor:
    """Fixed 2-D sin-cos position embedding [grid*grid, dim]: half of the channels encode one axis, half the other."""
    def one_axis(d, pos):
        omega = 1.0 / 10000 ** (torch.arange(d // 2, dtype=torch.float64) / (d / 2.0))
        out = pos.reshape(-1)[:, None] * omega[None, :]
        return torch.cat([out.sin(), out.cos()], dim=1)

    gh = torch.arange(grid, dtype=torch.float64)
    gw = torch.arange(grid, dtype=torch.float64)
    ww, hh = torch.meshgrid(gw, gh, indexing="xy")  # first grid varies along w
    emb = torch.cat([one_axis(dim // 2, ww), one_axis(dim // 2, hh)], dim=1)
    return emb.float()


def timestep_frequencies(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t.float()[:, None] * freqs[None]
    return torch.cat([a.cos(), a.sin()], dim=-1)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden: int, freq: int = 256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(freq, hidden), nn.SiLU(), nn.Linear(hidden, hidden))
        self.freq = freq

    def forward(self, t):
        return self.mlp(timestep_frequencies(t, self.freq))


class LabelEmbedder(nn.Module):
    def __init__(self, num_classes: int, hidden: int):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + 1, hidden)
        self.num_classes = num_classes

    def forward(self, labels):
        return self.embedding_table(labels)


class Attention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, T, D = x.shape
        q, k, v = self.qkv(x).reshape(B, T, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4)
        s = (q @ k.transpose(-1, -2)) * (D // self.heads) ** -0.5
        o = torch.softmax(s, dim=-1) @ v
        return self.proj(o.transpose(1, 2).reshape(B, T, D))


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x), approximate="tanh"))


def modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


class DiTBlock(nn.Module):
    def __init__(self, hidden: int, heads: int, mlp_ratio: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden, elementwise_affine=False, eps=1e-6)
        self.attn = Attention(hidden, heads)
        self.norm2 = nn.LayerNorm(hidden, elementwise_affine=False, eps=1e-6)
        self.mlp = Mlp(hidden, int(hidden * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden, 6 * hidden))

    def forward(self, x, c):
        sa, ca, ga, sm, cm, gm = self.adaLN_modulation(c).chunk(6, dim=1)
        x = x + ga.unsqueeze(1) * self.attn(modulate(self.norm1(x), sa, ca))
        x = x + gm.unsqueeze(1) * self.mlp(modulate(self.norm2(x), sm, cm))
        return x


class FinalLayer(nn.Module):
    def __init__(self, hidden: int, patch: int, out_channels: int):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden, patch * patch * out_channels)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden, 2 * hidden))

    def forward(self, x, c):
        shift, scale = self.adaLN_modulation(c).chunk(2, dim=1)
        return self.linear(modulate(self.norm_final(x), shift, scale))


class PatchEmbed(nn.Module):
    def __init__(self, patch: int, in_channels: int, hidden: int):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, hidden, patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class DiT(nn.Module):
    def __init__(self, **cfg):
        super().__init__()
        c = dict(DIT_XL_2_CONFIG)
        c.update(cfg)
        self.config = types.SimpleNamespace(**c)
        D, p = c["hidden_size"], c["patch_size"]
        self.out_channels = c["in_channels"] * (2 if c["learn_sigma"] else 1)
        self.x_embedder = PatchEmbed(p, c["in_channels"], D)
        self.t_embedder = TimestepEmbedder(D, c["frequency_embedding_size"])
        self.y_embedder = LabelEmbedder(c["num_classes"], D)
        grid = c["input_size"] // p
        self.register_buffer("pos_embed", sincos_pos_embed_2d(D, grid)[None], persistent=False)
        self.blocks = nn.ModuleList([DiTBlock(D, c["num_heads"], c["mlp_ratio"]) for _ in range(c["depth"])])
        self.final_layer = FinalLayer(D, p, self.out_channels)

    def init_weight(self):
        """DiT initialisation: xavier Linear weights, zero biases, N(0, 0.02) embeddings, zero adaLN / output layers."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        w = self.x_embedder.proj.weight
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.zeros_(self.x_embedder.proj.bias)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for blk in self.blocks:
            nn.init.zeros_(blk.adaLN_modulation[-1].weight)
            nn.init.zeros_(blk.adaLN_modulation[-1].bias)
        nn.init.zeros_(self.final_layer.adaLN_modulation[-1].weight)
        nn.init.zeros_(self.final_layer.adaLN_modulation[-1].bias)
        nn.init.zeros_(self.final_layer.linear.weight)
        nn.init.zeros_(self.final_layer.linear.bias)

    def unpatchify(self, x):
        c, p = self.out_channels, self.config.patch_size
        h = w = int(x.shape[1] ** 0.5)
        x = x.reshape(x.shape[0], h, w, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, h * p, w * p)

    def forward(self, sample, timestep, class_labels: Optional[torch.Tensor] = None, added_cond_kwargs=None, **_):
        if class_labels is None:
            class_labels = (added_cond_kwargs or {})["class_labels"]
        x = self.x_embedder(sample) + self.pos_embed
        c = self.t_embedder(timestep) + self.y_embedder(class_labels)
        for blk in self.blocks:
            x = blk(x, c)
        x = self.unpatchify(self.final_layer(x, c))
        return (x[:, : self.config.in_channels],)  # learn_sigma: the first C channels are the epsilon prediction


def tiny_config(**over):
    """A small DiT the kernels support (hidden multiple of 8; head dim 72 like DiT-XL) for parity tests."""
    c = dict(input_size=16, patch_size=2, in_channels=4, hidden_size=144, depth=2, num_heads=2, mlp_ratio=4.0, num_classes=10,
             learn_sigma=True, frequency_embedding_size=64)
    c.update(over)
    return c
