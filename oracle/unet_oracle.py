"""CPU restatement (plain PyTorch, fp32) of diffusers' `UNet2DConditionModel` as used by the duwu training step
— TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's denoiser is `UNet2DFromScratch(UNet2DConditionModel)` (src/duwu/modules/unet_patch.py:13-57), i.e. the
dense math lives in the un-vendored third-party package `diffusers` (unpinned, pyproject.toml:23; the only version hint
in the reference is a comment citing v0.30.2, src/duwu/loss/rectified_flow.py:101).  diffusers is not installable here,
so this file restates its published architecture for the block types the named configs use
(DownBlock2D / CrossAttnDownBlock2D / UNetMidBlock2DCrossAttn / CrossAttnUpBlock2D / UpBlock2D, ResnetBlock2D,
Transformer2DModel with linear projections, BasicTransformerBlock, GEGLU feed-forward, text_time addition embedding)
with diffusers' parameter names so state dicts interchange.  PARITY UNPINNED for this file (no executable diffusers);
anchors inside /root/reference:
  - BasicTransformerBlock.forward flow (norm1 -> attn1 -> +res -> norm2 -> attn2 -> +res -> norm3 -> ff -> +res):
    src/duwu/modules/rope_unet.py:288-415 (in-tree patched copy of diffusers' method)
  - AttnProcessor2_0 flow (to_q/to_k/to_v -> heads -> SDPA -> to_out[0] -> to_out[1]): src/duwu/modules/rope_unet.py:76-175
  - attribute paths attn1.to_out[0], attn2.to_out[0], ff.net[-1], conv2, conv_out and their N(0, 1e-5) init:
    src/duwu/modules/unet_patch.py:15-45
  - call protocol unet(sample, timestep, encoder_hidden_states=, added_cond_kwargs={"text_embeds","time_ids"})[0]:
    src/duwu/loss/diffusion.py:172-176
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

SDXL_UNET_CONFIG = dict(
    sample_size=128, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280),
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
    layers_per_block=2, transformer_layers_per_block=(1, 2, 10), attention_head_dim=(5, 10, 20),
    cross_attention_dim=2048, use_linear_projection=True, addition_embed_type="text_time",
    addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816, norm_num_groups=32, norm_eps=1e-5,
    act_fn="silu", flip_sin_to_cos=True, freq_shift=0,
)


def get_timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, shift: float = 0.0):
    """diffusers.models.embeddings.get_timestep_embedding (restated)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, L, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        d = q.shape[-1] // self.heads
        q = q.view(B, -1, self.heads, d).transpose(1, 2)
        k = k.view(B, -1, self.heads, d).transpose(1, 2)
        v = v.view(B, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, L, self.heads * d)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), ctx) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, heads: int, dim_head: int, in_channels: int, num_layers: int, cross_attention_dim: int,
                 norm_num_groups: int = 32, use_linear_projection: bool = True):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear_projection else nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels) if use_linear_projection else nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x)
        if self.use_linear_projection:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(B, H * W, -1)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        if self.use_linear_projection:
            h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.reshape(B, H, W, -1).permute(0, 3, 1, 2))
        return h + res


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """DownBlock2D (attn=False) / CrossAttnDownBlock2D (attn=True)."""

    def __init__(self, cin, cout, temb, num_layers, add_downsample, attn, heads=1, depth=1, cross_dim=None, groups=32,
                 eps=1e-5, linear_proj=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(num_layers)])
        if attn:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, cout // heads, cout, depth, cross_dim, groups, linear_proj) for _ in range(num_layers)])
        self.has_attn = attn
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb, heads, depth, cross_dim, groups=32, eps=1e-5, linear_proj=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups, eps), ResnetBlock2D(ch, ch, temb, groups, eps)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, ch // heads, ch, depth, cross_dim, groups, linear_proj)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    """UpBlock2D (attn=False) / CrossAttnUpBlock2D (attn=True)."""

    def __init__(self, cin, cout, prev, temb, num_layers, add_upsample, attn, heads=1, depth=1, cross_dim=None, groups=32,
                 eps=1e-5, linear_proj=True):
        super().__init__()
        rs = []
        for i in range(num_layers):
            skip = cin if i == num_layers - 1 else cout
            rin = prev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(rs)
        if attn:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, cout // heads, cout, depth, cross_dim, groups, linear_proj) for _ in range(num_layers)])
        self.has_attn = attn
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionModel(nn.Module):
    def __init__(self, **cfg):
        super().__init__()
        c = dict(SDXL_UNET_CONFIG)
        c.update(cfg)
        self.config = c
        boc = tuple(c["block_out_channels"])
        n = len(boc)
        tl = c["transformer_layers_per_block"]
        tl = (tl,) * n if isinstance(tl, int) else tuple(tl)
        hd = c["attention_head_dim"]
        hd = (hd,) * n if isinstance(hd, int) else tuple(hd)  # diffusers: attention_head_dim == number of heads
        groups, eps, cross = c["norm_num_groups"], c["norm_eps"], c["cross_attention_dim"]
        lin = c["use_linear_projection"]
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        if c.get("addition_embed_type") == "text_time":
            self.add_embedding = TimestepEmbedding(c["projection_class_embeddings_input_dim"], temb)
        else:
            self.add_embedding = None
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, t in enumerate(c["down_block_types"]):
            cin, out = out, boc[i]
            self.down_blocks.append(DownBlock(cin, out, temb, c["layers_per_block"], i != n - 1, t.startswith("CrossAttn"),
                                              hd[i], tl[i], cross, groups, eps, lin))
        self.mid_block = MidBlock(boc[-1], temb, hd[-1], tl[-1], cross, groups, eps, lin)
        self.up_blocks = nn.ModuleList()
        rb, rh, rt = boc[::-1], hd[::-1], tl[::-1]
        out = rb[0]
        for i, t in enumerate(c["up_block_types"]):
            prev, out = out, rb[i]
            cin = rb[min(i + 1, n - 1)]
            self.up_blocks.append(UpBlock(cin, out, prev, temb, c["layers_per_block"] + 1, i != n - 1,
                                          t.startswith("CrossAttn"), rh[i], rt[i], cross, groups, eps, lin))
        self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], c["out_channels"], 3, padding=1)

    def init_weight(self):
        """src/duwu/modules/unet_patch.py:15-45 — N(0, 1e-5) on every residual-branch output weight and conv_out."""
        for m in self.modules():
            if isinstance(m, BasicTransformerBlock):
                nn.init.normal_(m.attn1.to_out[0].weight, 0.0, 1e-5)
                nn.init.normal_(m.attn2.to_out[0].weight, 0.0, 1e-5)
                nn.init.normal_(m.ff.net[-1].weight, 0.0, 1e-5)
            if isinstance(m, ResnetBlock2D):
                nn.init.normal_(m.conv2.weight, 0.0, 1e-5)
        nn.init.normal_(self.conv_out.weight, 0.0, 1e-5)

    def forward(self, sample, timestep, encoder_hidden_states=None, encoder_attention_mask=None, added_cond_kwargs=None,
                cross_attention_kwargs=None, **_):
        c = self.config
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], device=sample.device)
        t = t.expand(sample.shape[0]) if t.dim() == 0 else t
        t_emb = get_timestep_embedding(t, c["block_out_channels"][0], c["flip_sin_to_cos"], c["freq_shift"]).to(sample.dtype)
        emb = self.time_embedding(t_emb)
        if self.add_embedding is not None:
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            te = get_timestep_embedding(time_ids.flatten(), c["addition_time_embed_dim"], c["flip_sin_to_cos"], c["freq_shift"])
            te = te.reshape(text_embeds.shape[0], -1)
            add = torch.cat([text_embeds, te], dim=-1).to(emb.dtype)
            emb = emb + self.add_embedding(add)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return (x,)


def tiny_config(**over):
    """A small SDXL-shaped config the kernels support (channels multiple of 64, head_dim 64) for parity tests."""
    c = dict(sample_size=16, block_out_channels=(64, 128), down_block_types=("DownBlock2D", "CrossAttnDownBlock2D"),
             up_block_types=("CrossAttnUpBlock2D", "UpBlock2D"), layers_per_block=1, transformer_layers_per_block=(1, 2),
             attention_head_dim=(1, 2), cross_attention_dim=128, projection_class_embeddings_input_dim=64 + 6 * 32,
             addition_time_embed_dim=32)
    c.update(over)
    return c
