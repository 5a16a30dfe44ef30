"""CPU restatement (plain PyTorch, fp32) of the ENCODER half of diffusers' `AutoencoderKL` — TEST INFRASTRUCTURE ONLY.

The reference step calls `self.vae.encode(x).latent_dist.sample()` (src/duwu/trainer/trainer.py:241-244) on
`diffusers.AutoencoderKL.from_pretrained("madebyollin/sdxl-vae-fp16-fix")` (configs/demo_training_lycoris.yaml:112-117).
diffusers is not installable here (unpinned in pyproject.toml:23) and the reference holds no vectors for it: PARITY UNPINNED.
Published architecture restated with diffusers parameter names: Encoder(conv_in, DownEncoderBlock2D x 4 [ResnetBlock2D
without time embedding, eps 1e-6; Downsample2D(padding=0) = F.pad(0,1,0,1) + stride-2 conv], UNetMidBlock2D [resnet,
single-head Attention with GroupNorm + residual, resnet], GroupNorm, SiLU, conv_out), quant_conv, DiagonalGaussianDistribution.
Known answer pinned in tests: 34 163 664 parameters (the SDXL VAE encoder + quant_conv).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

SDXL_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                       norm_num_groups=32, scaling_factor=0.13025)


class Resnet(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Downsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample(cout)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.downsamplers is None else self.downsamplers[0](x)


class MidAttention(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(ch, ch), nn.Linear(ch, ch), nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        n = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(n), self.to_k(n), self.to_v(n)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]  # one head of dim C
        o = self.to_out[0](o).transpose(1, 2).reshape(B, C, H, W)
        return o + x


class MidBlock(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(ch, ch, groups), Resnet(ch, ch, groups)])
        self.attentions = nn.ModuleList([MidAttention(ch, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups):
        super().__init__()
        boc = tuple(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        blocks, cin = [], boc[0]
        for i, ch in enumerate(boc):
            blocks.append(DownBlock(cin, ch, layers_per_block, norm_num_groups, i != len(boc) - 1))
            cin = ch
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = MidBlock(boc[-1], norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLEncoder(nn.Module):
    def __init__(self, **cfg):
        super().__init__()
        c = dict(SDXL_VAE_CONFIG)
        c.update(cfg)
        self.config = c
        self.encoder = Encoder(c["in_channels"], c["latent_channels"], c["block_out_channels"], c["layers_per_block"],
                               c["norm_num_groups"])
        self.quant_conv = nn.Conv2d(2 * c["latent_channels"], 2 * c["latent_channels"], 1)

    def moments(self, x):
        return self.quant_conv(self.encoder(x))

    def encode_mean_logvar(self, x):
        mean, logvar = torch.chunk(self.moments(x), 2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)


class Upsample(nn.Module):
    """diffusers Upsample2D(use_conv=True): nearest 2x, then a 3x3 conv."""

    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class UpBlock(nn.Module):
    """diffusers UpDecoderBlock2D: layers_per_block + 1 resnets (no time embedding), then the upsampler."""

    def __init__(self, cin, cout, layers, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    """diffusers.models.autoencoders.vae.Decoder (restated): conv_in, UNetMidBlock2D, UpDecoderBlock2D x len(block_out_channels)
    over the REVERSED channel list (layers_per_block + 1 resnets each, nearest-2x + conv except in the last), GroupNorm + SiLU,
    conv_out."""

    def __init__(self, out_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups):
        super().__init__()
        rev = tuple(reversed(tuple(block_out_channels)))
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = MidBlock(rev[0], norm_num_groups)
        blocks, cin = [], rev[0]
        for i, ch in enumerate(rev):
            blocks.append(UpBlock(cin, ch, layers_per_block + 1, norm_num_groups, i != len(rev) - 1))
            cin = ch
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, rev[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(rev[-1], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLFull(AutoencoderKLEncoder):
    """Encoder half + post_quant_conv + decoder: `decode(z)` = decoder(post_quant_conv(z)) (diffusers AutoencoderKL.decode)."""

    def __init__(self, **cfg):
        super().__init__(**cfg)
        c = self.config
        self.decoder = Decoder(c["out_channels"], c["latent_channels"], c["block_out_channels"], c["layers_per_block"],
                               c["norm_num_groups"])
        self.post_quant_conv = nn.Conv2d(c["latent_channels"], c["latent_channels"], 1)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))
