"""Frozen VAE of the reference step (SURVEY.md §8 f2 encode, f4 decode): `diffusers.AutoencoderKL.from_pretrained(...)` is called as
`latent_dist = self.vae.encode(x).latent_dist; x = latent_dist.sample(); x = (x - vae_mean) / vae_std`
(src/duwu/trainer/trainer.py:241-244; configs/demo_training_lycoris.yaml:112-117, madebyollin/sdxl-vae-fp16-fix).

`AutoencoderKL` here is the drop-in for that call path: `from_pretrained`, `.config.scaling_factor`, `.encode(x).latent_dist`
with `.sample()` / `.mode()` / `.mean` / `.logvar`, `.decode(z).sample` (what the reference's sampling callback turns latents
into pictures with, src/duwu/trainer/callbacks.py), diffusers parameter names (`encoder.*`, `quant_conv.*`, `decoder.*`,
`post_quant_conv.*`).  The arithmetic runs on the
sm_100a kernels: implicit-GEMM 3x3 / stride-2 / 1x1 convolutions, fused GroupNorm(+SiLU), and the mid-block's single-head
attention (head_dim 512) as Q K^T -> row softmax -> P V on the GEMM kernel.  Forward only (the VAE is frozen).  No CPU path.

diffusers is not installable here, so the architecture is restated (PARITY UNPINNED, like oracle/unet_oracle.py): Encoder =
conv_in, DownEncoderBlock2D x len(block_out_channels) (layers_per_block resnets without time embedding, eps 1e-6, then an
asymmetric-padding stride-2 conv except in the last block), UNetMidBlock2D (resnet, attention, resnet), GroupNorm + SiLU,
conv_out to 2 * latent_channels, quant_conv 1x1.  Decoder = post_quant_conv 1x1, conv_in, UNetMidBlock2D, UpDecoderBlock2D x
len(block_out_channels) over the reversed channel list (layers_per_block + 1 resnets, nearest-2x + conv except in the last),
GroupNorm + SiLU, conv_out.  Known answer: 83 653 863 parameters for the SDXL / SD VAE config (34 163 664 of them encoder +
quant_conv).
"""
from __future__ import annotations

import json
import os
import types
import warnings
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import B_KN
from .text_encoders import _KernelModule
from .unet import Conv2d, GroupNorm, Linear, _pad_to

BF16 = torch.bfloat16

SDXL_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                       layers_per_block=2, norm_num_groups=32, act_fn="silu", scaling_factor=0.13025, sample_size=1024,
                       mid_block_add_attention=True, use_quant_conv=True)
KNOWN_VAE_CONFIGS = {"madebyollin/sdxl-vae-fp16-fix": SDXL_VAE_CONFIG,
                     ("stabilityai/stable-diffusion-xl-base-1.0", "vae"): SDXL_VAE_CONFIG}


class _Resnet(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None (restated): GN+SiLU -> conv -> GN+SiLU -> conv, + (1x1 conv) input."""

    def __init__(self, cin: int, cout: int, groups: int, eps: float = 1e-6):
        super().__init__()
        self.norm1 = GroupNorm(groups, cin, eps=eps)
        self.conv1 = Conv2d(cin, cout, 3, padding=1)
        self.norm2 = GroupNorm(groups, cout, eps=eps)
        self.conv2 = Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = Conv2d(cin, cout, 1) if cin != cout else None

    def fwd(self, x, N, H, W):
        a, _ = self.norm1.fwd(x, N, H * W, True)
        h = self.conv1.fwd3x3(a, N, H, W)
        b, _ = self.norm2.fwd(h, N, H * W, True)
        res = x if self.conv_shortcut is None else self.conv_shortcut.fwd1x1(x, N * H * W)
        return self.conv2.fwd3x3(b, N, H, W, residual=res)


class _Downsample(nn.Module):
    """Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then a 3x3 stride-2 conv, i.e. out(i, j) = sum w[ky, kx] x(2i + ky, 2j + kx)
    with zeros past the bottom / right edge — read through the four phase planes of x like the UNet's stride-2 conv."""

    _TAPS = [((ky & 1) * 2 + (kx & 1), ky >> 1, kx >> 1) for ky in range(3) for kx in range(3)]

    def __init__(self, ch: int):
        super().__init__()
        self.conv = Conv2d(ch, ch, 3, stride=2, padding=0)

    def fwd(self, x, N, H, W):
        c = self.conv._pack()
        planes = ops.phase_split2(x, N, H, W, self.conv.in_channels).view(4 * N, H // 2, W // 2, self.conv.in_channels)
        taps = [(p * N, dh, dw) for (p, dh, dw) in self._TAPS]
        return ops.conv3x3_nhwc(planes, c.fwd, taps=taps, n_out_img=N, bias=c.bias)


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([_Downsample(cout)]) if add_down else None


class _AttnToOut(nn.ModuleList):
    pass


class _MidAttention(nn.Module):
    """diffusers Attention(heads=1, dim_head=C, residual_connection=True, norm_num_groups=32, eps 1e-6, bias=True) (restated)."""

    def __init__(self, ch: int, groups: int):
        super().__init__()
        self.group_norm = GroupNorm(groups, ch, eps=1e-6)
        self.to_q, self.to_k, self.to_v = Linear(ch, ch), Linear(ch, ch), Linear(ch, ch)
        self.to_out = _AttnToOut([Linear(ch, ch), nn.Dropout(0.0)])
        self.ch = ch

    def fwd(self, x, N, H, W):
        C, L = self.ch, H * W
        n, _ = self.group_norm.fwd(x, N, L, False)
        q, k, v = self.to_q.fwd(n, N * L), self.to_k.fwd(n, N * L), self.to_v.fwd(n, N * L)
        o = torch.empty((N * L, C), device=x.device, dtype=BF16)
        scale = C ** -0.5
        s = ops._workspace((L * L + 1) // 2, x.device, "vae_attn").view(BF16)[: L * L].view(L, L)
        for b in range(N):  # one image at a time: the score matrix is L x L (head_dim 512 is outside the flash kernels)
            sl = slice(b * L, (b + 1) * L)
            ops.gemm(q[sl], k[sl], L, L, C, alpha=scale, out=s)
            ops.softmax_rows_(s)
            ops.gemm(s, v[sl], L, C, L, b_layout=B_KN, ldb=C, out=o[sl])
        return self.to_out[0].fwd(o, N * L, residual=x)


class _MidBlock(nn.Module):
    def __init__(self, ch, groups, add_attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(ch, ch, groups), _Resnet(ch, ch, groups)])
        self.attentions = nn.ModuleList([_MidAttention(ch, groups)]) if add_attention else None


class Encoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        boc = tuple(c.block_out_channels)
        for ch in boc:
            if ch % 64:
                raise NotImplementedError(f"uwudiff_b200.vae: block_out_channels must be multiples of 64 (got {boc})")
        self.conv_in = Conv2d(c.in_channels, boc[0], 3, padding=1)
        blocks, cin = [], boc[0]
        for i, ch in enumerate(boc):
            blocks.append(_DownBlock(cin, ch, c.layers_per_block, c.norm_num_groups, i != len(boc) - 1))
            cin = ch
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _MidBlock(boc[-1], c.norm_num_groups, c.mid_block_add_attention)
        self.conv_norm_out = GroupNorm(c.norm_num_groups, boc[-1], eps=1e-6)
        self.conv_out = Conv2d(boc[-1], 2 * c.latent_channels, 3, padding=1)


class _Upsample(nn.Module):
    """Upsample2D(use_conv=True): nearest 2x (uwu_upsample2x) then a 3x3 conv."""

    def __init__(self, ch: int):
        super().__init__()
        self.conv = Conv2d(ch, ch, 3, padding=1)

    def fwd(self, x, N, H, W):
        up = ops.upsample2x(x, N, H, W, self.conv.in_channels)
        return self.conv.fwd3x3(up, N, 2 * H, 2 * W)


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([_Upsample(cout)]) if add_up else None


class Decoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        rev = tuple(reversed(tuple(c.block_out_channels)))
        self.conv_in = Conv2d(c.latent_channels, rev[0], 3, padding=1)
        self.mid_block = _MidBlock(rev[0], c.norm_num_groups, c.mid_block_add_attention)
        blocks, cin = [], rev[0]
        for i, ch in enumerate(rev):
            blocks.append(_UpBlock(cin, ch, c.layers_per_block + 1, c.norm_num_groups, i != len(rev) - 1))
            cin = ch
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = GroupNorm(c.norm_num_groups, rev[-1], eps=1e-6)
        self.conv_out = Conv2d(rev[-1], c.out_channels, 3, padding=1)


class DiagonalGaussianDistribution:
    """diffusers.models.autoencoders.vae.DiagonalGaussianDistribution (restated): moments = [mean | logvar] on dim 1."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKL(_KernelModule):
    def __init__(self, **cfg):
        super().__init__()
        c = dict(SDXL_VAE_CONFIG)
        c.update({k: v for k, v in cfg.items() if not k.startswith("_")})
        self.config = types.SimpleNamespace(**c)
        self.encoder = Encoder(self.config)
        lc = 2 * self.config.latent_channels
        self.quant_conv = nn.Conv2d(lc, lc, 1) if self.config.use_quant_conv else None
        self.decoder = Decoder(self.config)
        self.post_quant_conv = Conv2d(self.config.latent_channels, self.config.latent_channels, 1) if self.config.use_quant_conv else None
        self._out_op = None

    def drop_cache(self):
        self._out_op = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, **_):
        root = os.path.join(pretrained_model_name_or_path, subfolder or "")
        cfg_file = os.path.join(root, "config.json")
        if os.path.exists(cfg_file):
            with open(cfg_file) as f:
                raw = json.load(f)
            model = cls(**{k: raw[k] for k in SDXL_VAE_CONFIG if k in raw})
            for fn in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin"):
                path = os.path.join(root, fn)
                if os.path.exists(path):
                    if fn.endswith(".safetensors"):
                        from safetensors.torch import load_file

                        sd = load_file(path)
                    else:
                        sd = torch.load(path, map_location="cpu")
                    model.load_state_dict(sd)
                    return model
            warnings.warn(f"uwudiff_b200: no weight file under {root}: VAE initialised RANDOMLY")
            return model
        for key in (pretrained_model_name_or_path, (pretrained_model_name_or_path, subfolder)):
            if key in KNOWN_VAE_CONFIGS:
                warnings.warn(f"uwudiff_b200: '{pretrained_model_name_or_path}' is not a local directory and the HF hub is "
                              "unreachable: VAE built from the embedded config with RANDOM weights")
                return cls(**KNOWN_VAE_CONFIGS[key])
        raise OSError(f"VAE '{pretrained_model_name_or_path}': neither a local directory nor an embedded config")

    def load_state_dict(self, sd, *a, **k):
        """A full diffusers AutoencoderKL checkpoint (an encoder-only one leaves the decoder as initialised: pass strict=False);
        old-style mid-block attention names (query / key / value / proj_attn) are accepted."""
        ren = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}
        out = {}
        for kk, v in sd.items():
            parts = kk.split(".")
            if "attentions" in parts:
                for old, new in ren.items():
                    if parts[-2] == old:
                        parts[-2:-1] = new.split(".")
            v2 = v
            if parts[-1] == "weight" and v.dim() == 4 and ".attentions." in kk:  # 1x1-conv style attention projections
                v2 = v.reshape(v.shape[0], v.shape[1])
            out[".".join(parts)] = v2
        r = super().load_state_dict(out, *a, **k)
        for m in self.modules():
            if hasattr(m, "drop_cache"):
                m.drop_cache()
        return r

    def _conv_out_op(self):
        """conv_out composed with quant_conv (a 1x1 conv: W' = Wq Wc, b' = Wq bc + bq), built once: one 3x3 conv launch."""
        if self._out_op is None:
            co = self.encoder.conv_out
            Wc, bc = co.weight.detach().float(), co.bias.detach().float()
            if self.quant_conv is not None:
                Wq = self.quant_conv.weight.detach().float().view(self.quant_conv.out_channels, -1)
                Wc = torch.einsum("oc,cikl->oikl", Wq, Wc)
                bc = Wq @ bc + self.quant_conv.bias.detach().float()
            Co, Ci = Wc.shape[0], Wc.shape[1]
            ci_p, co_p = _pad_to(Ci, 64), _pad_to(Co, 16)
            fwd = torch.empty((co_p, 9 * ci_p), device=Wc.device, dtype=BF16)
            ops.conv_pack(Wc.contiguous(), ci_p, co_p, _pad_to(Co, 64), fwd, None)
            bias = torch.zeros((co_p,), device=Wc.device, dtype=torch.float32)
            bias[:Co] = bc
            self._out_op = (fwd, bias, Co, ci_p)
        return self._out_op

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        ops._req_cuda(x)
        enc = self.encoder
        N, Cin, H, W = x.shape
        if H % (2 ** (len(enc.down_blocks) - 1) * 1) or W % (2 ** (len(enc.down_blocks) - 1)):
            raise ValueError(f"AutoencoderKL.encode: image size {H}x{W} must be divisible by {2 ** (len(enc.down_blocks) - 1)}")
        h = ops.nchw_to_nhwc(x, _pad_to(Cin, 64))
        h = enc.conv_in.fwd3x3(h, N, H, W)
        for blk in enc.down_blocks:
            for r in blk.resnets:
                h = r.fwd(h, N, H, W)
            if blk.downsamplers is not None:
                h = blk.downsamplers[0].fwd(h, N, H, W)
                H, W = H // 2, W // 2
        h = enc.mid_block.resnets[0].fwd(h, N, H, W)
        if enc.mid_block.attentions is not None:
            h = enc.mid_block.attentions[0].fwd(h, N, H, W)
        h = enc.mid_block.resnets[1].fwd(h, N, H, W)
        y, _ = enc.conv_norm_out.fwd(h, N, H * W, True)
        fwd, bias, Co, ci_p = self._conv_out_op()
        mom = ops.conv3x3_nhwc(y.view(N, H, W, ci_p), fwd, bias=bias, out_dtype=torch.float32)
        moments = ops.nhwc_to_nchw(mom, N, Co, H, W).to(self.out_dtype)
        dist = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (dist,)
        return types.SimpleNamespace(latent_dist=dist)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """diffusers `AutoencoderKL.decode(z).sample`: latents [N, latent_channels, h, w] (already divided by the scaling factor by
        the caller, as in the reference's sampling callback) -> images [N, out_channels, 8h, 8w]."""
        ops._req_cuda(z)
        dec = self.decoder
        N, Cz, H, W = z.shape
        M = N * H * W
        h = ops.nchw_to_nhwc(z, _pad_to(Cz, 64))
        if self.post_quant_conv is not None:
            # 1x1 conv straight into a zero-padded 64-channel buffer (the operand width of conv_in)
            c = self.post_quant_conv._pack()
            buf = torch.zeros((M, c.ci_p), device=z.device, dtype=BF16)
            ops.gemm(h, c.fwd, M, c.co_p, c.ci_p, lda=h.stride(0), bias=c.bias, out=buf[:, :c.co_p])
            h = buf
        h = dec.conv_in.fwd3x3(h, N, H, W)
        h = dec.mid_block.resnets[0].fwd(h, N, H, W)
        if dec.mid_block.attentions is not None:
            h = dec.mid_block.attentions[0].fwd(h, N, H, W)
        h = dec.mid_block.resnets[1].fwd(h, N, H, W)
        for blk in dec.up_blocks:
            for r in blk.resnets:
                h = r.fwd(h, N, H, W)
            if blk.upsamplers is not None:
                h = blk.upsamplers[0].fwd(h, N, H, W)
                H, W = 2 * H, 2 * W
        y, _ = dec.conv_norm_out.fwd(h, N, H * W, True)
        c = dec.conv_out._pack()
        img = ops.conv3x3_nhwc(y.view(N, H, W, c.ci_p), c.fwd, bias=c.bias, out_dtype=torch.float32)
        sample = ops.nhwc_to_nchw(img, N, dec.conv_out.out_channels, H, W).to(self.out_dtype)
        if not return_dict:
            return (sample,)
        return types.SimpleNamespace(sample=sample)
