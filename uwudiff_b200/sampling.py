"""Sampling path of the reference on the kernel-backed denoiser, up to the generated LATENTS:

  * `DiscreteSchedule` / `DiscreteEpsDDPMDenoiser` — src/duwu/sampling/k_diffusion_wrapper.py:23-103 (sigma <-> fractional
    timestep by log-sigma interpolation, c_in = (sigma^2 + 1)^-1/2, denoised = x - sigma * eps),
  * `cfg_wrapper` / `cond_text_wrapper` — src/duwu/sampling/cfg.py:9-127 (classifier-free guidance by batch doubling:
    one UNet forward over [x | x] with [cond | uncond] conditioning, `uncond + (cond - uncond) * cfg`),
  * `sample_euler_ancestral` — src/duwu/sampling/k_diffusion_euler.py:8-51 with the two k-diffusion helpers it imports
    (`to_d`, `get_ancestral_step`; k-diffusion is not installable here, their published formulas are restated),
  * `sample_latents` — the latent-space body of `diffusion_sampling` (src/duwu/sampling/sampling.py:17-118): sigma ladder from
    the scheduler, `randn * sqrt(1 + sigma_0^2)`, the sampler loop, optional std rescale, `* vae_std + vae_mean`.

The VAE *decoder* (sampling.py:119-128) is not built: the training step never decodes; callers get latents.
Every UNet evaluation is the forward pass of uwudiff_b200.unet (tcgen05 GEMM / conv / flash attention kernels) under no_grad;
fractional timesteps reach the sinusoidal embedding kernel as floats.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn


def append_dims(x: torch.Tensor, target_dims: int) -> torch.Tensor:
    d = target_dims - x.ndim
    if d < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * d]


def append_zero(x: torch.Tensor) -> torch.Tensor:
    return torch.cat([x, x.new_zeros([1])])


class DiscreteSchedule(nn.Module):
    """k_diffusion_wrapper.py:23-78."""

    def __init__(self, sigmas: torch.Tensor, quantize: bool):
        super().__init__()
        self.register_buffer("sigmas", sigmas)
        self.register_buffer("log_sigmas", sigmas.log())
        self.quantize = quantize

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]

    def get_sigmas(self, n: Optional[int] = None):
        if n is None:
            return append_zero(self.sigmas.flip(0))
        t_max = len(self.sigmas) - 1
        t = torch.linspace(t_max, 0, n, device=self.sigmas.device)
        return append_zero(self.t_to_sigma(t))

    def sigma_to_t(self, sigma: torch.Tensor, quantize: Optional[bool] = None):
        quantize = self.quantize if quantize is None else quantize
        log_sigma = sigma.log()
        dists = log_sigma - self.log_sigmas[:, None]
        if quantize:
            return dists.abs().argmin(dim=0).view(sigma.shape)
        low_idx = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=self.log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = self.log_sigmas[low_idx], self.log_sigmas[high_idx]
        w = ((low - log_sigma) / (low - high)).clamp(0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.view(sigma.shape)

    def t_to_sigma(self, t: torch.Tensor):
        t = t.float()
        low_idx, high_idx, w = t.floor().long(), t.ceil().long(), t.frac()
        return ((1 - w) * self.log_sigmas[low_idx] + w * self.log_sigmas[high_idx]).exp()


class DiscreteEpsDDPMDenoiser(DiscreteSchedule):
    """k_diffusion_wrapper.py:81-103: eps-prediction model -> denoiser D(x, sigma) = x - sigma * eps(c_in x, t(sigma))."""

    def __init__(self, model: Callable, alphas_cumprod: torch.Tensor, quantize: bool):
        super().__init__(((1 - alphas_cumprod) / alphas_cumprod) ** 0.5, quantize)
        self.inner_model = model
        self.sigma_data = 1.0

    def get_scalings(self, sigma):
        return -sigma, 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5

    def get_eps(self, *args, **kwargs):
        return self.inner_model(*args, **kwargs)

    def forward(self, input, sigma, sigma_cond=None, **kwargs):
        c_out, c_in = [append_dims(x, input.ndim) for x in self.get_scalings(sigma)]
        sigma_cond = sigma_cond if sigma_cond is not None else sigma
        t = self.sigma_to_t(sigma_cond)
        eps = self.get_eps(input * c_in, t, **kwargs)
        return input + eps * c_out


def _added_cond(pool, n, width, height, time_ids, like):
    if time_ids is None:
        time_ids = torch.tensor([height, width, 0, 0, height, width]).repeat(n, 1).to(like)
    return None if pool is None else {"time_ids": time_ids.to(like), "text_embeds": pool}


def cond_wrapper_from_embeddings(emb, pool, mask, width: int, height: int, unet: DiscreteSchedule, time_ids=None):
    """cfg.py:9-51 with the text already encoded (emb [B, L, D], pooled [B, P] or None)."""
    added = _added_cond(pool, emb.size(0), width, height, time_ids, emb)

    def model_fn(x, sigma, sigma_cond=None):
        return unet(x, sigma, sigma_cond=sigma_cond, encoder_hidden_states=emb, encoder_attention_mask=mask,
                    added_cond_kwargs=added), None

    return model_fn


def cfg_wrapper_from_embeddings(emb, pool, mask, neg_emb, neg_pool, neg_mask, width: int, height: int, unet: DiscreteSchedule,
                                cfg: float = 5.0, time_ids=None):
    """cfg.py:54-127 with the prompts already encoded: pad the shorter context, stack [cond | uncond], one batched forward."""
    import torch.nn.functional as F

    if time_ids is not None:
        time_ids = time_ids.repeat(2, 1)
    added = _added_cond(None if pool is None else torch.concat([pool, neg_pool]), 2 * emb.size(0), width, height, time_ids, emb)
    if emb.size(1) > neg_emb.size(1):
        pad = (0, 0, 0, emb.size(1) - neg_emb.size(1))
        neg_emb = F.pad(neg_emb, pad)
        if neg_mask is not None:
            neg_mask = F.pad(neg_mask, pad[2:])
    if neg_emb.size(1) > emb.size(1):
        pad = (0, 0, 0, neg_emb.size(1) - emb.size(1))
        emb = F.pad(emb, pad)
        if mask is not None:
            mask = F.pad(mask, pad[2:])
    attn_mask = torch.concat([mask, neg_mask]) if (mask is not None and neg_mask is not None) else None
    ctx = torch.concat([emb, neg_emb])

    def cfg_fn(x, sigma, sigma_cond=None):
        if sigma_cond is not None:
            sigma_cond = torch.cat([sigma_cond, sigma_cond])
        cond, uncond = unet(torch.cat([x, x]), torch.cat([sigma, sigma]), sigma_cond=sigma_cond, encoder_hidden_states=ctx,
                            encoder_attention_mask=attn_mask, added_cond_kwargs=added).chunk(2)
        return uncond + (cond - uncond) * cfg, uncond

    return cfg_fn


def cfg_wrapper(prompt, neg_prompt, width: int, height: int, unet: DiscreteSchedule, te, cfg: float = 5.0, time_ids=None):
    """cfg.py:54-127: encode both prompts with the text-encoder wrapper, then `cfg_wrapper_from_embeddings`."""
    emb, normed, pool, mask = te.encode(prompt, padding=True, truncation=True)
    nemb, nnormed, npool, nmask = te.encode(neg_prompt, padding=True, truncation=True)
    if te.use_normed_ctx:
        emb, nemb = normed, nnormed
    return cfg_wrapper_from_embeddings(emb, pool, mask, nemb, npool, nmask, width, height, unet, cfg, time_ids)


def to_d(x, sigma, denoised):
    """k_diffusion.sampling.to_d (restated): the Karras ODE derivative (x - D(x)) / sigma."""
    return (x - denoised) / append_dims(sigma, x.ndim)


def get_ancestral_step(sigma_from, sigma_to, eta: float = 1.0):
    """k_diffusion.sampling.get_ancestral_step (restated): (sigma_down, sigma_up) of an ancestral step."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


@torch.no_grad()
def sample_euler_ancestral(model, x, sigmas, extra_args=None, callback=None, eta: float = 1.0, s_noise: float = 1.0,
                           noise_sampler: Optional[Callable] = None, image_to_noise: bool = False):
    """k_diffusion_euler.py:8-51."""
    extra_args = {} if extra_args is None else extra_args
    noise_sampler = (lambda s, s_next: torch.randn_like(x)) if noise_sampler is None else noise_sampler
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        sigma_cond = sigmas[i + 1] if image_to_noise else sigmas[i]
        denoised, _ = model(x, sigmas[i] * s_in, sigma_cond=sigma_cond * s_in, **extra_args)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        d = to_d(x, sigmas[i], denoised)
        x = x + d * (sigma_down - sigmas[i])
        if sigmas[i + 1] > 0:
            x = x + noise_sampler(sigmas[i], sigmas[i + 1]) * s_noise * sigma_up
    return x


def truncate_or_pad_to_length(xs: Sequence, n: int, padding_mode: str = "cycling") -> List:
    """src/duwu/utils/__init__.py:119-153."""
    xs = list(xs)
    if len(xs) >= n:
        return xs[:n]
    if padding_mode == "repeat_last":
        return xs + [xs[-1]] * (n - len(xs))
    if padding_mode == "cycling":
        return [xs[i % len(xs)] for i in range(n)]
    if padding_mode == "uniform_expansion":
        reps, rem = divmod(n, len(xs))
        out = []
        for i, v in enumerate(xs):
            out += [v] * (reps + (1 if i < rem else 0))
        return out
    raise ValueError(padding_mode)


@torch.no_grad()
def sample_latents(unet, train_scheduler, model_fn_factory: Callable, num_steps: int = 16, sample_scheduler=None,
                   get_sigma_func: Optional[Callable] = None, num_samples: int = 1, seed: Optional[int] = 42, width: int = 1024,
                   height: int = 1024, rescale: bool = False, vae_std: float = 1.0, vae_mean: float = 0.0,
                   internal_sampling_func: Optional[Callable] = None, noise: Optional[torch.Tensor] = None,
                   noise_sampler: Optional[Callable] = None):
    """Latent-space body of `diffusion_sampling` (sampling.py:41-118).  `model_fn_factory(model_wrapper)` returns the
    (cfg_output, uncond) callable (`cfg_wrapper(...)` / `cfg_wrapper_from_embeddings(...)` partially applied)."""
    if seed is not None:
        torch.manual_seed(seed)
    sampler = internal_sampling_func or sample_euler_ancestral
    ref = next(unet.parameters())
    wrapper = DiscreteEpsDDPMDenoiser(lambda *a, **k: unet(*a, **k)[0], train_scheduler.alphas_cumprod, False).to(ref.device)
    model_fn = model_fn_factory(wrapper)
    sch = sample_scheduler or train_scheduler
    if get_sigma_func is None:
        sigmas = sch.sigmas[torch.linspace(0, sch.config.num_train_timesteps, num_steps + 1).long()]
    else:
        sigmas = get_sigma_func(num_steps)
    if not isinstance(sigmas, torch.Tensor):
        sigmas = torch.tensor(sigmas)
    sigmas = sigmas.to(device=ref.device, dtype=torch.float32)
    if noise is None:
        noise = torch.randn(num_samples, unet.config.in_channels, height // 8, width // 8)
    x = noise.to(device=ref.device, dtype=torch.float32) * torch.sqrt(1 + sigmas[0] ** 2)
    kw = {} if noise_sampler is None else {"noise_sampler": noise_sampler}
    lat = sampler(model_fn, x, sigmas, **kw)
    if rescale:
        lat = lat / lat.std([1, 2, 3], keepdim=True)
    return lat * vae_std + vae_mean


def decode_latents(vae, latents: torch.Tensor, to_uint8: bool = True):
    """Tail of `diffusion_sampling` (src/duwu/sampling/sampling.py:117-126): each generated latent is decoded on its own
    (`vae.decode(latent.unsqueeze(0)).sample`) and post-processed like `vae_image_postprocess` (src/duwu/data/utils.py:10-19):
    (x * 0.5 + 0.5) * 255, clamped to [0, 255], uint8, HWC.  Returns a list of [H, W, 3] uint8 CPU tensors (PIL is the caller's
    business), or the raw [N, 3, H, W] decoder output when `to_uint8` is False."""
    outs = [vae.decode(lat.unsqueeze(0)).sample for lat in latents]
    imgs = torch.cat(outs)
    if not to_uint8:
        return imgs
    return [((im.float() * 0.5 + 0.5) * 255).cpu().clamp(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous() for im in imgs]


def diffusion_sampling(unet, vae, train_scheduler, model_fn_factory: Callable, **kw):
    """`duwu.sampling.sampling.diffusion_sampling` (sampling.py:41-126) end to end: `sample_latents(...)` -> `decode_latents(...)`."""
    return decode_latents(vae, sample_latents(unet, train_scheduler, model_fn_factory, **kw))
