"""§8 f4 — the sampling path on the kernel-backed denoiser: `sample_latents` (latent-space body of
src/duwu/sampling/sampling.py:17-118) with classifier-free guidance by batch doubling (cfg.py:54-127) and the Euler-ancestral
sampler (k_diffusion_euler.py:8-51), product UNet on the GPU vs the fp32 oracle UNet on the CPU through the SAME loop with
identical injected noise.  Four denoiser evaluations chained: tolerance 2e-2 on the final latents (bf16 compute)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import diffusers_shim  # noqa: E402  (checker only)
from oracle import unet_oracle as U  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    assert torch.isfinite(a).all()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def test_cfg_euler_ancestral_sampling_matches_oracle():
    from uwudiff_b200 import sampling as S
    from uwudiff_b200 import unet as P
    from uwudiff_b200.scheduler import EulerDiscreteScheduler

    torch.manual_seed(0)
    cfg = U.tiny_config()
    o = U.UNet2DConditionModel(**cfg).eval()
    p = P.UNet2DFromScratch.from_config(cfg)
    p.load_state_dict(o.state_dict())
    p = p.cuda().requires_grad_(False)
    o.requires_grad_(False)
    n, steps = 2, 4
    g = torch.Generator().manual_seed(1)
    emb, nemb = torch.randn(n, 77, cfg["cross_attention_dim"], generator=g), torch.randn(n, 50, cfg["cross_attention_dim"], generator=g)
    pool, npool = torch.randn(n, 64, generator=g), torch.randn(n, 64, generator=g)
    noise0 = torch.randn(n, 4, 16, 16, generator=g)
    step_noise = [torch.randn(n, 4, 16, 16, generator=g) for _ in range(steps)]

    def run(unet, dev):
        it = iter(step_noise)
        sch = (EulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler")
               if dev == "cuda" else diffusers_shim.EulerDiscreteScheduler.from_pretrained("x"))
        fac = lambda w: S.cfg_wrapper_from_embeddings(emb.to(dev), pool.to(dev), None, nemb.to(dev), npool.to(dev), None, 128, 128,
                                                      w, cfg=3.0)
        return S.sample_latents(unet, sch, fac, num_steps=steps, num_samples=n, seed=None, width=128, height=128, noise=noise0,
                                noise_sampler=lambda s, s_next: next(it).to(dev), vae_std=1 / 0.13025)

    ref = run(o, "cpu")
    got = run(p, "cuda")
    assert got.shape == ref.shape == (n, 4, 16, 16)
    assert rel(got, ref) < 2e-2, rel(got, ref)


def test_decode_tail_of_diffusion_sampling_matches_the_reference_postprocess():
    """Latents -> `vae.decode` one at a time -> `vae_image_postprocess` (sampling.py:117-126, data/utils.py:10-19) through the
    kernel-backed VAE decoder, against the oracle decoder + the same post-processing in fp32."""
    from oracle import vae_oracle as VO
    from uwudiff_b200 import sampling as S
    from uwudiff_b200.vae import AutoencoderKL

    torch.manual_seed(0)
    cfg = dict(block_out_channels=(64, 128), layers_per_block=1)
    o = VO.AutoencoderKLFull(**cfg).eval()
    p = AutoencoderKL(**cfg)
    p.load_state_dict(o.state_dict())
    p = p.cuda().requires_grad_(False)
    lat = torch.randn(3, 4, 8, 8, generator=torch.Generator().manual_seed(5))
    imgs = S.decode_latents(p, lat.cuda())
    with torch.no_grad():
        ref = torch.cat([o.decode(z.unsqueeze(0)) for z in lat])
    want = [((im * 0.5 + 0.5) * 255).clamp(0, 255).to(torch.uint8).permute(1, 2, 0) for im in ref]
    assert len(imgs) == 3 and imgs[0].shape == (16, 16, 3) and imgs[0].dtype == torch.uint8
    for a, b in zip(imgs, want):
        assert (a.int() - b.int()).abs().max() <= 3  # bf16 decoder vs fp32, in 8-bit pixel steps
    raw = S.decode_latents(p, lat.cuda(), to_uint8=False)
    assert raw.shape == (3, 3, 16, 16)
