"""§8 f2 / f4 — `AutoencoderKL.encode` (the frozen VAE of src/duwu/trainer/trainer.py:241-244) and `.decode` (the sampling
callback's latents -> pictures) on the sm_100a kernels against the fp32 restatement oracle/vae_oracle.py (diffusers absent:
parity unpinned; the published parameter counts of the SD / SDXL VAE — 83 653 863 in total, 34 163 664 of them encoder +
quant_conv — are checked as known answers).  Same state dict on both sides; bf16 compute vs fp32: rel <= max(1e-2, 2 x the error of
the same oracle under torch.autocast(cuda, bf16)) on the moments."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vae_oracle as VO  # noqa: E402  (checker only)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    assert torch.isfinite(a).all()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def pair(**cfg):
    from uwudiff_b200.vae import AutoencoderKL

    torch.manual_seed(0)
    o = VO.AutoencoderKLFull(**cfg).eval()
    p = AutoencoderKL(**cfg)
    r = p.load_state_dict(o.state_dict())
    assert not r.missing_keys and not r.unexpected_keys
    return o, p.cuda().requires_grad_(False)


@pytest.mark.parametrize("px,B", [(64, 2), (128, 1)])
def test_vae_encoder_matches_oracle_tiny(px, B):
    o, p = pair(block_out_channels=(64, 128, 128), layers_per_block=1)
    x = torch.randn(B, 3, px, px, generator=torch.Generator().manual_seed(px))
    with torch.no_grad():
        m_ref, lv_ref = o.encode_mean_logvar(x)
    dist = p.encode(x.cuda()).latent_dist
    assert dist.mean.shape == m_ref.shape == (B, 4, px // 4, px // 4)
    o2 = o.cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        m_y, lv_y = o2.encode_mean_logvar(x.cuda())
    for got, yard, ref in ((dist.mean, m_y, m_ref), (dist.logvar, lv_y, lv_ref)):
        assert rel(got, ref) <= max(1e-2, 2 * rel(yard, ref)), (rel(got, ref), rel(yard, ref))
    g = torch.Generator(device="cuda").manual_seed(3)
    s = dist.sample(generator=g)
    g = torch.Generator(device="cuda").manual_seed(3)
    noise = torch.randn(dist.mean.shape, generator=g, device="cuda", dtype=dist.mean.dtype)
    assert torch.equal(s, dist.mean + dist.std * noise) and torch.equal(dist.mode(), dist.mean)


def test_sdxl_vae_encoder_full_width_matches_oracle():
    """The real SDXL VAE config (128-256-512-512, 2 resnets per block, mid attention with one 512-wide head) at 256 x 256 px."""
    o, p = pair()
    assert sum(q.numel() for q in p.parameters()) == 83_653_863 == sum(q.numel() for q in o.parameters())
    enc = lambda m: sum(q.numel() for n, q in m.named_parameters() if n.startswith(("encoder.", "quant_conv.")))  # noqa: E731
    assert enc(p) == 34_163_664 == enc(o)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = o.moments(x)
    mom = p.encode(x.cuda()).latent_dist.parameters
    assert mom.shape == ref.shape == (2, 8, 32, 32)
    o = o.cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        yard = o.moments(x.cuda())
    e, y = rel(mom, ref), rel(yard, ref)
    print(f"[vae] moments: kernels {e:.3e}  torch-autocast-bf16 {y:.3e}")
    assert e <= max(1e-2, 2 * y), (e, y)


def test_trainer_latents_come_from_the_kernel_vae():
    """`get_latent_and_conditioning` with a VAE: pixels -> encode -> sample -> (x - mean) / std (trainer.py:241-244)."""
    from uwudiff_b200 import config as ucfg
    from uwudiff_b200.vae import AutoencoderKL

    prev = ucfg.use_synthetic_conditioning(False)
    try:
        with pytest.warns(UserWarning, match="RANDOM"):
            vae = ucfg.load_any({"_target_": "diffusers.AutoencoderKL.from_pretrained",
                                 "_load_config_": {"precision": "torch.float16", "to_freeze": True},
                                 "pretrained_model_name_or_path": "madebyollin/sdxl-vae-fp16-fix"})
    finally:
        ucfg.use_synthetic_conditioning(prev)
    assert isinstance(vae, AutoencoderKL) and abs(vae.config.scaling_factor - 0.13025) < 1e-9
    vae = vae.cuda()
    lat = vae.encode(torch.randn(1, 3, 64, 64, device="cuda")).latent_dist.sample()
    assert lat.shape == (1, 4, 8, 8) and lat.dtype == torch.float16 and torch.isfinite(lat).all()


@pytest.mark.parametrize("lat,B", [(8, 2), (16, 1)])
def test_vae_decoder_matches_oracle_tiny(lat, B):
    o, p = pair(block_out_channels=(64, 128, 128), layers_per_block=1)
    z = torch.randn(B, 4, lat, lat, generator=torch.Generator().manual_seed(lat))
    with torch.no_grad():
        ref = o.decode(z)
    img = p.decode(z.cuda()).sample
    assert img.shape == ref.shape == (B, 3, lat * 4, lat * 4)
    o2 = o.cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        yard = o2.decode(z.cuda())
    assert rel(img, ref) <= max(1e-2, 2 * rel(yard, ref)), (rel(img, ref), rel(yard, ref))
    assert torch.equal(p.decode(z.cuda(), return_dict=False)[0], img)


def test_sdxl_vae_decoder_full_width_matches_oracle():
    """The real SDXL VAE decoder (512-512-256-128, 3 resnets per block, mid attention) from 32 x 32 latents to 256 x 256 px, and
    the encode -> decode round trip shape contract of the sampling callback."""
    o, p = pair()
    z = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = o.decode(z)
    img = p.decode(z.cuda()).sample
    assert img.shape == ref.shape == (1, 3, 256, 256)
    o = o.cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        yard = o.decode(z.cuda())
    e, y = rel(img, ref), rel(yard, ref)
    print(f"[vae] decoded image: kernels {e:.3e}  torch-autocast-bf16 {y:.3e}")
    assert e <= max(1e-2, 2 * y), (e, y)
    lat = p.encode(img).latent_dist.mode()
    assert lat.shape == z.shape
