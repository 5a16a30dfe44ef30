"""CPU tests: the oracle restatement against the golden vectors made from the reference run verbatim,
and (where /root/reference exists) against the reference itself."""
import numpy as np
import pytest
import torch

from oracle import diffusers_shim, loss_oracle, philox, ref_loss

CASE_NAMES = ["eps_plain", "eps_minsnr", "eps_minsnr_debiased", "v_plain", "v_minsnr", "sample_plain", "rf_plain",
              "eps_bf16", "v_bf16", "ragged_fp32", "pixel_c1"]


def _tables():
    sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("sdxl")
    return sch, loss_oracle.scheduler_tables(sch)


def test_scheduler_known_answers(golden):
    sch, tab = _tables()
    # in-tree known answer: configs/sampling/demo_sampling.yaml:49 (max_sigma: 14.6146)
    assert abs(float(sch.sigmas[0]) - 14.6146) < 1e-4
    assert float(sch.sigmas[-1]) == 0.0 and sch.sigmas.numel() == 1001
    # SURVEY.md Appendix D probe values (reference loss file run verbatim)
    assert float(sch.sigmas[0]) == 14.614646911621094
    assert float(sch.sigmas[-2]) == 0.029167532920837402
    assert float(tab.acp[0]) == 0.9991499781608582
    assert float(tab.acp[999]) == 0.00466009508818388
    assert float(tab.snr[0]) == 1175.4405517578125
    assert float(tab.snr[999]) == 0.004681912250816822
    np.testing.assert_array_equal(tab.acp.numpy(), golden["tab_acp"])
    np.testing.assert_array_equal(tab.snr.numpy(), golden["tab_snr"])
    np.testing.assert_array_equal(tab.sigma_t.numpy(), golden["tab_sigma_by_t"])
    # sigma(t) == sqrt((1-acp)/acp) bit-exactly (Appendix A.1)
    np.testing.assert_array_equal(tab.sigma_t.numpy(), (((1 - tab.acp) / tab.acp) ** 0.5).numpy())


@pytest.mark.parametrize("name", CASE_NAMES)
def test_restatement_matches_golden(golden, name):
    ttype, snr, deb, dtype = golden[f"{name}/meta"]
    dt = getattr(torch, str(dtype))
    _, tab = _tables()
    x0 = torch.from_numpy(golden[f"{name}/x0"]).to(dt)
    eps = torch.from_numpy(golden[f"{name}/eps"]).to(dt)
    t = torch.from_numpy(golden[f"{name}/t"])
    loss, aux = loss_oracle.diffusion_loss(
        x0, eps, t, lambda x, t, **kw: (0.5 * x,), tab, target_type=str(ttype), prediction_type=str(ttype),
        use_snr_weight=bool(int(snr)), use_debiased=bool(int(deb)))
    # bit-exact: same ops, same order, same machine arithmetic
    np.testing.assert_array_equal(aux["noisy_latent"].float().numpy(), golden[f"{name}/x_t"])
    np.testing.assert_array_equal(aux["target"].float().numpy(), golden[f"{name}/target"])
    if dt == torch.float32:
        np.testing.assert_array_equal(aux["losses"].float().numpy(), golden[f"{name}/losses"])
        assert np.float32(loss.item()) == golden[f"{name}/loss"]
    else:
        # bf16 latents: the reference on CPU (no autocast) evaluates MSELoss in bf16; the oracle and the kernels
        # follow the CUDA-autocast policy of `precision: bf16-mixed` (mse_loss in fp32) -> bf16-level tolerance
        np.testing.assert_allclose(aux["losses"].float().numpy(), golden[f"{name}/losses"], rtol=2e-2)
        assert abs(loss.item() - float(golden[f"{name}/loss"])) < 2e-2 * abs(float(golden[f"{name}/loss"]))


MIXED = [("v_prediction", "epsilon"), ("epsilon", "sample"), ("sample", "v_prediction"), ("rectified_flow", "epsilon"),
         ("epsilon", "rectified_flow"), ("v_prediction", "sample")]


@pytest.mark.parametrize("ptype,ttype", MIXED)
def test_restatement_matches_golden_mixed_types(golden, ptype, ttype):
    """prediction_type != target_type: get_prediction_for_training -> get_x0_eps_from_pred -> get_target
    (src/duwu/loss/diffusion.py:100-139), golden vectors from the reference run verbatim."""
    name = f"mixed_{ptype}_to_{ttype}"
    _, tab = _tables()
    x0, eps, t = (torch.from_numpy(golden[f"{name}/{k}"]) for k in ("x0", "eps", "t"))
    loss, aux = loss_oracle.diffusion_loss(x0, eps, t, lambda x, tt, **kw: (0.5 * x,), tab, target_type=ttype,
                                           prediction_type=ptype)
    np.testing.assert_array_equal(aux["noisy_latent"].numpy(), golden[f"{name}/x_t"])
    np.testing.assert_array_equal(aux["target"].numpy(), golden[f"{name}/target"])
    np.testing.assert_array_equal(aux["pred"].numpy(), golden[f"{name}/pred"])
    np.testing.assert_allclose(aux["losses"].numpy(), golden[f"{name}/losses"], rtol=1e-6)


@pytest.mark.parametrize("name", ["rf_time_rf", "rf_time_eps", "rf_time_v_paired", "rf_time_sample_rescaled"])
def test_rf_time_sampling_host_logic_matches_golden(golden, name):
    """RectifiedFlowLoss.sample_timesteps_and_sigmas / sigma_to_timestep (host logic, no kernels): the fractional timesteps
    the reference produced for the recorded uniform draws."""
    from uwudiff_b200.loss import RectifiedFlowLoss
    from uwudiff_b200.scheduler import EulerDiscreteScheduler

    sch = EulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler")
    L = RectifiedFlowLoss(scheduler=sch)
    time = torch.from_numpy(golden[f"{name}/time"])
    ts, sig = L.sample_timesteps_and_sigmas(torch.zeros(time.numel(), 1), time)
    np.testing.assert_allclose(ts.numpy(), golden[f"{name}/timesteps"], rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(sig.numpy(), (time / (1 - time)).numpy(), rtol=1e-6)


def test_survey_probe_case():
    """seed 1215, x=randn(4,4,32,32), unet = 0.5*x, min-SNR + debiased."""
    _, tab = _tables()
    torch.manual_seed(1215)
    x = torch.randn(4, 4, 32, 32)
    eps = torch.randn_like(x)
    t = torch.randint(0, 1000, (4,))
    assert t.tolist() == [556, 39, 226, 371]
    loss, aux = loss_oracle.diffusion_loss(x, eps, t, lambda x, t, **kw: (0.5 * x,), tab, use_snr_weight=True,
                                           use_debiased=True)
    np.testing.assert_allclose(aux["losses"].numpy(), [0.7387838959693909, 0.03956311196088791, 0.45207294821739197,
                                                        0.5394063591957092], rtol=1e-6)
    assert abs(loss.item() - 0.44245660305023193) < 1e-6


def test_unsupported_target_type_raises():
    _, tab = _tables()
    x = torch.zeros(1, 4, 2, 2)
    with pytest.raises(ValueError):
        loss_oracle.target(x, x, torch.zeros(1, dtype=torch.long), tab, "o_prediction")


@pytest.mark.skipif(not ref_loss.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("ttype,snr,deb", [("epsilon", True, True), ("v_prediction", True, False),
                                            ("sample", False, False), ("rectified_flow", False, False)])
def test_restatement_matches_reference_verbatim(ttype, snr, deb):
    mod = ref_loss.load_reference_loss_module()
    sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("x", prediction_type=ttype)
    L = mod.DiffusionLoss(sch, use_snr_weight=snr, use_debiased_estimation=deb)
    tab = loss_oracle.scheduler_tables(sch)
    x0 = torch.randn(6, 4, 16, 16, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(99)
    loss_ref, aux = L(x0, lambda x, t, **kw: (torch.tanh(x),))
    torch.manual_seed(99)
    eps = torch.randn_like(x0)
    t = torch.randint(0, 1000, (6,))
    loss, a = loss_oracle.diffusion_loss(x0, eps, t, lambda x, t, **kw: (torch.tanh(x),), tab, target_type=ttype,
                                         prediction_type=ttype, use_snr_weight=snr, use_debiased=deb)
    assert torch.equal(a["noisy_latent"], aux.noisy_latent)
    assert torch.equal(a["target"], aux.target)
    assert torch.equal(a["losses"], aux.losses)
    assert torch.equal(loss, loss_ref)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, exp in kat:
        got = philox.philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == exp


def test_philox_sampling_properties():
    t = philox.sample_timesteps(4096, 1000, seed=1215, offset=3)
    assert t.min() >= 0 and t.max() < 1000 and len(np.unique(t)) > 900
    z = philox.normals(64, 1023, seed=1215, offset=0)
    assert z.shape == (64, 1023) and abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    # different offsets / seeds give different streams; same key reproduces
    assert not np.array_equal(z, philox.normals(64, 1023, seed=1215, offset=1))
    np.testing.assert_array_equal(z, philox.normals(64, 1023, seed=1215, offset=0))


def test_sinusoidal_embedding_shape_and_values():
    e = loss_oracle.sinusoidal_embedding(torch.tensor([0, 10, 999]), 320)
    assert e.shape == (3, 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))
    assert abs(e[1, 0].item() - np.cos(10.0)) < 1e-6 and abs(e[1, 160].item() - np.sin(10.0)) < 1e-6


def test_edm_weight_known_answers():
    """lambda(sigma) = (sigma^2 + sd^2) / (sigma sd)^2: equals 2 / sd^2 at sigma = sd, -> 1 / sd^2 for sigma -> inf,
    ~ 1 / sigma^2 for sigma -> 0 (Karras et al. 2022, Table 1)."""
    import torch

    from oracle import loss_oracle

    sig = torch.tensor([0.5, 1e4, 1e-3], dtype=torch.float32)
    tab = loss_oracle.Tables(None, sig, None)
    w = loss_oracle.edm_weight(torch.arange(3), tab, 0.5)
    assert w.shape == (2, 3) and torch.equal(w[1], torch.ones(3))
    assert abs(w[0, 0].item() - 8.0) < 1e-5
    assert abs(w[0, 1].item() - 4.0) < 1e-4
    assert abs(w[0, 2].item() * 1e-6 - 1.0) < 1e-4
