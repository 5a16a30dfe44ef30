"""Generate tests/golden/loss_golden.npz by running the REFERENCE's loss code verbatim (container only).

    python -m oracle.make_golden

For each case: seed torch, run `DiffusionLoss.forward(x, unet)` from /root/reference/src/duwu/loss/diffusion.py with
a fixed stand-in denoiser (unet(x_t, t) = 0.5 * x_t), and record inputs (x0, the eps/t the reference drew —
recovered by replaying randn_like -> randint in the reference's order, src/duwu/loss/diffusion.py:75-76,68-70) and
outputs (noisy_latent, target, losses, loss).  Also records the scheduler tables and the in-tree known answers
(sigma_max = 14.6146, configs/sampling/demo_sampling.yaml:49).  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import diffusers_shim, ref_loss

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "loss_golden.npz")

CASES = [
    # name, target/pred type, use_snr, use_debiased, dtype, shape
    ("eps_plain", "epsilon", False, False, torch.float32, (4, 4, 8, 8)),
    ("eps_minsnr", "epsilon", True, False, torch.float32, (4, 4, 8, 8)),
    ("eps_minsnr_debiased", "epsilon", True, True, torch.float32, (4, 4, 32, 32)),
    ("v_plain", "v_prediction", False, False, torch.float32, (4, 4, 8, 8)),
    ("v_minsnr", "v_prediction", True, False, torch.float32, (5, 4, 8, 8)),
    ("sample_plain", "sample", False, False, torch.float32, (3, 4, 8, 8)),
    ("rf_plain", "rectified_flow", False, False, torch.float32, (3, 4, 8, 8)),
    ("eps_bf16", "epsilon", True, True, torch.bfloat16, (4, 4, 8, 8)),
    ("v_bf16", "v_prediction", True, False, torch.bfloat16, (4, 4, 8, 8)),
    ("ragged_fp32", "epsilon", True, True, torch.float32, (3, 3, 5, 7)),  # n_per = 105, not a multiple of 4
    ("pixel_c1", "epsilon", False, False, torch.float32, (4, 3, 32, 32)),  # BASELINE configs[0] shape
]


MIXED = [("v_prediction", "epsilon"), ("epsilon", "sample"), ("sample", "v_prediction"), ("rectified_flow", "epsilon"),
         ("epsilon", "rectified_flow"), ("v_prediction", "sample")]


RF_CASES = [
    # name, time_sampling_type, prediction_type, paired-noise input, rescale image+noise
    ("rf_time_rf", "uniform_time", "rectified_flow", False, False),
    ("rf_time_eps", "uniform_time", "epsilon", False, False),
    ("rf_time_v_paired", "uniform_time", "v_prediction", True, False),
    ("rf_timestep_rf", "uniform_timestep", "rectified_flow", False, False),
    ("rf_time_sample_rescaled", "uniform_time", "sample", False, True),
]


def unet_stub(x, t, **kw):
    return (0.5 * x,)


def main():
    assert ref_loss.available(), "needs /root/reference"
    mod = ref_loss.load_reference_loss_module()
    out = {}
    sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler")
    loss0 = mod.DiffusionLoss(sch)
    out["tab_acp"] = sch.alphas_cumprod.numpy()
    out["tab_sigmas"] = sch.sigmas.numpy()
    out["tab_timesteps"] = sch.timesteps.numpy()
    out["tab_snr"] = sch.all_snr.numpy()
    out["tab_sigma_by_t"] = loss0.get_sigmas_for_timesteps(torch.arange(1000)).numpy()
    for i, (name, ttype, snr, deb, dtype, shape) in enumerate(CASES):
        sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("x", prediction_type=ttype)
        L = mod.DiffusionLoss(sch, use_snr_weight=snr, use_debiased_estimation=deb, prediction_type=ttype, target_type=ttype)
        g = torch.Generator().manual_seed(1215 + i)
        x0 = torch.randn(shape, generator=g).to(dtype)
        torch.manual_seed(777 + i)
        loss, aux = L(x0, unet_stub)
        # replay the reference's RNG order to recover eps (randn_like, then randint)
        torch.manual_seed(777 + i)
        eps = torch.randn_like(x0)
        t = torch.randint(0, 1000, (shape[0],))
        assert torch.equal(t, aux.timesteps)
        out[f"{name}/x0"] = x0.float().numpy()
        out[f"{name}/eps"] = eps.float().numpy()
        out[f"{name}/t"] = t.numpy()
        out[f"{name}/x_t"] = aux.noisy_latent.float().numpy()
        out[f"{name}/target"] = aux.target.float().numpy()
        out[f"{name}/pred"] = aux.pred.float().numpy()
        out[f"{name}/losses"] = aux.losses.float().numpy()
        out[f"{name}/loss"] = np.float32(loss.float().item())
        out[f"{name}/meta"] = np.array([ttype, str(int(snr)), str(int(deb)), str(dtype).replace("torch.", "")])
    # prediction type != target type: get_prediction_for_training -> get_x0_eps_from_pred -> get_target (:100-139)
    for j, (ptype, ttype) in enumerate(MIXED):
        name = f"mixed_{ptype}_to_{ttype}"
        sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("x", prediction_type=ptype)
        L = mod.DiffusionLoss(sch, prediction_type=ptype, target_type=ttype)
        g = torch.Generator().manual_seed(3000 + j)
        x0 = torch.randn((4, 4, 8, 8), generator=g)
        torch.manual_seed(4000 + j)
        loss, aux = L(x0, unet_stub)
        torch.manual_seed(4000 + j)
        eps = torch.randn_like(x0)
        t = torch.randint(0, 1000, (4,))
        assert torch.equal(t, aux.timesteps)
        out[f"{name}/x0"] = x0.numpy()
        out[f"{name}/eps"] = eps.numpy()
        out[f"{name}/t"] = t.numpy()
        out[f"{name}/x_t"] = aux.noisy_latent.numpy()
        out[f"{name}/target"] = aux.target.numpy()
        out[f"{name}/pred"] = aux.pred.numpy()
        out[f"{name}/losses"] = aux.losses.numpy()
        out[f"{name}/loss"] = np.float32(loss.item())
        out[f"{name}/meta"] = np.array([ptype, ttype])
    # RectifiedFlowLoss (src/duwu/loss/rectified_flow.py) run verbatim
    rf = ref_loss.load_reference_rf_module()
    for j, (name, sampling, ptype, paired, rescale) in enumerate(RF_CASES):
        sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("x", prediction_type=ptype)
        L = rf.RectifiedFlowLoss(time_sampling_type=sampling, rescale_image=rescale, rescale_noise=rescale, scheduler=sch,
                                 prediction_type=ptype)
        g = torch.Generator().manual_seed(5000 + j)
        x_in = torch.randn((4, 2, 4, 8, 8) if paired else (4, 4, 8, 8), generator=g)
        torch.manual_seed(6000 + j)
        loss, aux = L(x_in, unet_stub)
        torch.manual_seed(6000 + j)  # replay: randn_like (unless paired) -> rand / randint
        noises = x_in[:, 1] if paired else torch.randn_like(x_in)
        if sampling == "uniform_time":
            smax = sch.sigmas[0]
            time = torch.rand(4) * (smax / (1 + smax))
            out[f"{name}/time"] = time.numpy()
        else:
            t_int = torch.randint(0, 1000, (4,))
            assert torch.equal(t_int, aux.timesteps)
        out[f"{name}/x_in"] = x_in.numpy()
        out[f"{name}/noise"] = noises.numpy()
        out[f"{name}/timesteps"] = aux.timesteps.numpy()
        out[f"{name}/x_t"] = aux.noisy_latent.numpy()
        out[f"{name}/target"] = aux.target.numpy()
        out[f"{name}/pred"] = aux.pred.numpy()
        out[f"{name}/losses"] = aux.losses.numpy()
        out[f"{name}/loss"] = np.float32(loss.item())
        out[f"{name}/meta"] = np.array([sampling, ptype, str(int(paired)), str(int(rescale))])
    # NNWeightedRFLoss (rectified_flow.py:144-203) run verbatim with a PARAMETRIC stand-in denoiser (out = a * x_t) and a
    # one-parameter loss head: records the loss and the gradients the reference sends to BOTH (the rescaled rf loss must
    # reach the denoiser, the log-loss regression the head)
    class Den(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.tensor(0.5))

        def forward(self, x, t, **kw):
            return (self.a * x,)

    class Head(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor(0.3))

        def forward(self, x_t, sigmas, **kw):
            return self.w * torch.log1p(sigmas) - 0.5

    for j, ptype in enumerate(["rectified_flow", "epsilon"]):
        name = f"nnw_{ptype}"
        sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("x", prediction_type=ptype)
        den, head = Den(), Head()
        L = rf.NNWeightedRFLoss(loss_pred_module=head, scheduler=sch, prediction_type=ptype)
        g = torch.Generator().manual_seed(7000 + j)
        x_in = torch.randn((4, 4, 8, 8), generator=g)
        torch.manual_seed(8000 + j)
        loss, aux = L(x_in, den)
        loss.backward()
        torch.manual_seed(8000 + j)
        noises = torch.randn_like(x_in)
        smax = sch.sigmas[0]
        time = torch.rand(4) * (smax / (1 + smax))
        out[f"{name}/x_in"] = x_in.numpy()
        out[f"{name}/noise"] = noises.numpy()
        out[f"{name}/time"] = time.numpy()
        out[f"{name}/loss"] = np.float32(loss.item())
        out[f"{name}/losses"] = aux.losses.detach().numpy()
        out[f"{name}/rescaled_losses"] = aux.rescaled_losses.detach().numpy()
        out[f"{name}/pred_losses"] = aux.pred_losses.detach().numpy()
        out[f"{name}/loss_pred_losses"] = aux.loss_pred_losses.detach().numpy()
        out[f"{name}/grad_denoiser"] = np.float32(den.a.grad.item())
        out[f"{name}/grad_head"] = np.float32(head.w.grad.item())
    # unsupported target type -> ValueError in the reference (:98)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(CASES), "cases")


if __name__ == "__main__":
    main()
