"""`EulerDiscreteScheduler` duck type used by the training loss (tables only).

The reference instantiates `diffusers.EulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0",
subfolder="scheduler")` from YAML (configs/demo_training_lycoris.yaml:64-67; src/duwu/trainer/trainer.py:173-177) and
only reads `.alphas_cumprod`, `.timesteps`, `.sigmas`, `.config.{prediction_type,num_train_timesteps}`,
`.get_velocity` and the writable `.all_snr` (src/duwu/loss/diffusion.py:37-51,57-62,67,90).  diffusers is not
installable here (no network), so the public scheduler_config.json constants of the named checkpoints are embedded.
"""
from __future__ import annotations

import json
import os
import types

import numpy as np
import torch

_KNOWN_CONFIGS = {
    # stabilityai/stable-diffusion-xl-base-1.0 / scheduler/scheduler_config.json
    "stabilityai/stable-diffusion-xl-base-1.0": dict(
        num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
        prediction_type="epsilon", timestep_spacing="leading", steps_offset=1, interpolation_type="linear",
        use_karras_sigmas=False, trained_betas=None),
    # runwayml/stable-diffusion-v1-5 (PNDM config re-used with Euler: same betas)
    "runwayml/stable-diffusion-v1-5": dict(
        num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
        prediction_type="epsilon", timestep_spacing="leading", steps_offset=1, interpolation_type="linear",
        use_karras_sigmas=False, trained_betas=None),
}


class EulerDiscreteScheduler:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, prediction_type: str = "epsilon", **extra):
        if trained_betas is not None:
            betas = torch.as_tensor(np.asarray(trained_betas), dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        self.betas = betas
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).flip(0)
        ts = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()
        self.timesteps = torch.from_numpy(ts).to(dtype=torch.float32)
        self.sigmas = torch.cat([sig, torch.zeros(1, dtype=torch.float32)])
        self.config = types.SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=trained_betas, prediction_type=prediction_type, **extra)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str = "stabilityai/stable-diffusion-xl-base-1.0",
                        subfolder: str | None = None, **overrides):
        path = pretrained_model_name_or_path
        cfg = None
        cand = os.path.join(path, subfolder or "", "scheduler_config.json")
        if os.path.exists(cand):
            with open(cand) as f:
                cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        elif path in _KNOWN_CONFIGS:
            cfg = dict(_KNOWN_CONFIGS[path])
        if cfg is None:
            raise OSError(f"scheduler config for '{path}' is neither a local directory nor an embedded config "
                          f"(known: {sorted(_KNOWN_CONFIGS)}); the HF hub is unreachable")
        cfg.update(overrides)
        return cls(**cfg)

    def get_velocity(self, sample: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        acp = self.alphas_cumprod.to(device=sample.device, dtype=sample.dtype)
        timesteps = timesteps.to(sample.device)
        sa = acp[timesteps] ** 0.5
        s1a = (1 - acp[timesteps]) ** 0.5
        while sa.dim() < sample.dim():
            sa, s1a = sa.unsqueeze(-1), s1a.unsqueeze(-1)
        return sa * noise - s1a * sample
