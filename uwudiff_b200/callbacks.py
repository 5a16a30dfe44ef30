"""Validation-side statistics of the training path: `duwu.trainer.callbacks.PlotValLossPerTimestep`
(/root/reference/src/duwu/trainer/callbacks.py:48-158), SURVEY.md §8(f) rank 4.

The reference accumulates, for every validation batch, count / sum / sum of squares of the per-sample losses per diffusion
timestep with a Python loop over ALL N_t timesteps and boolean masks (`:83-92`: 3 N_t tiny kernels and masks per batch),
gathers the three vectors over the ranks at epoch end (`:96-105`) and plots mean +- std (`:111-158`).  Here the per-batch
accumulation is ONE scatter-add launch (`uwu_timestep_hist`), the cross-rank reduction one all-reduce of a [3, N_t] buffer;
the statistics (`:111-126`) are returned as tensors — drawing the figure and logging it is left to the caller's logger (the
Lightning / wandb control plane is out of scope).  Hook names and call signatures are those of the Lightning callback.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class PlotValLossPerTimestep:
    def __init__(self, n_diffusion_time_steps: Optional[int] = None, loss_key: str = "losses"):
        self.n_diffusion_time_steps = n_diffusion_time_steps
        self.loss_key = loss_key
        self._buf = None

    def _n(self, pl_module) -> int:
        return int(self.n_diffusion_time_steps or pl_module.n_diffusion_time_steps)

    def on_validation_epoch_start(self, trainer, pl_module):
        dev = pl_module.ema_loss.device if hasattr(pl_module, "ema_loss") else next(pl_module.parameters()).device
        self._buf = torch.zeros((3, self._n(pl_module)), device=dev, dtype=torch.float32)  # counts, sums, squared sums

    @property
    def validation_timestep_counts(self):
        return self._buf[0]

    @property
    def validation_timestep_losses(self):
        return self._buf[1]

    @property
    def validation_timestep_squared_losses(self):
        return self._buf[2]

    def on_validation_batch_end(self, trainer, pl_module, outputs, batch, idx):
        _, aux_output = outputs
        losses = getattr(aux_output, self.loss_key)
        ops.timestep_hist(losses, aux_output.timesteps, self._buf[0], self._buf[1], self._buf[2])

    def on_validation_epoch_end(self, trainer, pl_module):
        """Returns (timesteps, mean loss, std of the loss) over the timesteps that were drawn; identical on every rank."""
        buf = self._buf
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            buf = buf.clone()
            torch.distributed.all_reduce(buf)  # == sum over ranks of all_gather (callbacks.py:96-105)
        counts, sums, sq = buf[0], buf[1], buf[2]
        valid = counts > 0
        n = counts[valid]
        mean = sums[valid] / n
        std = torch.sqrt(torch.clamp(sq[valid] / n - mean ** 2, min=0))
        return torch.nonzero(valid).flatten(), mean, std
