"""`DMTrainer` — drop-in for `duwu.trainer.DMTrainer` (src/duwu/trainer/trainer.py:95-318) without Lightning.

Same constructor keywords, same attributes (`unet, te, vae, lycoris_model, loss, ema_loss, n_diffusion_time_steps,
train_params`), same `training_step(batch, idx) -> {"loss", "aux_output"}`, `validation_step`, `configure_optimizers`,
`merge_lycoris`, LyCORIS weight dump.  What Lightning did implicitly around the step (autocast `bf16-mixed`, backward,
`gradient_clip_val`, optimizer / LR-scheduler step, DDP gradient all-reduce) is the explicit
`fit_step()` here: loss.backward() runs the hand-scheduled kernel backward, gradient buckets are all-reduced over NCCL on
a side stream while earlier blocks are still in backward, and clip + AdamW are two kernels with no host sync.
The reference logs `loss.item()` every step (2 host syncs, trainer.py:280-293); here the EMA stays on the device and the
host reads it only every `log_every_n_steps`.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Iterator, Optional

import torch
import torch.nn as nn

from .config import instantiate_any, load_any
from .data import BaseTextEncoder


class BaseTrainer(nn.Module):
    def __init__(self, *args, name: str = "", lr: float = 1e-5, optimizer="torch.optim.AdamW",
                 opt_config: Dict[str, Any] = {"weight_decay": 0.01, "betas": (0.9, 0.999)},
                 lr_scheduler="torch.optim.lr_scheduler.CosineAnnealingLR",
                 lr_scheduler_config: Dict[str, Any] = {"T_max": 100_000, "eta_min": 1e-7}, use_warm_up: bool = True,
                 warm_up_period: int = 1000, **kwargs):
        super().__init__()
        self.name = name
        self.train_params: Optional[Iterator[nn.Parameter]] = None
        self.optimizer = instantiate_any(optimizer)
        if self.optimizer is torch.optim.AdamW:  # the class itself (reference default argument) -> fused kernel
            from .optim import FusedAdamW

            self.optimizer = FusedAdamW
        self.opt_config = dict(opt_config)
        self.lr = lr
        self.lr_sch = instantiate_any(lr_scheduler)
        self.lr_sch_config = dict(lr_scheduler_config)
        self.use_warm_up = use_warm_up
        self.warm_up_period = warm_up_period
        self.global_step = 0

    def configure_optimizers(self, max_grad_norm: Optional[float] = None):
        """src/duwu/trainer/trainer.py:52-74.  `torch.optim.AdamW` resolves to the fused multi-tensor kernel."""
        assert self.train_params is not None
        kw = dict(self.opt_config)
        from .optim import FusedAdamW

        if self.optimizer is FusedAdamW and max_grad_norm is not None:
            kw["max_grad_norm"] = max_grad_norm
        optimizer = self.optimizer(list(self.train_params), lr=self.lr, **kw)
        lr_sch = self.lr_sch(optimizer, **self.lr_sch_config) if self.lr_sch is not None else None
        if self.use_warm_up:
            lr_scheduler = GradualWarmup(optimizer, self.warm_up_period, lr_sch)
        else:
            lr_scheduler = lr_sch
        if lr_scheduler is None:
            return optimizer
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": lr_scheduler, "interval": "step"}}


class GradualWarmup:
    """`warmup_scheduler.GradualWarmupScheduler(optimizer, multiplier=1, total_epoch, after_scheduler)` as the reference
    uses it (trainer.py:62-65), step for step: construction leaves lr = base * 0 / N (the first optimizer step runs at 0),
    the k-th scheduler step sets base * k / N up to k = N, step N + 1 hands over at the after-scheduler's current lr and
    later steps advance the after-scheduler."""

    def __init__(self, optimizer, total_epoch: int, after_scheduler=None):
        self.optimizer, self.total_epoch, self.after_scheduler = optimizer, total_epoch, after_scheduler
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        if after_scheduler is not None and hasattr(after_scheduler, "base_lrs"):
            self.base_lrs = list(after_scheduler.base_lrs)  # constructing it may already have touched group["lr"]
        self.last_epoch = -1
        self.finished = False
        self.step()

    def step(self):
        if self.finished and self.after_scheduler is not None:
            self.after_scheduler.step()
            return
        self.last_epoch += 1
        if self.last_epoch > self.total_epoch:
            if self.after_scheduler is not None:
                self.finished = True
                lrs = self.after_scheduler.get_last_lr()
            else:
                lrs = self.base_lrs
        else:
            lrs = [b * self.last_epoch / self.total_epoch for b in self.base_lrs]
        for g, lr in zip(self.optimizer.param_groups, lrs):
            g["lr"] = lr

    def get_last_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]

    def state_dict(self):
        sd = {"last_epoch": self.last_epoch, "finished": self.finished, "base_lrs": list(self.base_lrs),
              "total_epoch": self.total_epoch}
        if self.after_scheduler is not None:
            sd["after_scheduler"] = self.after_scheduler.state_dict()
        return sd

    def load_state_dict(self, sd):
        self.last_epoch, self.finished = sd["last_epoch"], sd["finished"]
        self.base_lrs, self.total_epoch = list(sd["base_lrs"]), sd["total_epoch"]
        if self.after_scheduler is not None and "after_scheduler" in sd:
            self.after_scheduler.load_state_dict(sd["after_scheduler"])
        if self.finished and self.after_scheduler is not None:
            lrs = self.after_scheduler.get_last_lr()
        elif self.last_epoch > self.total_epoch:
            lrs = self.base_lrs
        else:
            lrs = [b * self.last_epoch / self.total_epoch for b in self.base_lrs]
        for g, lr in zip(self.optimizer.param_groups, lrs):
            g["lr"] = lr


class DMTrainer(BaseTrainer):
    def __init__(self, model_config: dict, te_use_normed_ctx: bool = False, vae_std: Optional[float] = None,
                 vae_mean: Optional[float] = None, lycoris_config=None, *args, name: str = "", lr: float = 1e-5,
                 optimizer="torch.optim.AdamW", opt_config: Dict[str, Any] = {"weight_decay": 0.01, "betas": (0.9, 0.999)},
                 lr_scheduler="torch.optim.lr_scheduler.CosineAnnealingLR",
                 lr_scheduler_config: Dict[str, Any] = {"T_max": 100_000, "eta_min": 1e-7}, use_warm_up: bool = True,
                 warm_up_period: int = 1000, loss_config: Optional[dict] = None, device: Optional[str] = None):
        super().__init__(*args, name=name, lr=lr, optimizer=optimizer, opt_config=opt_config, lr_scheduler=lr_scheduler,
                         lr_scheduler_config=lr_scheduler_config, use_warm_up=use_warm_up, warm_up_period=warm_up_period)
        dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        with torch.device(dev):
            unet = load_any(model_config["unet"])
        self.unet = unet.to(dev)
        te = load_any(model_config.get("te"))
        vae = load_any(model_config.get("vae"))
        self.te = te.to(dev) if isinstance(te, nn.Module) else te
        self.vae = vae.to(dev) if isinstance(vae, nn.Module) else vae
        self.te_use_normed_ctx = te_use_normed_ctx
        self.vae_std = vae_std
        self.vae_mean = vae_mean or 0
        if self.vae_std is None and self.vae is not None:
            self.vae_std = 1 / self.vae.config.scaling_factor

        if isinstance(lycoris_config, str):
            import toml

            lycoris_config = toml.load(lycoris_config)
        if lycoris_config is not None:
            from .lycoris import LycorisNetwork, create_lycoris

            LycorisNetwork.apply_preset(lycoris_config["preset"])
            lycoris_model = create_lycoris(self.unet, **lycoris_config["config"])
            lycoris_model.apply_to()
        else:
            lycoris_model = None
        self.lycoris_model = lycoris_model

        self.register_buffer("ema_loss", torch.tensor(0.0, device=dev))
        self.ema_decay = 0.99
        if lycoris_model is not None:
            self.lycoris_model.train()
            self.unet.requires_grad_(False)
            self.train_params = self.lycoris_model.parameters()
        else:
            self.unet.requires_grad_(True).train()
            self.train_params = self.unet.parameters()

        if loss_config is None:
            from .loss import DiffusionLoss
            from .scheduler import EulerDiscreteScheduler

            scheduler = EulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler")
            self.loss = DiffusionLoss(scheduler)
        else:
            self.loss = instantiate_any(loss_config)
        self.n_diffusion_time_steps = self.loss.n_diffusion_time_steps
        self._hyper = None
        if type(self) is DMTrainer and any(p.requires_grad for p in self.loss.parameters()):
            raise ValueError("the loss module has trainable parameters (e.g. NNWeightedRFLoss.loss_pred_module) that DMTrainer's "
                             "optimizer would never see: use uwudiff_b200.trainer.NNWeightedLossTrainer")
        # fused sinusoidal timestep embedding from the noising kernel (diffusers Timesteps(block_out_channels[0]))
        # (the kernel emits the flip_sin_to_cos = True, freq_shift = 0 form; any other config takes the unfused path)
        cfg = self.unet.config
        if getattr(cfg, "flip_sin_to_cos", True) and getattr(cfg, "freq_shift", 0) == 0:
            self.loss.temb_dim = int(cfg.block_out_channels[0] if hasattr(cfg, "block_out_channels") else cfg.frequency_embedding_size)
        self._fit = None

    # ---- reference API ---------------------------------------------------------------------------------------
    def merge_lycoris(self):
        self.lycoris_model.restore()
        self.lycoris_model.merge_to()

    def save_lycoris_weight(self, dirpath: str = "./lycoris_weight", epoch: int = 0):
        """on_train_epoch_end (trainer.py:189-215): lycoris state dict ∪ trainable unet params -> epoch=N.pt."""
        os.makedirs(dirpath, exist_ok=True)
        model_weight = {k: v for k, v in self.unet.named_parameters() if v.requires_grad}
        lycoris_weight = {k: v.detach().clone() for k, v in self.lycoris_model.state_dict().items()} | model_weight
        path = os.path.join(dirpath, f"epoch={epoch}.pt")
        torch.save(lycoris_weight, path)
        return path

    def get_latent_and_conditioning(self, batch):
        x, captions, tokenizer_outputs, added_cond, cross_attn_kwargs = batch
        dev = self.ema_loss.device
        x = x.to(dev, non_blocking=True)
        added_cond = {k: v.to(dev, non_blocking=True) for k, v in added_cond.items()}
        attn_mask = None
        with torch.no_grad():
            if self.vae is not None:
                latent_dist = self.vae.encode(x).latent_dist
                x = latent_dist.sample()
                x = (x - self.vae_mean) / self.vae_std
            if self.te is None:  # class-conditional denoisers (DiT): labels travel in added_cond["class_labels"]
                return x, None, None, added_cond, cross_attn_kwargs
            if isinstance(self.te, BaseTextEncoder):
                try:
                    embedding, normed_embedding, pooled_embedding, attn_mask = self.te(tokenizer_outputs, batch_size=x.shape[0])
                except TypeError:
                    embedding, normed_embedding, pooled_embedding, attn_mask = self.te(tokenizer_outputs)
            else:
                normed_embedding, pooled_embedding, *embeddings = self.te(**tokenizer_outputs[0], return_dict=False,
                                                                           output_hidden_states=True)
                embedding = embeddings[-1][-1]
            ctx = normed_embedding if self.te_use_normed_ctx else embedding
        added_cond["text_embeds"] = pooled_embedding
        return x, ctx, attn_mask, added_cond, cross_attn_kwargs

    def training_step(self, batch, idx, _prepared=None):
        x, ctx, attn_mask, added_cond, cross_attn_kwargs = _prepared or self.get_latent_and_conditioning(batch)
        loss, aux_output = self.loss(x, self.unet, encoder_hidden_states=ctx, encoder_attention_mask=attn_mask,
                                     added_cond_kwargs=added_cond, cross_attention_kwargs=cross_attn_kwargs)
        hyper = getattr(self, "_hyper", None)
        if hyper is not None:  # CUDA-graph capture: the decay is a device scalar refreshed before every replay
            d = hyper[-1, 0]
            self.ema_loss.mul_(d).add_(loss.detach() * (1 - d))
        else:
            ema_decay = min(self.global_step / (10 + self.global_step), self.ema_decay)
            self.ema_loss = ema_decay * self.ema_loss + (1 - ema_decay) * loss.detach()  # stays on the device, no .item()
        return {"loss": loss, "aux_output": aux_output}

    @torch.no_grad()
    def validation_step(self, batch, idx):
        x, ctx, attn_mask, added_cond, cross_attn_kwargs = self.get_latent_and_conditioning(batch)
        return self.loss(x, self.unet, encoder_hidden_states=ctx, encoder_attention_mask=attn_mask,
                         added_cond_kwargs=added_cond, cross_attention_kwargs=cross_attn_kwargs)

    # ---- what Lightning's fit loop did around training_step -----------------------------------------------------
    def setup_fit(self, gradient_clip_val: Optional[float] = None, process_group=None, seed: Optional[int] = None,
                  n_buckets: int = 0, accumulate_grad_batches: int = 1, cuda_graph: bool = False, graph_warmup_steps: int = 2):
        """`accumulate_grad_batches` is Lightning's Trainer option of the same name (the `lightning_config` block of the YAMLs
        is passed to pl.Trainer verbatim): k micro-batches share one optimizer step, each loss scaled by 1/k; the gradient
        exchange runs once, during the last micro-batch's backward (BASELINE.json configs[4]: global batch 128 on fewer GPUs
        = micro-batches of 16 per GPU)."""
        from .parallel import GradientBuckets

        opt = self.configure_optimizers(max_grad_norm=gradient_clip_val)
        sched = None
        if isinstance(opt, dict):
            sched = opt["lr_scheduler"]["scheduler"]
            opt = opt["optimizer"]
        buckets = None
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                         and torch.distributed.get_world_size() > 1):
            buckets = GradientBuckets(self, process_group=process_group, n_buckets=n_buckets)
        if seed is not None:
            rank = torch.distributed.get_rank() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
            self.loss.seed = int(seed) + rank  # pl.seed_everything(seed + global_rank), test_scripts/test_train.py:68-69
        self._fit = dict(opt=opt, sched=sched, buckets=buckets, accum=max(1, int(accumulate_grad_batches)), micro=0, graph=None)
        if cuda_graph:
            from .optim import FusedAdamW

            if not isinstance(opt, FusedAdamW):
                raise NotImplementedError("cuda_graph=True needs the fused AdamW (its lr / step are read from device memory)")
            self._fit["graph"] = dict(state="warmup", seen=0, warm=max(1, int(graph_warmup_steps)))
        return self._fit

    def fit_step(self, batch, idx: int = 0):
        """forward + backward + (DDP all-reduce) + clip + optimizer + lr schedule for one batch; returns the step dict.
        With `setup_fit(cuda_graph=True)` the first `graph_warmup_steps` optimizer steps run eagerly, then forward + backward
        and clip + AdamW are captured once as two CUDA graphs and replayed (shapes are static)."""
        if self._fit is None:
            self.setup_fit()
        f = self._fit
        g = f["graph"]
        if g is None or g["state"] == "warmup":
            out, stepped = self._fit_step_eager(batch, idx)
            if g is not None and stepped:
                g["seen"] += 1
                if g["seen"] >= g["warm"]:
                    g["state"] = "capture"
            return out
        return self._fit_step_graph(batch, idx)

    def _fit_step_eager(self, batch, idx: int = 0):
        f = self._fit
        k = f["accum"]
        last = (f["micro"] + 1) % k == 0
        if f["buckets"] is not None:
            f["buckets"].enabled = last  # earlier micro-batches only accumulate locally
            if last:
                f["buckets"].begin_step()
        if k > 1 and f["micro"] % k != 0 and self.lycoris_model is not None:
            object.__setattr__(self.lycoris_model, "skip_next_fold", True)  # adapters untouched since the window's first forward
        out = self.training_step(batch, idx)
        (out["loss"] if k == 1 else out["loss"] / k).backward()
        f["micro"] += 1
        if not last:
            return out, False
        if f["buckets"] is not None:
            f["buckets"].finish()
        f["opt"].step()
        if f["sched"] is not None:
            f["sched"].step()
        if self.lycoris_model is not None:
            self.lycoris_model.zero_grad()  # one fill over the flat gradient buffer instead of one per adapter tensor
        else:
            f["opt"].zero_grad(set_to_none=False)  # gradients keep their (flat) storage
        self.global_step += 1
        return out, True

    # ---- CUDA-graph replay of the step ---------------------------------------------------------------------------
    def _graph_capture(self, prepared):
        """Capture (a) noising -> UNet forward -> loss -> backward -> EMA and (b) clip + AdamW + gradient reset.  Everything
        that changes from step to step is read from device memory: the noise-stream position (`loss._step_dev`, advanced
        inside graph (a)), lr / bias corrections / EMA decay (`self._hyper`, refreshed by a small H2D copy before each replay),
        the inputs (static buffers the batch is copied into)."""
        from . import ops

        f = self._fit
        g = f["graph"]
        x, ctx, attn_mask, added_cond, cak = prepared
        dev = x.device
        g["x"] = x.clone()
        g["ctx"] = None if ctx is None else ctx.clone()
        g["added"] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in added_cond.items()}
        g["mask"], g["cak"] = attn_mask, cak
        n_groups = len(f["opt"].param_groups)
        self._hyper = torch.zeros((n_groups + 1, 4), device=dev, dtype=torch.float32)
        self._write_hyper()
        self.loss._step_dev = torch.zeros((1,), device=dev, dtype=torch.int64)
        if f["buckets"] is not None:
            f["buckets"].enabled = False  # the exchange runs between the two graphs, on the flat gradient buffer
        k = f["accum"]
        # autograd leaves created before the capture carry the (legacy) stream they were created on: their gradient
        # accumulation would be scheduled there, which stream capture forbids -> the denoiser's graph hook is re-created
        # inside the capture
        if hasattr(self.unet, "_hook"):
            self.unet._hook = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()  # the graphs' private pool needs the room the eager steps' cached blocks occupy
        n0 = ops.launch_count()
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            out = self.training_step(None, 0, _prepared=(g["x"], g["ctx"], g["mask"], g["added"], g["cak"]))
            (out["loss"] if k == 1 else out["loss"] / k).backward()
            self.loss._step_dev += 1
        n1 = ops.launch_count()
        ga2, out2 = None, None
        if k > 1 and self.lycoris_model is not None:
            # micro-batches 2..k of an accumulation window: the same graph without the adapter fold (operands are still valid)
            if hasattr(self.unet, "_hook"):
                self.unet._hook = None
            ga2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga2, pool=ga.pool()):
                object.__setattr__(self.lycoris_model, "skip_next_fold", True)
                out2 = self.training_step(None, 0, _prepared=(g["x"], g["ctx"], g["mask"], g["added"], g["cak"]))
                (out2["loss"] / k).backward()
                self.loss._step_dev += 1
        n1b = ops.launch_count()
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, pool=ga.pool()):
            tables = f["opt"]._tables or []
            f["opt"].step(hyper_dev=[self._hyper[t["gi"]] for t in tables] if tables else None)
            f["opt"]._step -= 1  # capture does not execute: the counter advances once per replay (below)
            if self.lycoris_model is not None:
                self.lycoris_model.zero_grad()
            else:
                f["opt"].zero_grad(set_to_none=False)
        n2 = ops.launch_count()
        ops.add_graph_launches(-(n2 - n0))  # recorded, not executed: only replays count as launches
        n2 -= n1b - n1  # (launch counts per graph: fwdbwd n1 - n0, its fold-free twin n1b - n1, optimizer step n2 - n1b)
        # host staging buffers of tables uploaded inside the capture are re-read by every replay: keep them alive
        keep = list(self.lycoris_model._grad_tables.values()) if self.lycoris_model is not None else []
        g.update(fwdbwd=ga, optstep=gb, n_fwdbwd=n1 - n0, n_opt=n2 - n1, out=out, state="replay", keep=keep, fwdbwd_nofold=ga2,
                 out_nofold=out2, n_fwdbwd_nofold=n1b - n1)

    def _write_hyper(self):
        """[lr, 1 - b1^t, sqrt(1 - b2^t), 0] per param group + [ema decay, 0, 0, 0]: one pinned H2D copy per step."""
        opt = self._fit["opt"]
        rows = [opt.hyper_values(gi) + [0.0] for gi in range(len(opt.param_groups))]
        rows.append([min(self.global_step / (10 + self.global_step), self.ema_decay), 0.0, 0.0, 0.0])
        self._hyper.copy_(torch.tensor(rows, dtype=torch.float32).pin_memory(), non_blocking=True)

    def _fit_step_graph(self, batch, idx: int = 0):
        from . import ops

        f = self._fit
        g = f["graph"]
        prepared = self.get_latent_and_conditioning(batch)
        if g["state"] == "capture":
            if f["opt"]._tables is None:
                raise RuntimeError("cuda_graph: the optimizer has not stepped yet (graph_warmup_steps >= 1)")
            self._graph_capture(prepared)
        x, ctx, _mask, added_cond, _cak = prepared
        g["x"].copy_(x, non_blocking=True)
        if ctx is not None:
            g["ctx"].copy_(ctx, non_blocking=True)
        for kk, v in added_cond.items():
            if torch.is_tensor(v):
                g["added"][kk].copy_(v, non_blocking=True)
        k = f["accum"]
        last = (f["micro"] + 1) % k == 0
        self._write_hyper()
        first = f["micro"] % k == 0
        if first or g.get("fwdbwd_nofold") is None:
            g["fwdbwd"].replay()
            ops.add_graph_launches(g["n_fwdbwd"])
            out = g["out"]
        else:
            g["fwdbwd_nofold"].replay()
            ops.add_graph_launches(g["n_fwdbwd_nofold"])
            out = g["out_nofold"]
        f["micro"] += 1
        if not last:
            return out
        if f["buckets"] is not None:
            f["buckets"]._reduce(f["buckets"].flat)  # one exchange of the whole flat buffer, ordered on this stream
        g["optstep"].replay()
        ops.add_graph_launches(g["n_opt"])
        f["opt"]._step += 1
        if f["sched"] is not None:
            f["sched"].step()
        self.global_step += 1
        return out

    # ---- checkpoint / resume of what Lightning's .ckpt carries besides the weights ----------------------------------
    def fit_state_dict(self) -> Dict[str, Any]:
        """Optimizer moments + step counter, LR-scheduler state, global step and the noise-stream position."""
        if self._fit is None:
            self.setup_fit()
        f = self._fit
        return {"optimizer": f["opt"].state_dict(), "lr_scheduler": f["sched"].state_dict() if f["sched"] is not None else None,
                "global_step": self.global_step, "loss_step": getattr(self.loss, "_step", 0), "ema_loss": float(self.ema_loss)}

    def load_fit_state_dict(self, sd: Dict[str, Any]):
        if self._fit is None:
            self.setup_fit()
        f = self._fit
        f["opt"].load_state_dict(sd["optimizer"])
        if f["sched"] is not None and sd.get("lr_scheduler") is not None:
            f["sched"].load_state_dict(sd["lr_scheduler"])
        self.global_step = int(sd["global_step"])
        if hasattr(self.loss, "_step"):
            self.loss._step = int(sd.get("loss_step", 0))
        self.ema_loss.fill_(float(sd.get("ema_loss", 0.0)))


class NNWeightedLossTrainer(DMTrainer):
    """Drop-in for `duwu.trainer.nn_weighted_loss_trainer.NNWeightedLossTrainer` (nn_weighted_loss_trainer.py:12-95): the loss
    module (an `NNWeightedRFLoss` with its `loss_pred_module`) is trained too, in its own parameter group with
    `loss_opt_config`.  (The reference class passes `lycoris_model=` to a `DMTrainer` that takes `lycoris_config=` and is
    un-constructible as shipped; here the keyword is `lycoris_config`.)"""

    def __init__(self, *args, loss_opt_config: Dict[str, Any] = {"lr": 1e-3, "weight_decay": 0, "betas": (0.9, 0.999)}, **kwargs):
        super().__init__(*args, **kwargs)
        self.loss = self.loss.to(self.ema_loss.device)
        self.loss.requires_grad_(True).train()
        self.loss_params = list(self.loss.parameters())
        self.loss_opt_config = dict(loss_opt_config)
        self.train_params = list(self.train_params)

    def configure_optimizers(self, max_grad_norm: Optional[float] = None):
        kw = {}
        from .optim import FusedAdamW

        if self.optimizer is FusedAdamW and max_grad_norm is not None:
            kw["max_grad_norm"] = max_grad_norm
        optimizer = self.optimizer([{"params": self.loss_params, **self.loss_opt_config},
                                    {"params": self.train_params, "lr": self.lr, **self.opt_config}], **kw)
        lr_sch = self.lr_sch(optimizer, **self.lr_sch_config) if self.lr_sch is not None else None
        lr_scheduler = GradualWarmup(optimizer, self.warm_up_period, lr_sch) if self.use_warm_up else lr_sch
        if lr_scheduler is None:
            return optimizer
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": lr_scheduler, "interval": "step"}}
