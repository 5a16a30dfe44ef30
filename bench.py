#!/usr/bin/env python
"""bench.py — the duwu diffusion training step on B200 (BASELINE.json metric: train samples/s at 1/2/4/8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--latent S]

Workload (config.workload): BASELINE.json configs[2]/[4] — SDXL UNet frozen + LyCORIS (LoKr on Attention/FeedForward,
LoRA r4 on proj_in/out, norm deltas: configs/lycoris preset of the reference), 4x128x128 latents, batch 16 PER GPU,
v-prediction + min-SNR loss, AdamW + grad-clip 1.0, data-parallel gradient all-reduce.  Weak scaling: 16 samples per GPU
at every N (N=8 is the global-batch-128 configuration of configs[4]).  A "step" = noising -> UNet forward -> weighted
MSE -> backward -> (all-reduce) -> clip + AdamW, through `DMTrainer.fit_step` (the public API).

  value : samples/s, inputs already resident in HBM, CUDA-event timed, max over ranks.
  e2e   : same step fed from pinned host memory every step (H2D inside the timed region) + a D2H read of the loss.
  roofline      : the dominant kernel (tcgen05 GEMM / implicit-GEMM conv): algorithmic FLOPs of every launch of one step
                  / their CUDA-event durations, against MEASURED_PEAKS.json bf16_tflops_sustained.
  cpu_baseline  : the CPU oracle (PyTorch restatement of the diffusers/lycoris reference path) on the host cores, on a
                  bounded sample (1 image at reduced latent size), scaled by the algorithmic-FLOP ratio.
`--impl reference` times that CPU path alone (the reference's own stack — diffusers, lycoris, lightning — is not
installable here), same metric/config/unit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LYCORIS_CONFIG = {
    "config": dict(linear_dim=4, linear_alpha=1, conv_dim=4, conv_alpha=1, algo="lora", use_tucker=True, train_norm=True),
    "preset": dict(enable_conv=False, target_module=["Transformer2DModel"], target_name=[],
                   module_algo_map={"Attention": dict(algo="lokr", factor=64, full_matrix=True),
                                    "FeedForward": dict(algo="lokr", factor=6, full_matrix=True)}),
}


def trainer_config(latent: int, batch: int):
    return {
        "seed": 1215,
        "lightning_config": {"precision": "bf16-mixed", "gradient_clip_val": 1.0},
        "data": {"_target_": "duwu.data.TrainDataModule", "_recursive_": False,
                 "dataset_config": {"_target_": "duwu.data.DummyDataset", "sample_size": [4, latent, latent], "n_samples": batch},
                 "dataloader_config": {"batch_size": batch, "num_workers": 0}},
        "trainer": {
            "_target_": "duwu.trainer.DMTrainer", "_recursive_": False, "lr": 1.0e-6, "optimizer": "torch.optim.AdamW",
            "opt_config": {"weight_decay": 0.01, "betas": [0.9, 0.999]}, "use_warm_up": False, "warm_up_period": 100,
            "lycoris_config": LYCORIS_CONFIG,
            "loss_config": {"_target_": "duwu.loss.DiffusionLoss",
                            "scheduler": {"_target_": "diffusers.EulerDiscreteScheduler.from_pretrained",
                                          "pretrained_model_name_or_path": "stabilityai/stable-diffusion-xl-base-1.0",
                                          "subfolder": "scheduler", "prediction_type": "v_prediction"},
                            "use_snr_weight": True},
            "model_config": {
                "unet": {"_target_": "duwu.modules.unet_patch.UNet2DFromScratch.from_config",
                         "_load_config_": {"precision": "torch.float32"},
                         "config": "stabilityai/stable-diffusion-xl-base-1.0", "subfolder": "unet"},
                "te": {"_target_": "duwu.modules.text_encoders.ConcatTextEncoders",
                       "_load_config_": {"precision": "torch.float16", "to_freeze": True}},
                "vae": None,
            },
        },
    }


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows)}


# ======================================================================================================
# CPU oracle arm (cpu_baseline / --impl reference)
# ======================================================================================================
class CpuOracleStep:
    """The reference path restated on CPU: oracle UNet (diffusers restatement) + oracle LyCORIS (kron forward patch) +
    the reference loss arithmetic + torch.optim.AdamW, under torch.autocast('cpu', bf16) to mirror `bf16-mixed`."""

    def __init__(self):
        import torch

        from oracle import diffusers_shim, loss_oracle, lycoris_oracle, unet_oracle

        self.torch = torch
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        with torch.device("meta"):
            unet = unet_oracle.UNet2DConditionModel()
        unet = unet.to_empty(device="cpu")
        g = torch.Generator().manual_seed(0)
        for p in unet.parameters():
            if p.dim() > 1:
                p.data.uniform_(-0.02, 0.02, generator=g)
            else:
                p.data.zero_()
        for m in unet.modules():
            if isinstance(m, (torch.nn.GroupNorm, torch.nn.LayerNorm)):
                m.weight.data.fill_(1.0)
        lycoris_oracle.LycorisNetwork.apply_preset(LYCORIS_CONFIG["preset"])
        self.net = lycoris_oracle.create_lycoris(unet, **LYCORIS_CONFIG["config"])
        self.net.apply_to()
        unet.requires_grad_(False)
        self.unet = unet
        sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("x", prediction_type="v_prediction")
        self.tab = loss_oracle.scheduler_tables(sch)
        self.loss_oracle = loss_oracle
        self.opt = torch.optim.AdamW(self.net.parameters(), lr=1e-6, weight_decay=0.01, betas=(0.9, 0.999))
        self.cfg = unet_oracle.SDXL_UNET_CONFIG

    def step(self, batch: int, latent: int) -> float:
        torch = self.torch
        g = torch.Generator().manual_seed(1215)
        x0 = torch.randn((batch, 4, latent, latent), generator=g)
        eps = torch.randn(x0.shape, generator=g)
        t = torch.randint(0, 1000, (batch,), generator=g)
        ctx = torch.randn((batch, 77, 2048), generator=g)
        ac = dict(text_embeds=torch.randn((batch, 1280), generator=g),
                  time_ids=torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * batch))
        t0 = time.time()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss, _ = self.loss_oracle.diffusion_loss(x0, eps, t, self.unet, self.tab, target_type="v_prediction",
                                                      prediction_type="v_prediction", use_snr_weight=True,
                                                      encoder_hidden_states=ctx, added_cond_kwargs=ac)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(self.net.parameters()), 1.0)
        self.opt.step()
        self.opt.zero_grad()
        return time.time() - t0

    def scaled_samples_per_s(self, seconds: float, batch: int, latent: int, target_latent: int) -> float:
        from uwudiff_b200.flops import unet_forward_flops

        f_s = unet_forward_flops(self.cfg, latent, latent)["total"]
        f_t = unet_forward_flops(self.cfg, target_latent, target_latent)["total"]
        return (batch / seconds) * (f_s / f_t)


def cpu_sample_choice(oracle: CpuOracleStep, budget_s: float, n_steps: int):
    """Warm up at 32x32 latents, then pick the largest latent size whose n_steps fit the budget."""
    t32 = oracle.step(1, 32)
    t32 = oracle.step(1, 32)
    from uwudiff_b200.flops import unet_forward_flops

    r = unet_forward_flops(oracle.cfg, 64, 64)["total"] / unet_forward_flops(oracle.cfg, 32, 32)["total"]
    latent = 64 if t32 * r * n_steps <= budget_s else 32
    return latent, t32


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle = CpuOracleStep()
    latent, _ = cpu_sample_choice(oracle, 150.0, args.steps + args.warmup)
    for _ in range(args.warmup):
        oracle.step(1, latent)
    ts = [oracle.step(1, latent) for _ in range(args.steps)]
    sec = sum(ts) / len(ts)
    v = oracle.scaled_samples_per_s(sec, 1, latent, args.latent)
    sample = (f"1 image at {latent}x{latent} latents per step (full SDXL UNet + LyCORIS, fwd+bwd+clip+AdamW, autocast bf16), "
              f"scaled by algorithmic forward FLOPs {latent}^2 -> {args.latent}^2")
    line = {
        "impl": "reference", "metric": "train_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": oracle.cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "SDXL UNet (2.57 B params, frozen) + LyCORIS LoKr/LoRA/norm adapters (52.4 M trainable), "
                        f"4x{args.latent}x{args.latent} latents, batch {args.batch} per GPU, v-pred min-SNR, AdamW + clip 1.0 "
                        "[BASELINE.json configs[2]; at 8 GPUs = configs[4] global batch 128]",
            "global_batch": args.batch * world, "per_gpu_batch": args.batch, "latent": [4, args.latent, args.latent],
            "parallelism": f"dp{world}", "l2": "working set >> L2 (~100 GB of activations per step), no flush needed",
            "gradient_checkpointing": False}


# ======================================================================================================
# GPU arm
# ======================================================================================================
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from uwudiff_b200 import config as ucfg
    from uwudiff_b200 import ops
    from uwudiff_b200.flops import unet_forward_flops
    from uwudiff_b200.unet import SDXL_UNET_CONFIG

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the uwudiff_b200 kernels have no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1215 + rank)

    conf = trainer_config(args.latent, args.batch)
    trainer = ucfg.instantiate_any(conf["trainer"])
    trainer.setup_fit(gradient_clip_val=conf["lightning_config"]["gradient_clip_val"], seed=conf["seed"])
    B, S = args.batch, args.latent
    host_x = torch.randn((B, 4, S, S)).pin_memory()
    host_ids = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B).pin_memory()
    dev_batch = (host_x.to(dev), ["DUMMY TEST"] * B, [], {"time_ids": host_ids.to(dev)}, {})

    def host_batch():
        return (host_x, ["DUMMY TEST"] * B, [], {"time_ids": host_ids}, {})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, batch_fn, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        w0 = time.time()
        e0.record()
        last = None
        for i in range(n):
            out = trainer.fit_step(batch_fn(), i)
            if read_loss:
                last = out["loss"].item()
            else:
                last = out["loss"]
        e1.record()
        barrier()
        w1 = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / n, ops.launch_count() - l0, (w0, w1), last

    def note(msg):
        if rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    note(f"trainer ready (world {world})")
    for i in range(max(args.warmup, 3)):
        trainer.fit_step(dev_batch, i)
    barrier()
    note("warm-up done")
    clocks = ClockSampler(local) if rank == 0 else None
    ms_dev, launches, (w0, _), loss_dev = timed(args.steps, lambda: dev_batch, False)
    ms_e2e, _, (_, w1), loss_host = timed(args.steps, host_batch, True)
    clk = clocks.stop(w0, w1) if clocks is not None else None
    note(f"timed: {ms_dev:.1f} ms/step resident, {ms_e2e:.1f} ms/step end-to-end")

    # ---- dominant kernel: every tcgen05 GEMM / conv launch of one more step, CUDA events on the launching stream ----
    # (every rank runs this extra step: it contains the gradient all-reduce; only rank 0 keeps the timings)
    roof = None
    if True:
        real_gemm = ops.gemm
        recs = []

        def timed_gemm(a, b, M, N, K, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = real_gemm(a, b, M, N, K, **kw)
            e.record()
            recs.append((s, e, 2.0 * M * N * K * max(1, kw.get("k_segs", 0))))
            return r

        ops.gemm = timed_gemm
        try:
            trainer.fit_step(dev_batch, 0)
            torch.cuda.synchronize()
        finally:
            ops.gemm = real_gemm
    if rank == 0:
        t_ms = sum(s.elapsed_time(e) for s, e, _ in recs)
        fl = sum(f for _, _, f in recs)
        pk, pk_src = peaks()
        achieved = fl / (t_ms * 1e-3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        fwd = unet_forward_flops(SDXL_UNET_CONFIG, S, S)["total"] * B
        traffic, traffic_note = None, None
        try:  # DRAM bytes of the representative launch (GEGLU projection, M16384 N10240 K1280) from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")) as f:
                g0 = json.load(f)["gemm"][0]
            traffic = (g0["dram_read_mb"] + g0["dram_write_mb"]) * 1e6
            traffic_note = ("ncu --set full, one launch of M16384 N10240 K1280 (the largest Linear class): "
                            f"{g0['dram_read_mb']:.0f} MB read + {g0['dram_write_mb']:.0f} MB written vs 403 MB algorithmic "
                            f"(A 42 + B 26 + out 335), tensor pipe {g0['tensor_pct']:.1f}% active, {g0['time_us']:.0f} us")
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (Linear + implicit-GEMM conv, fwd/dgrad/adapter-wgrad)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note,
                "peak_source": pk_src + " bf16_tflops_sustained", "launches_per_step": len(recs),
                "kernel_ms_per_step": t_ms, "kernel_share_of_step": t_ms / ms_dev,
                "step_algorithmic_tflop": 2 * fwd / 1e12, "step_algorithmic_tflops": 2 * fwd / 1e12 / (ms_dev * 1e-3),
                "step_frac_of_peak": 2 * fwd / 1e12 / (ms_dev * 1e-3) / peak}
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            oracle = CpuOracleStep()
            latent, _ = cpu_sample_choice(oracle, 25.0, 1)
            sec = oracle.step(1, latent)
            cpu = {"value": oracle.scaled_samples_per_s(sec, 1, latent, S), "unit": "samples/s", "cores": oracle.cores,
                   "kind": "port",
                   "sample": f"1 image at {latent}x{latent} latents (full SDXL UNet + LyCORIS oracle, fwd+bwd+clip+AdamW, autocast "
                             f"bf16; {sec:.1f} s), scaled by algorithmic forward FLOPs {latent}^2 -> {S}^2"}
        except Exception as e:  # the baseline is informational; never lose the GPU line to it
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}

    if rank == 0:
        h2d = host_x.numel() * 4 + host_ids.numel() * 4
        line = {
            "metric": "train_samples_per_s", "value": B * world / (ms_dev * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world), "per_gpu": B / (ms_dev * 1e-3), "clocks": clk,
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "loss": float(loss_host) if loss_host is not None else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="uwu", choices=["uwu", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU")
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == "__main__":
    main()
