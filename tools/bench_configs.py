"""Throughput of the other named configurations of BASELINE.json (the headline, configs[2], is `bench.py`).

    python tools/bench_configs.py c1|c2|latent [--steps 5] [--warmup 3] [--batch B]

  c1      configs/demo_training.yaml-style pixel UNet, batch 4 at 3x32x32, eps-pred MSE, full fine-tune
  c2      SD-1.5-size latent UNet (860 M params), 4x64x64 latents, batch 32, bf16 compute, full fine-tune
  c4      DiT-XL/2 class-conditional (675 M params), 4x32x32 latents, batch 256 on one GPU (32 per GPU at 8), full fine-tune
  latent  the reference's configs/demo_training_latent.yaml as shipped: SDXL UNet, 4x32x32 latents, batch 16, full fine-tune

Same timing rules as bench.py: >= 3 warm-up steps, CUDA events around K whole `DMTrainer.fit_step` calls (forward, backward,
clip, fused AdamW over every parameter), resident inputs for `value`, pinned host inputs + loss read-back for `e2e`.
Prints one JSON line.  `tensor_tflops` = 3 x algorithmic forward FLOPs (fwd + dgrad + wgrad) / time.
"""
import argparse
import json
import os

os.environ.setdefault("UWU_SYNTHETIC_CONDITIONING", "1")  # synthetic text-encoder outputs (no weights offline)
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from uwudiff_b200 import config as ucfg
from uwudiff_b200 import ops
from uwudiff_b200.flops import unet_forward_flops

C1_UNET = dict(in_channels=3, out_channels=3, sample_size=32, block_out_channels=(64, 128, 256),
               down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
               up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"), layers_per_block=1,
               transformer_layers_per_block=1, attention_head_dim=(1, 2, 4), cross_attention_dim=128,
               projection_class_embeddings_input_dim=64 + 6 * 32, addition_time_embed_dim=32)


def make_conf(kind: str, batch: int):
    conf = bench.trainer_config(32, batch)
    tr = conf["trainer"]
    tr["lycoris_config"] = None
    sched = tr["loss_config"]["scheduler"]
    if kind == "c1":
        tr["model_config"]["unet"] = {"_target_": "duwu.modules.unet_patch.UNet2DFromScratch.from_config", "config": C1_UNET}
        tr["model_config"]["te"]["hidden_dim"] = 128
        tr["model_config"]["te"]["pooled_dim"] = 64
        sched["prediction_type"] = "epsilon"
        tr["loss_config"]["use_snr_weight"] = False
        return conf, C1_UNET, (3, 32, 32), "pixel UNet (block_out 64/128/256), eps-pred MSE [configs[0]]"
    if kind == "c2":
        tr["model_config"]["unet"]["config"] = "runwayml/stable-diffusion-v1-5"
        tr["model_config"]["te"]["hidden_dim"] = 768
        sched["prediction_type"] = "epsilon"
        tr["loss_config"]["use_snr_weight"] = False
        from uwudiff_b200.unet import SD15_UNET_CONFIG
        return conf, SD15_UNET_CONFIG, (4, 64, 64), "SD-1.5-size latent UNet (860 M params) full fine-tune [configs[1]]"
    if kind == "c4":
        from uwudiff_b200.dit import DIT_XL_2_CONFIG
        tr["model_config"]["unet"] = {"_target_": "uwudiff_b200.dit.DiT.from_config", "config": "DiT-XL/2"}
        tr["model_config"]["te"] = None
        sched["prediction_type"] = "epsilon"
        tr["loss_config"]["use_snr_weight"] = False
        return conf, DIT_XL_2_CONFIG, (4, 32, 32), "DiT-XL/2 class-conditional, adaLN-Zero, full fine-tune [configs[3]]"
    from uwudiff_b200.unet import SDXL_UNET_CONFIG
    sched["prediction_type"] = "epsilon"
    tr["loss_config"]["use_snr_weight"] = False
    return conf, SDXL_UNET_CONFIG, (4, 32, 32), "SDXL UNet full fine-tune (reference demo_training_latent.yaml as shipped)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["c1", "c2", "c4", "latent"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    B = args.batch or {"c1": 4, "c2": 32, "c4": 256, "latent": 16}[args.kind]
    dev = torch.device("cuda")
    conf, ucfg_dict, shape, label = make_conf(args.kind, B)
    trainer = ucfg.instantiate_any(conf["trainer"])
    trainer.setup_fit(gradient_clip_val=1.0, seed=1215)
    host_x = torch.randn((B, *shape)).pin_memory()
    host_ids = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B).pin_memory()
    host_cond = {"time_ids": host_ids}
    if args.kind == "c4":
        host_cond = {"class_labels": torch.randint(0, 1000, (B,)).pin_memory()}
    dev_batch = (host_x.to(dev), ["DUMMY TEST"] * B, [], {k: v.to(dev) for k, v in host_cond.items()}, {})

    def timed(batch_fn, read):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        e0.record()
        for i in range(args.steps):
            out = trainer.fit_step(batch_fn(), i)
            if read:
                out["loss"].item()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, (ops.launch_count() - l0) // args.steps, float(out["loss"])

    for i in range(max(3, args.warmup)):
        trainer.fit_step(dev_batch, i)
    ms, launches, loss = timed(lambda: dev_batch, False)
    ms_e2e, _, _ = timed(lambda: (host_x, ["DUMMY TEST"] * B, [], dict(host_cond), {}), True)
    if args.kind == "c4":
        from uwudiff_b200.dit import dit_forward_flops
        fwd = dit_forward_flops(ucfg_dict)["total"]
    else:
        fwd = unet_forward_flops(ucfg_dict, shape[1], shape[2])["total"]
    peak, _ = bench.peaks()
    sus = peak.get("bf16_tflops_sustained") or peak.get("bf16_sustained_tflops")
    tf = 3.0 * fwd * B / (ms * 1e-3) / 1e12
    n_params = sum(p.numel() for p in trainer.unet.parameters())
    print(json.dumps({
        "metric": "train_samples_per_s", "value": B / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "dtype": "bf16", "data": "synthetic", "loss": loss,
        "config": {"workload": label, "batch": B, "input": list(shape), "params": n_params, "trainable": "all (full fine-tune)"},
        "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": host_x.numel() * 4 + sum(v.numel() * v.element_size() for v in host_cond.values()),
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "tensor_tflops": {"achieved": tf, "algorithmic_tflop_per_step": 3.0 * fwd * B / 1e12,
                          "peak_sustained": sus, "frac": tf / sus if sus else None},
    }))


if __name__ == "__main__":
    main()
