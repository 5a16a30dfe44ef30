#!/usr/bin/env python
"""Same-box library comparator (SURVEY.md §8d "secondary comparator"): the ORACLE UNet + LyCORIS restatement of the reference
path, run with stock PyTorch on `cuda` under `torch.autocast(bf16)` — i.e. cuDNN convolutions, cuBLASLt GEMMs, SDPA flash
attention, ATen norms, autograd, torch.optim.AdamW — on the same workload and timed the same way as bench.py (CUDA events,
W warm-up + K timed steps).  This is what the reference stack would run on this GPU; none of this repo's kernels are on it.

    python tools/bench_eager_cuda.py [--batch 16] [--latent 128] [--steps 5] [--warmup 3]

Prints one JSON line; the builder commits it under profiles/ and DESIGN.md §5 quotes it next to bench.py's number.
Test infrastructure / measurement only (imports oracle/): never part of the product path.
"""
from __future__ import annotations

import argparse
import json
import os

os.environ.setdefault("UWU_SYNTHETIC_CONDITIONING", "1")  # synthetic text-encoder outputs (no weights offline)
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--checkpointing", action="store_true", help="recompute each top-level block in backward (reference default)")
    args = ap.parse_args()
    import torch

    import bench

    assert torch.cuda.is_available()
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    st = bench.CpuOracleStep(device="cuda")
    if args.checkpointing:
        from torch.utils.checkpoint import checkpoint

        for blk in list(st.unet.down_blocks) + [st.unet.mid_block] + list(st.unet.up_blocks):
            f = blk.forward
            blk.forward = (lambda f: lambda *a, **k: checkpoint(f, *a, use_reentrant=False, **k))(f)
    B = args.batch
    result = None
    while B >= 1:
        try:
            for _ in range(args.warmup):
                st.step(B, args.latent)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                st.step(B, args.latent)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            result = dict(batch=B, ms_per_step=ms, samples_per_s=B / (ms * 1e-3),
                          peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
            break
        except torch.OutOfMemoryError:
            st.opt.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
            print(f"[eager] batch {B} does not fit without this repo's memory plan; halving", file=sys.stderr)
            B //= 2
    from uwudiff_b200.flops import unet_forward_flops

    fwd = unet_forward_flops(st.cfg, args.latent, args.latent)["total"]
    line = {"impl": "torch eager cuda (oracle UNet + LyCORIS, autocast bf16: cuDNN / cuBLASLt / SDPA)", "metric": "train_samples_per_s",
            "requested_batch": args.batch, "latent": args.latent, "steps": args.steps, "warmup": args.warmup,
            "gradient_checkpointing": args.checkpointing, "torch": torch.__version__,
            "gpu": torch.cuda.get_device_name(0)}
    if result:
        line.update(result)
        line["step_algorithmic_tflops"] = 2 * fwd * result["batch"] / 1e12 / (result["ms_per_step"] * 1e-3)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
