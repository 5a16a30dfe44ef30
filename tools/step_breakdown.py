"""Per-op CUDA-event breakdown of one SDXL + LyCORIS training step (in situ: warm caches, real clocks / power state).

    python tools/step_breakdown.py [--batch 16] [--latent 128] [--top 40]

Every `uwudiff_b200.ops` entry point is wrapped with a pair of CUDA events on the launching stream; one step after warm-up
is recorded and aggregated by op (and by shape for the GEMM / conv / attention calls).
"""
import argparse
import collections
import os

os.environ.setdefault("UWU_SYNTHETIC_CONDITIONING", "1")  # synthetic text-encoder outputs (no weights offline)
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import bench
from uwudiff_b200 import config as ucfg
from uwudiff_b200 import ops

WRAP = ["gemm", "noise_fwd", "sincos_embed", "wmse_fwd", "wmse_bwd", "attn_fwd", "attn_bwd", "groupnorm_fwd", "groupnorm_bwd",
        "layernorm_fwd", "layernorm_bwd", "geglu_fwd", "geglu_bwd", "elementwise", "nchw_to_nhwc", "nhwc_to_nchw", "upsample2x",
        "phase_split2", "colsum", "fold_lokr", "fold_lora", "axpy_f32", "lokr_grad", "lora_grad", "copy2d", "lokr_z", "lokr_dw1", "lokr_fused_grad",
        "im2col3x3", "conv_wgrad_unpack", "colsum_groups", "pred_convert", "adaln_fwd", "adaln_bwd", "gate_residual_fwd", "gate_residual_bwd", "patchify",
        "unpatchify", "embed_gather", "embed_scatter_add", "conv_pack", "fold_loha", "loha_grad", "timestep_hist"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--config", default="c3", choices=["c3", "c1", "c2", "c4", "latent"], help="c3 = bench.py headline workload")
    args = ap.parse_args()
    dev = torch.device("cuda")
    shape = (4, args.latent, args.latent)
    if args.config == "c3":
        conf = bench.trainer_config(args.latent, args.batch)
    else:
        import bench_configs

        args.batch = {"c1": 4, "c2": 32, "c4": 256, "latent": 16}[args.config]
        conf, _, shape, _ = bench_configs.make_conf(args.config, args.batch)
    trainer = ucfg.instantiate_any(conf["trainer"])
    trainer.setup_fit(gradient_clip_val=1.0, seed=1215)
    B, S = args.batch, args.latent
    batch = (torch.randn((B, *shape), device=dev), ["DUMMY TEST"] * B, [],
             ({"class_labels": torch.randint(0, 1000, (B,), device=dev)} if args.config == "c4" else
              {"time_ids": torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B, device=dev)}), {})
    for i in range(3):
        trainer.fit_step(batch, i)
    torch.cuda.synchronize()
    recs = []
    real = {n: getattr(ops, n) for n in WRAP if hasattr(ops, n)}

    def wrap(name, fn):
        def f(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            key = name
            if name == "gemm":
                M, N, K = a[2], a[3], a[4]
                kind = "conv" if k.get("conv") is not None else ("wgrad" if k.get("a_layout", 0) == 1 else "lin")
                key = f"gemm[{kind}] M{M} N{N} K{K}"
                fl = 2.0 * M * N * K
            elif name in ("attn_fwd", "attn_bwd"):
                Bq, h, Lq, Lk = a[3:7] if name == "attn_fwd" else a[6:10]
                hd = k.get("head_dim", 64)
                key = f"{name} B{Bq} h{h} Lq{Lq} Lk{Lk} d{hd}"
                fl = 4.0 * Bq * h * Lq * Lk * hd * (1.0 if name == "attn_fwd" else 2.5)
            else:
                t0 = next((t for t in a if torch.is_tensor(t)), None)
                if t0 is not None:
                    key = f"{name} {tuple(t0.shape)}"
                fl = 0.0
            recs.append((key, name, s, e, fl))
            return r
        return f

    for n, fn in real.items():
        setattr(ops, n, wrap(n, fn))
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    trainer.fit_step(batch, 3)
    s1.record()
    torch.cuda.synchronize()
    for n, fn in real.items():
        setattr(ops, n, fn)
    total = s0.elapsed_time(s1)
    by_key = collections.defaultdict(lambda: [0, 0.0, 0.0])
    by_op = collections.defaultdict(lambda: [0, 0.0])
    for key, name, s, e, fl in recs:
        ms = s.elapsed_time(e)
        by_key[key][0] += 1
        by_key[key][1] += ms
        by_key[key][2] += fl
        by_op[name][0] += 1
        by_op[name][1] += ms
    acc = sum(v[1] for v in by_op.values())
    print(f"step {total:.1f} ms (instrumented); ops account for {acc:.1f} ms over {len(recs)} calls")
    print("--- by op ---")
    for k, v in sorted(by_op.items(), key=lambda x: -x[1][1]):
        print(f"{v[1]:9.2f} ms {v[0]:6d}  {100*v[1]/total:5.1f}%  {k}")
    print("--- by op+shape ---")
    for k, v in sorted(by_key.items(), key=lambda x: -x[1][1])[: args.top]:
        tf = f" {v[2]/v[1]/1e9:7.1f} TFLOP/s" if v[2] else ""
        print(f"{v[1]:9.2f} ms {v[0]:5d} x {1e3*v[1]/v[0]:8.1f} us{tf}  {k}")


if __name__ == "__main__":
    main()
