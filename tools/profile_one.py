"""Run one op a few times so that `ncu -k regex:<kernel> --launch-skip N -c 1` can capture it.

    python tools/profile_one.py attn   | gemm_lin | gemm_res | ln | gn | geglu
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uwudiff_b200 import ops

dev = "cuda"


def mk(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


def main(what):
    if what == "attn":
        B, heads, L = 16, 10, 4096
        C = heads * 64
        qkv = mk(B * L, 3 * C)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        do = mk(B * L, C)
        for _ in range(2):
            o, lse = ops.attn_fwd(q, k, v, B, heads, L, L)
            ops.attn_bwd(q, k, v, o, do, lse, B, heads, L, L)
    elif what in ("gemm_lin", "gemm_res"):
        M, N, K = 16384, 1280, 1280
        a, b = mk(M, K), mk(N, K)
        res = mk(M, N) if what == "gemm_res" else None
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, b, M, N, K, out=out, residual=res)
        M, N, K = 16384, 10240, 1280
        a, b = mk(M, K), mk(N, K)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(N, device=dev)
        for _ in range(3):
            ops.gemm(a, b, M, N, K, out=out, bias=bias)
    elif what == "conv320":
        # the weakest GEMM class against cuDNN (profiles/r02_gemm_perf12*.log): 3x3 conv 320 -> 320 at 128 x 128, batch 16
        x = mk(16, 128, 128, 320)
        wp = mk(320, 9 * 320)
        out = torch.empty(16 * 128 * 128, 320, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(320, device=dev)
        for _ in range(3):
            ops.conv3x3_nhwc(x, wp, out=out, bias=bias)
        x = mk(16, 64, 64, 640)
        wp = mk(640, 9 * 640)
        out = torch.empty(16 * 64 * 64, 640, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.conv3x3_nhwc(x, wp, out=out)
    elif what == "bw":
        # one launch each of the bandwidth-bound kernels at step shapes
        from uwudiff_b200.loss import DiffusionLoss
        from uwudiff_b200.scheduler import EulerDiscreteScheduler

        M, C = 16384, 1280
        x, dy, dres = mk(M, C), mk(M, C), mk(M, C)
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        for _ in range(2):
            y, stats = ops.layernorm_fwd(x, gamma, beta)
            ops.layernorm_bwd(x, dy, gamma, stats, dres=dres, dgamma=dg, dbeta=db)
        N, HW, Cg = 16, 128 * 128, 320
        xg, dyg, drg = mk(N * HW, Cg), mk(N * HW, Cg), mk(N * HW, Cg)
        gg, bg = torch.ones(Cg, device=dev), torch.zeros(Cg, device=dev)
        for _ in range(2):
            yg, sg = ops.groupnorm_fwd(xg, N, HW, Cg, 32, 1e-5, gg, bg, True)
            ops.groupnorm_bwd(xg, dyg, N, HW, Cg, 32, gg, bg, sg, True, dres=drg)
        xp, dout = mk(M, 10240), mk(M, 5120)
        for _ in range(2):
            ops.geglu_fwd(xp)
            ops.geglu_bwd(xp, dout)
        sch = EulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler",
                                                     prediction_type="v_prediction")
        L = DiffusionLoss(sch, use_snr_weight=True)
        x0 = torch.randn(128, 4, 128, 128, device=dev)
        tab = L._device_tables(x0.device)
        for _ in range(2):
            x_t, target, _, t, sig, w, temb = ops.noise_fwd(x0, tab, target_type="v_prediction", pred_type="v_prediction",
                                                            use_snr_weight=True, use_debiased=False, gamma=5.0, seed=1, offset=0,
                                                            temb_dim=320, want_eps=False)
            ops.wmse_fwd(x_t, target, w)
            ops.wmse_bwd(x_t, target, w)
    elif what == "dit":
        # DiT-XL/2 shapes: token-reduction weight gradient under the three split-K schedules, d = 72 attention, adaLN glue
        from uwudiff_b200._lib import A_COL, B_KN

        Mtok, Co, Ci = 65536, 1152, 4608
        dy, x = mk(Mtok, Co), mk(Mtok, Ci)
        G = torch.zeros(Co, Ci, device=dev)
        for sk in (0, 1, -1):
            ops.gemm(dy, x, Co, Ci, Mtok, a_layout=A_COL, lda=Co, b_layout=B_KN, ldb=Ci, out=G, stream_k=sk, accumulate=True)
        B, heads, L, d = 256, 16, 256, 72
        C = heads * d
        qkv, do = mk(B * L, 3 * C), mk(B * L, C)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        o, lse = ops.attn_fwd(q, k, v, B, heads, L, L, head_dim=d)
        ops.attn_bwd(q, k, v, o, do, lse, B, heads, L, L, head_dim=d)
        xt, dyt = mk(B * L, C), mk(B * L, C)
        mod = torch.randn(B, 6 * C, device=dev)
        dmod = torch.zeros(B, 6 * C, device=dev, dtype=torch.bfloat16)
        y, st = ops.adaln_fwd(xt, mod, 0, C, L)
        ops.adaln_bwd(xt, dyt, mod, C, st, L, dmod, 0, C, dres=dyt)
        ops.gate_residual_fwd(xt, y, mod, 2 * C, L)
        ops.gate_residual_bwd(dyt, y, mod, 2 * C, L, dmod, 2 * C)
        ops.elementwise(mk(B * L, 4 * C), None, ops.EW_GELU_TANH)
    elif what == "lokr_fused":
        M, ol, im = 16384, 20, 20
        dy, x = mk(M, ol * 64), mk(M, im * 64)
        w1, w2 = torch.randn(ol, im, device=dev), torch.randn(64, 64, device=dev)
        dw1, dw2 = torch.zeros_like(w1), torch.zeros_like(w2)
        for _ in range(3):
            ops.lokr_fused_grad(dy, x, M, w1, w2, dw1, dw2)
    elif what == "cross":
        # cross-attention forward at the step shape on both paths (UWU_ATTN_SHORT is read per call), column sums, conv pack
        B, heads, L, Lk = 16, 20, 1024, 77
        C = heads * 64
        q, k, v, do = mk(B * L, C), mk(B * Lk, C), mk(B * Lk, C), mk(B * L, C)
        for flag in ("1", "0"):
            os.environ["UWU_ATTN_SHORT"] = flag
            o, lse = ops.attn_fwd(q, k, v, B, heads, L, Lk)
        ops.attn_bwd(q, k, v, o, do, lse, B, heads, L, Lk)
        x = mk(65536, 4608)
        ops.colsum(x)
        W = torch.randn(1280, 1280, 3, 3, device=dev)
        fwd = torch.empty(1280, 9 * 1280, device=dev, dtype=torch.bfloat16)
        dg = torch.empty(1280, 9 * 1280, device=dev, dtype=torch.bfloat16)
        ops.conv_pack(W, 1280, 1280, 1280, fwd, dg)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main(sys.argv[1])
