"""GPU diagnostic for the attention / norm / glue kernels (run under gpurun).

    python tools/diag_kernels.py <case> [...]

Each case prints max-abs / relative error against a torch fp32 evaluation of the same bf16 inputs.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from uwudiff_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def report(name, got, ref):
    got = got.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    bad = (~torch.isfinite(got)).sum().item()
    print(f"{name}: max_abs_err={err:.4e} ref_max={scale:.3e} rel={err/scale:.3e} nonfinite={bad}", flush=True)
    return err / scale


def mk(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).to(torch.bfloat16)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def attn_ref(q, k, v, B, heads, Lq, Lk, dout=None):
    qf = q.float().reshape(B, Lq, heads, 64).transpose(1, 2).detach().requires_grad_(True)
    kf = k.float().reshape(B, Lk, heads, 64).transpose(1, 2).detach().requires_grad_(True)
    vf = v.float().reshape(B, Lk, heads, 64).transpose(1, 2).detach().requires_grad_(True)
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    p = torch.softmax(s, dim=-1)
    o = p @ vf
    lse = torch.logsumexp(s, dim=-1)
    o2 = o.transpose(1, 2).reshape(B * Lq, heads * 64)
    if dout is None:
        return o2, lse
    o2.backward(dout.float())
    back = lambda t, L: t.grad.transpose(1, 2).reshape(B * L, heads * 64)
    return o2, lse, back(qf, Lq), back(kf, Lk), back(vf, Lk)


def case_attn_fwd():
    for (B, heads, Lq, Lk) in [(1, 1, 128, 128), (1, 1, 256, 128), (1, 1, 128, 256), (2, 3, 256, 256), (2, 2, 1024, 1024),
                               (2, 2, 64, 64), (2, 5, 256, 77), (1, 2, 320, 200)]:
        C = heads * 64
        q, k, v = mk(B * Lq, C), mk(B * Lk, C), mk(B * Lk, C)
        o, lse = ops.attn_fwd(q, k, v, B, heads, Lq, Lk)
        torch.cuda.synchronize()
        ro, rlse = attn_ref(q, k, v, B, heads, Lq, Lk)
        report(f"attn_fwd B{B} h{heads} Lq{Lq} Lk{Lk} O", o, ro)
        Lp = (Lq + 127) // 128 * 128
        report(f"attn_fwd B{B} h{heads} Lq{Lq} Lk{Lk} lse", lse.view(B, heads, Lp)[:, :, :Lq], rlse)
    # fused-QKV strided views
    B, heads, L = 2, 2, 256
    C = heads * 64
    qkv = mk(B * L, 3 * C)
    o, lse = ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads, L, L)
    ro, _ = attn_ref(qkv[:, :C].contiguous(), qkv[:, C:2 * C].contiguous(), qkv[:, 2 * C:].contiguous(), B, heads, L, L)
    report("attn_fwd strided qkv", o, ro)
    print("DONE attn_fwd")


def case_attn_bwd():
    for (B, heads, Lq, Lk) in [(1, 1, 128, 128), (1, 1, 256, 128), (1, 1, 128, 256), (2, 3, 256, 256), (2, 2, 1024, 1024),
                               (2, 2, 64, 64), (2, 5, 256, 77), (1, 2, 320, 200)]:
        C = heads * 64
        q, k, v, do = mk(B * Lq, C), mk(B * Lk, C), mk(B * Lk, C), mk(B * Lq, C)
        o, lse = ops.attn_fwd(q, k, v, B, heads, Lq, Lk)
        dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, B, heads, Lq, Lk)
        torch.cuda.synchronize()
        _, _, rdq, rdk, rdv = attn_ref(q, k, v, B, heads, Lq, Lk, do)
        tag = f"attn_bwd B{B} h{heads} Lq{Lq} Lk{Lk}"
        report(tag + " dq", dq, rdq)
        report(tag + " dk", dk, rdk)
        report(tag + " dv", dv, rdv)
    print("DONE attn_bwd")


def case_attn_perf():
    for (B, heads, L, Lk) in [(16, 10, 4096, 4096), (16, 20, 1024, 1024), (16, 10, 4096, 77), (16, 20, 1024, 77)]:
        C = heads * 64
        qkv = mk(B * L, 3 * C) if Lk == L else None
        if qkv is not None:
            q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        else:
            q, k, v = mk(B * L, C), mk(B * Lk, C), mk(B * Lk, C)
        do = mk(B * L, C)
        o, lse = ops.attn_fwd(q, k, v, B, heads, L, Lk)
        us = timeit(lambda: ops.attn_fwd(q, k, v, B, heads, L, Lk))
        fl = 4.0 * B * heads * L * Lk * 64
        print(f"perf attn_fwd B{B} h{heads} L{L} Lk{Lk}: {us:.1f} us  {fl/us/1e6:.1f} TFLOP/s", flush=True)
        us = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, B, heads, L, Lk))
        print(f"perf attn_bwd B{B} h{heads} L{L} Lk{Lk}: {us:.1f} us  {2.5*fl/us/1e6:.1f} TFLOP/s (2.5x fwd flops)", flush=True)
        qf = q.reshape(B, L, heads, 64).transpose(1, 2)
        kf = k.reshape(B, Lk, heads, 64).transpose(1, 2)
        vf = v.reshape(B, Lk, heads, 64).transpose(1, 2)
        us = timeit(lambda: F.scaled_dot_product_attention(qf, kf, vf))
        print(f"perf torch sdpa fwd: {us:.1f} us  {fl/us/1e6:.1f} TFLOP/s", flush=True)
    print("DONE attn_perf")


def case_norm():
    # GroupNorm (+SiLU) fwd/bwd
    for (N, H, W, C, G, silu, eps) in [(2, 16, 16, 64, 32, True, 1e-5), (3, 8, 8, 320, 32, True, 1e-5),
                                       (2, 32, 32, 640, 32, False, 1e-6), (2, 16, 16, 1920, 32, True, 1e-5),
                                       (1, 64, 64, 960, 32, True, 1e-5)]:
        HW = H * W
        x = mk(N * HW, C)
        dy = mk(N * HW, C)
        dres = mk(N * HW, C)
        gamma = torch.randn(C, device=dev) * 0.5 + 1
        beta = torch.randn(C, device=dev) * 0.5
        y, stats = ops.groupnorm_fwd(x, N, HW, C, G, eps, gamma, beta, silu)
        xr = x.float().reshape(N, HW, C).permute(0, 2, 1).detach().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        yr = F.group_norm(xr, G, gr, br, eps)
        if silu:
            yr = F.silu(yr)
        yr2 = yr.permute(0, 2, 1).reshape(N * HW, C)
        tag = f"gn N{N} HW{HW} C{C} silu{int(silu)}"
        report(tag + " y", y, yr2.detach())
        yr2.backward(dy.float())
        dgamma = torch.zeros(C, device=dev)
        dbeta = torch.zeros(C, device=dev)
        dx = ops.groupnorm_bwd(x, dy, N, HW, C, G, gamma, beta, stats, silu, dres=dres, dgamma=dgamma, dbeta=dbeta)
        torch.cuda.synchronize()
        report(tag + " dx", dx, xr.grad.permute(0, 2, 1).reshape(N * HW, C) + dres.float())
        report(tag + " dgamma", dgamma, gr.grad)
        report(tag + " dbeta", dbeta, br.grad)
    # LayerNorm fwd/bwd
    for (M, C) in [(256, 64), (1000, 640), (4096, 1280), (77, 128)]:
        x, dy, dres = mk(M, C), mk(M, C), mk(M, C)
        gamma = torch.randn(C, device=dev) * 0.5 + 1
        beta = torch.randn(C, device=dev) * 0.5
        y, stats = ops.layernorm_fwd(x, gamma, beta, 1e-5)
        xr = x.float().detach().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        yr = F.layer_norm(xr, (C,), gr, br, 1e-5)
        report(f"ln M{M} C{C} y", y, yr.detach())
        yr.backward(dy.float())
        dgamma = torch.zeros(C, device=dev)
        dbeta = torch.zeros(C, device=dev)
        dx = ops.layernorm_bwd(x, dy, gamma, stats, dres=dres, dgamma=dgamma, dbeta=dbeta, accumulate=True)
        torch.cuda.synchronize()
        report(f"ln M{M} C{C} dx", dx, xr.grad + dres.float())
        report(f"ln M{M} C{C} dgamma", dgamma, gr.grad)
        report(f"ln M{M} C{C} dbeta", dbeta, br.grad)
    print("DONE norm")


def case_glue():
    M, Fh = 512, 256
    x, dout = mk(M, 2 * Fh), mk(M, Fh)
    y = ops.geglu_fwd(x)
    xr = x.float().detach().requires_grad_(True)
    h, g = xr.chunk(2, dim=-1)
    yr = h * F.gelu(g)
    report("geglu fwd", y, yr.detach())
    yr.backward(dout.float())
    report("geglu bwd", ops.geglu_bwd(x, dout), xr.grad)
    a, b = mk(1024, 64), mk(1024, 64)
    report("silu", ops.elementwise(a, None, ops.EW_SILU), F.silu(a.float()))
    ar = a.float().detach().requires_grad_(True)
    F.silu(ar).backward(b.float())
    report("silu bwd", ops.elementwise(b, a, ops.EW_SILU_BWD), ar.grad)
    report("add", ops.elementwise(a, b, ops.EW_ADD), a.float() + b.float())
    # layout
    x4 = torch.randn(2, 4, 8, 16, device=dev)
    nh = ops.nchw_to_nhwc(x4, 64)
    ref = torch.zeros(2 * 8 * 16, 64, device=dev)
    ref[:, :4] = x4.permute(0, 2, 3, 1).reshape(-1, 4)
    report("nchw_to_nhwc", nh, ref.bfloat16().float())
    back = ops.nhwc_to_nchw(nh, 2, 4, 8, 16)
    report("nhwc_to_nchw", back, x4.bfloat16().float())
    N, H, W, C = 2, 8, 8, 64
    x = mk(N * H * W, C)
    up = ops.upsample2x(x, N, H, W, C)
    ref = F.interpolate(x.float().reshape(N, H, W, C).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    report("upsample2x", up, ref.permute(0, 2, 3, 1).reshape(-1, C))
    dy = mk(N * 4 * H * W, C)
    dxr = dy.float().reshape(N, H, 2, W, 2, C).sum(dim=(2, 4)).reshape(-1, C)
    report("upsample2x bwd", ops.upsample2x(dy, N, H, W, C, backward=True), dxr)
    ps = ops.phase_split2(x, N, H, W, C)
    xx = x.reshape(N, H // 2, 2, W // 2, 2, C).permute(2, 4, 0, 1, 3, 5).reshape(-1, C)
    report("phase_split", ps, xx.float())
    report("phase_split inverse", ops.phase_split2(ps, N, H, W, C, inverse=True), x.float())
    report("colsum", ops.colsum(mk(3000, 320)), None) if False else None
    xm = mk(3000, 320)
    report("colsum", ops.colsum(xm), xm.float().sum(0))
    # stride-2 conv through phase planes vs torch conv2d
    N, H, W, Ci, Co = 2, 16, 16, 64, 128
    x = mk(N * H * W, Ci, s=0.5)
    w = (torch.randn(Co, Ci, 3, 3, device=dev) * 0.05)
    planes = ops.phase_split2(x, N, H, W, Ci).reshape(4 * N, H // 2, W // 2, Ci)
    taps = []
    for ky in range(3):
        for kx in range(3):
            dy_, dx_ = ky - 1, kx - 1
            py, px = dy_ & 1, dx_ & 1
            taps.append(((py * 2 + px) * N, -1 if dy_ == -1 else 0, -1 if dx_ == -1 else 0))
    wp = w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).to(torch.bfloat16).contiguous()
    y = ops.conv3x3_nhwc(planes, wp, taps=taps, n_out_img=N)
    ref = F.conv2d(x.float().reshape(N, H, W, Ci).permute(0, 3, 1, 2), wp.float().reshape(Co, 3, 3, Ci).permute(0, 3, 1, 2),
                   stride=2, padding=1)
    report("conv stride2", y, ref.permute(0, 2, 3, 1).reshape(-1, Co))
    # conv with N=4 outputs padded to 16 (conv_out) and K from 64-padded 4-channel input (conv_in)
    N, H, W = 2, 16, 16
    x = mk(N * H * W, 64, s=0.5)
    w = torch.zeros(16, 9 * 64, device=dev)
    w[:4] = torch.randn(4, 9 * 64, device=dev) * 0.05
    wp = w.to(torch.bfloat16)
    y = ops.conv3x3_nhwc(x.reshape(N, H, W, 64), wp)
    ref = F.conv2d(x.float().reshape(N, H, W, 64).permute(0, 3, 1, 2), wp.float().reshape(16, 3, 3, 64).permute(0, 3, 1, 2), padding=1)
    report("conv Cout16", y, ref.permute(0, 2, 3, 1).reshape(-1, 16))
    print("DONE glue")


def case_glue_perf():
    for (N, HW, C) in [(16, 128 * 128, 320), (16, 64 * 64, 640), (16, 32 * 32, 1280)]:
        x = mk(N * HW, C)
        gamma = torch.ones(C, device=dev)
        beta = torch.zeros(C, device=dev)
        us = timeit(lambda: ops.groupnorm_fwd(x, N, HW, C, 32, 1e-5, gamma, beta, True))
        by = x.numel() * 2 * 2
        print(f"perf gn+silu fwd N{N} HW{HW} C{C}: {us:.1f} us {by/us/1e3:.0f} GB/s (4 B/elem algorithmic)", flush=True)
        y, stats = ops.groupnorm_fwd(x, N, HW, C, 32, 1e-5, gamma, beta, True)
        us = timeit(lambda: ops.groupnorm_bwd(x, x, N, HW, C, 32, gamma, beta, stats, True))
        print(f"perf gn+silu bwd: {us:.1f} us {x.numel()*2*3/us/1e3:.0f} GB/s (6 B/elem algorithmic)", flush=True)
    for (M, C) in [(65536, 640), (16384, 1280)]:
        x = mk(M, C)
        gamma = torch.ones(C, device=dev)
        beta = torch.zeros(C, device=dev)
        us = timeit(lambda: ops.layernorm_fwd(x, gamma, beta))
        print(f"perf ln fwd M{M} C{C}: {us:.1f} us {x.numel()*4/us/1e3:.0f} GB/s", flush=True)
        y, stats = ops.layernorm_fwd(x, gamma, beta)
        us = timeit(lambda: ops.layernorm_bwd(x, x, gamma, stats, dres=x))
        print(f"perf ln bwd(+dres): {us:.1f} us {x.numel()*8/us/1e3:.0f} GB/s", flush=True)
        xx = mk(M, 8 * C)
        us = timeit(lambda: ops.geglu_fwd(xx))
        print(f"perf geglu fwd M{M} F{4*C}: {us:.1f} us {xx.numel()*3/us/1e3:.0f} GB/s", flush=True)
    print("DONE glue_perf")


if __name__ == "__main__":
    for c in sys.argv[1:]:
        t0 = time.time()
        try:
            globals()["case_" + c]()
        except Exception as e:  # keep going so one call reports every case
            import traceback

            traceback.print_exc()
            print(f"CASE {c} FAILED: {e}", flush=True)
        print(f"case {c} took {time.time()-t0:.1f}s", flush=True)
