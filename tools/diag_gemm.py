"""GPU diagnostic for the tcgen05 GEMM/conv kernel (run under gpurun, one case per process).

    python tools/diag_gemm.py <case>

Each case prints max-abs / relative error against a torch fp32 matmul/conv2d of the same bf16 inputs.
"""
import sys
import os

os.environ.setdefault("UWU_SYNTHETIC_CONDITIONING", "1")  # synthetic text-encoder outputs (no weights offline)
import itertools

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from uwudiff_b200 import ops
from uwudiff_b200._lib import A_COL, A_ROW, B_KN, B_NK

torch.manual_seed(0)
dev = "cuda"


def report(name, got, ref):
    got = got.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    bad = (~torch.isfinite(got)).sum().item()
    print(f"{name}: max_abs_err={err:.4e} ref_max={scale:.3e} rel={err/scale:.3e} nonfinite={bad}", flush=True)
    return err / scale


def mk(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(torch.bfloat16)


def case_kmajor():
    for (M, N, K, bn, odt) in [(128, 128, 64, 128, torch.float32), (128, 128, 256, 128, torch.float32),
                               (256, 320, 512, 160, torch.bfloat16), (300, 200, 136, 0, torch.float32),
                               (1024, 1280, 1280, 0, torch.bfloat16), (16, 1280, 320, 0, torch.bfloat16),
                               (4096, 256, 2048, 256, torch.bfloat16), (512, 48, 64, 48, torch.float32)]:
        a, b = mk(M, K), mk(N, K)
        ref = a.float() @ b.float().t()
        out = ops.gemm(a, b, M, N, K, out_dtype=odt, block_n=bn)
        torch.cuda.synchronize()
        report(f"kmajor M{M} N{N} K{K} bn{bn} {odt}", out, ref)
    # epilogue: bias + residual + bias_rows + alpha
    M, N, K = 512, 320, 256
    a, b = mk(M, K), mk(N, K)
    bias = torch.randn(N, device=dev)
    res = mk(M, N)
    brows = torch.randn(M // 128, N, device=dev)
    ref = 0.5 * (a.float() @ b.float().t()) + bias + res.float() + brows.repeat_interleave(128, 0)
    out = ops.gemm(a, b, M, N, K, out_dtype=torch.float32, bias=bias, residual=res, bias_rows=brows, rows_per_bias=128,
                   alpha=0.5)
    torch.cuda.synchronize()
    report("epilogue", out, ref)
    out2 = ops.gemm(a, b, M, N, K, out=out.clone(), accumulate=True)
    torch.cuda.synchronize()
    report("accumulate", out2, out + a.float() @ b.float().t())
    # split output
    o1 = torch.empty(M, 160, device=dev, dtype=torch.bfloat16)
    o2 = torch.empty(M, 160, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, b, M, N, K, out=o1, out2=o2, n_split=160, block_n=160)
    torch.cuda.synchronize()
    r = a.float() @ b.float().t()
    report("split lo", o1, r[:, :160])
    report("split hi", o2, r[:, 160:])


def _sweep(name, fn, ref, variants):
    for v in variants:
        try:
            out = fn(v)
            torch.cuda.synchronize()
            rel = report(f"{name} {v}", out, ref)
            if rel < 2e-2:
                print(f"{name}: OK with {v}", flush=True)
                return v
        except Exception as e:  # noqa
            print(f"{name} {v}: EXC {e}", flush=True)
    print(f"{name}: NO VARIANT WORKED", flush=True)
    return None


def case_acol():
    M, N, K = 128, 128, 128
    at = mk(K, M)  # stored [K, M]
    b = mk(N, K)
    ref = at.float().t() @ b.float().t()
    variants = [dict(), dict(a_lbo=1024, a_sbo=8192), dict(a_lbo=8192, a_sbo=1024, a_kadv=1024),
                dict(a_lbo=1024, a_sbo=8192, a_kadv=1024)]
    v = _sweep("acol", lambda v: ops.gemm(at, b, M, N, K, a_layout=A_COL, out_dtype=torch.float32, dbg=v), ref, variants)
    if v is not None:
        for (M, N, K) in [(256, 320, 512), (1280, 1280, 4096), (200, 136, 328)]:
            at, b = mk(K, M), mk(N, K)
            ref = at.float().t() @ b.float().t()
            out = ops.gemm(at, b, M, N, K, a_layout=A_COL, out_dtype=torch.float32, dbg=v)
            torch.cuda.synchronize()
            report(f"acol M{M} N{N} K{K}", out, ref)


def case_bkn():
    M, N, K = 128, 128, 128
    a = mk(M, K)
    bt = mk(K, N)  # stored [K, N]
    ref = a.float() @ bt.float()
    variants = [dict(), dict(b_lbo=1024, b_sbo=8192), dict(b_lbo=8192, b_sbo=1024, b_kadv=1024),
                dict(b_lbo=1024, b_sbo=8192, b_kadv=1024)]
    v = _sweep("bkn", lambda v: ops.gemm(a, bt, M, N, K, b_layout=B_KN, out_dtype=torch.float32, dbg=v), ref, variants)
    if v is not None:
        for (M, N, K, bn) in [(256, 320, 512, 160), (1024, 1280, 1280, 0), (200, 136, 328, 0)]:
            a, bt = mk(M, K), mk(K, N)
            ref = a.float() @ bt.float()
            out = ops.gemm(a, bt, M, N, K, b_layout=B_KN, out_dtype=torch.float32, block_n=bn, dbg=v)
            torch.cuda.synchronize()
            report(f"bkn M{M} N{N} K{K}", out, ref)
        # both MN-major (weight-gradient form): dW[N_out,K_in] = dY^T X
        Mtok, Co, Ci = 4096, 320, 640
        dy, x = mk(Mtok, Co), mk(Mtok, Ci)
        ref = dy.float().t() @ x.float()
        out = ops.gemm(dy, x, Co, Ci, Mtok, a_layout=A_COL, b_layout=B_KN, out_dtype=torch.float32)
        torch.cuda.synchronize()
        report("wgrad form", out, ref)


def case_conv():
    for (Nimg, H, W, C1, C2, Co) in [(2, 16, 16, 64, 0, 64), (2, 32, 32, 128, 0, 320), (1, 128, 128, 64, 0, 64),
                                     (4, 8, 8, 64, 64, 128), (2, 64, 64, 320, 640, 320), (3, 16, 16, 64, 0, 16)]:
        x = mk(Nimg, H, W, C1)
        x2 = mk(Nimg, H, W, C2) if C2 else None
        w = mk(Co, C1 + C2, 3, 3) * 0.2  # torch layout [Co, Ci, kh, kw]
        bias = torch.randn(Co, device=dev)
        xin = torch.cat([x, x2], -1) if C2 else x
        ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
        wp = w.permute(0, 2, 3, 1).reshape(Co, 9 * (C1 + C2)).contiguous()  # [Co, tap, c]
        out = ops.conv3x3_nhwc(x, wp, x2=x2, bias=bias, out_dtype=torch.float32)
        torch.cuda.synchronize()
        report(f"conv N{Nimg} H{H} W{W} C{C1}+{C2} Co{Co}", out, ref)


def case_perf():
    def bench(fn, flops, name, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        print(f"perf {name}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s", flush=True)

    for (M, N, K, bn) in [(16384, 1280, 1280, 0), (16384, 1280, 1280, 160), (16384, 1280, 1280, 128),
                          (16384, 10240, 1280, 0), (16384, 1280, 5120, 0), (65536, 640, 640, 0),
                          (65536, 5120, 640, 0), (8192, 8192, 8192, 256)]:
        a, b = mk(M, K), mk(N, K)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        bench(lambda: ops.gemm(a, b, M, N, K, out=out, block_n=bn), 2 * M * N * K, f"gemm M{M} N{N} K{K} bn{bn}")
        ref_fn = lambda: torch.matmul(a, b.t())
        bench(ref_fn, 2 * M * N * K, f"torch M{M} N{N} K{K}")
    for (Nimg, H, W, C, Co) in [(16, 32, 32, 1280, 1280), (16, 64, 64, 640, 640), (16, 128, 128, 320, 320)]:
        x = mk(Nimg, H, W, C)
        wp = mk(Co, 9 * C)
        out = torch.empty(Nimg * H * W, Co, device=dev, dtype=torch.bfloat16)
        bench(lambda: ops.conv3x3_nhwc(x, wp, out=out), 2 * Nimg * H * W * Co * 9 * C, f"conv {Nimg}x{H}x{W} C{C}->{Co}")


def case_perf12():
    """The most expensive GEMM / conv classes of the SDXL + LyCORIS step (profiles/r01_step_breakdown_441ms.log), isolated,
    next to torch.matmul (cuBLASLt) / cuDNN on the same shapes.  Run with UWU_GEMM_PAIR=0 and =1 to compare the 1-CTA and
    the CTA-pair kernels."""
    import os

    def bench(fn, flops, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        return ms * 1e3, flops / ms / 1e9

    print(f"UWU_GEMM_PAIR={os.environ.get('UWU_GEMM_PAIR', '1 (default)')}")
    print(f"{'shape':44s} {'uwu us':>9s} {'TF/s':>8s} {'torch us':>9s} {'TF/s':>8s} {'ratio':>6s}")
    lin = [(16384, 1280, 1280, "nk", True), (16384, 10240, 1280, "nk", False), (16384, 1280, 10240, "kn", False),
           (16384, 1280, 5120, "nk", True), (16384, 5120, 1280, "kn", False), (16384, 3840, 1280, "nk", False),
           (16384, 1280, 3840, "kn", False), (16384, 1280, 2048, "nk", False), (65536, 640, 640, "nk", True),
           (65536, 5120, 640, "nk", False), (65536, 640, 5120, "kn", False), (65536, 640, 2560, "nk", True),
           (8192, 8192, 8192, "nk", False)]
    for (M, N, K, bl, with_res) in lin:
        a = mk(M, K)
        b = mk(N, K) if bl == "nk" else mk(K, N)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        res = mk(M, N) if with_res else None
        bias = torch.randn(N, device=dev) if with_res else None
        fn = lambda: ops.gemm(a, b, M, N, K, out=out, b_layout=B_NK if bl == "nk" else B_KN, residual=res, bias=bias)
        us, tf = bench(fn, 2 * M * N * K)
        if bl == "nk":
            ref = (lambda: torch.addmm(res, a, b.t())) if with_res else (lambda: torch.matmul(a, b.t()))
        else:
            ref = lambda: torch.matmul(a, b)
        rus, rtf = bench(ref, 2 * M * N * K)
        tag = f"lin M{M} N{N} K{K} {bl}{' +bias+res' if with_res else ''}"
        print(f"{tag:44s} {us:9.1f} {tf:8.1f} {rus:9.1f} {rtf:8.1f} {tf / rtf:6.2f}", flush=True)
    torch.backends.cudnn.benchmark = True
    for (Nimg, H, W, C, Co) in [(16, 128, 128, 320, 320), (16, 64, 64, 640, 640), (16, 32, 32, 1280, 1280),
                                (16, 32, 32, 2560, 1280), (16, 128, 128, 640, 320)]:
        x = mk(Nimg, H, W, C)
        wp = mk(Co, 9 * C)
        out = torch.empty(Nimg * H * W, Co, device=dev, dtype=torch.bfloat16)
        fl = 2 * Nimg * H * W * Co * 9 * C
        us, tf = bench(lambda: ops.conv3x3_nhwc(x, wp, out=out), fl, iters=10)
        xc = x.permute(0, 3, 1, 2)  # channels-last NCHW view for cuDNN
        wc = wp.view(Co, 3, 3, C).permute(0, 3, 1, 2)
        rus, rtf = bench(lambda: F.conv2d(xc, wc, padding=1), fl, iters=10)
        tag = f"conv {Nimg}x{H}x{W} C{C}->{Co}"
        print(f"{tag:44s} {us:9.1f} {tf:8.1f} {rus:9.1f} {rtf:8.1f} {tf / rtf:6.2f}", flush=True)


def case_perfw():
    """Token-reduction (adapter / weight gradient) GEMMs of the step: G[N_out, K_in] = dY^T X, fp32 output, split-K schedules."""
    import os

    print(f"UWU_GEMM_PAIR={os.environ.get('UWU_GEMM_PAIR', '2 (default)')}")
    for (Co, Ci, Mtok) in [(1280, 5120, 16384), (1280, 1280, 16384), (3840, 1280, 16384), (5120, 1280, 16384),
                           (640, 640, 65536), (640, 2560, 65536), (1920, 640, 65536), (2560, 2048, 1232)]:
        dy, x = mk(Mtok, Co), mk(Mtok, Ci)
        G = torch.empty(Co, Ci, device=dev, dtype=torch.float32)
        fn = lambda: ops.gemm(dy, x, Co, Ci, Mtok, a_layout=A_COL, lda=Co, b_layout=B_KN, ldb=Ci, out=G)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        ref = dy.float().t() @ x.float()
        err = ((G - ref).abs().max() / ref.abs().max()).item()
        s.record()
        for _ in range(20):
            torch.matmul(dy.t(), x)
        e.record()
        torch.cuda.synchronize()
        rms_ = s.elapsed_time(e) / 20
        fl = 2.0 * Co * Ci * Mtok
        print(f"wgrad G[{Co},{Ci}] K{Mtok}: {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TF/s   torch bf16-out {rms_*1e3:8.1f} us "
              f"{fl/rms_/1e9:8.1f} TF/s   rel err {err:.2e}", flush=True)


if __name__ == "__main__":
    globals()["case_" + sys.argv[1]]()
    print("DONE", sys.argv[1], flush=True)
