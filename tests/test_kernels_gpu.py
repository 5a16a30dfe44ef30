"""GPU parity of every dense / glue kernel, called through the C ABI (uwudiff_b200.ops -> libuwu_b200.so), against a plain
PyTorch fp32 evaluation of the same bf16 inputs.

Tolerances (north_star: per-layer activations / gradients <= 1e-2 relative, fp32 accumulate):
  * fp32 outputs of bf16 operands: 2e-5 of the reference max (only summation order differs),
  * bf16 outputs: 8e-3 of the reference max (one bf16 rounding of the result = 2^-8 relative worst case),
  * attention: 1e-2 (P is rounded to bf16 before the PV product, as in every flash kernel).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"
TOL_F32, TOL_BF16, TOL_ATTN = 2e-5, 8e-3, 1e-2


@pytest.fixture(scope="module")
def ops():
    from uwudiff_b200 import ops as o

    torch.manual_seed(0)
    return o


def mk(*shape, s=0.5):
    return (torch.randn(*shape, device=DEV) * s).to(torch.bfloat16)


def relerr(got, ref):
    got = got.float()
    assert torch.isfinite(got).all(), "non-finite output"
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-9)).item()


# ------------------------------------------------------------------------------------------------------------------
# tcgen05 GEMM: operand layouts, tile widths, ragged edges, epilogues
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn,odt", [
    (128, 128, 64, 128, torch.float32), (128, 128, 256, 128, torch.float32), (256, 320, 512, 160, torch.bfloat16),
    (300, 200, 136, 0, torch.float32), (1024, 1280, 1280, 0, torch.bfloat16), (16, 1280, 320, 0, torch.bfloat16),
    (4096, 256, 2048, 256, torch.bfloat16), (512, 48, 64, 48, torch.float32), (1232, 640, 2048, 0, torch.bfloat16),
    (16384, 1280, 1280, 0, torch.bfloat16)])
def test_gemm_row_nk(ops, M, N, K, bn, odt):
    a, b = mk(M, K), mk(N, K)
    out = ops.gemm(a, b, M, N, K, out_dtype=odt, block_n=bn)
    assert relerr(out, a.float() @ b.float().t()) < (TOL_F32 if odt == torch.float32 else TOL_BF16)


@pytest.mark.parametrize("M,N,K,layout", [(38528, 512, 256, "nk"), (38500, 640, 192, "nk"), (40960, 1280, 320, "kn"),
                                            (38528 + 64, 256, 128, "kn"), (65536, 320, 128, "nk")])
def test_gemm_large_tile_counts(ops, M, N, K, layout):
    """Shapes with >= 2 tiles per SM (with UWU_GEMM_CLUSTER=1 these run as 2-CTA clusters sharing every B tile through TMA
    multicast: odd row-tile counts, ragged M, both B layouts)."""
    from uwudiff_b200._lib import B_KN

    a = mk(M, K)
    bias, res = torch.randn(N, device=DEV), mk(M, N)
    if layout == "nk":
        b = mk(N, K)
        out = ops.gemm(a, b, M, N, K, bias=bias, residual=res)
        ref = a.float() @ b.float().t()
    else:
        b = mk(K, N)
        out = ops.gemm(a, b, M, N, K, b_layout=B_KN, ldb=N, bias=bias, residual=res)
        ref = a.float() @ b.float()
    assert relerr(out, ref + bias + res.float()) < TOL_BF16


def test_gemm_epilogues(ops):
    M, N, K = 512, 320, 256
    a, b = mk(M, K), mk(N, K)
    bias, res, brows = torch.randn(N, device=DEV), mk(M, N), torch.randn(M // 128, N, device=DEV)
    ref = 0.5 * (a.float() @ b.float().t()) + bias + res.float() + brows.repeat_interleave(128, 0)
    out = ops.gemm(a, b, M, N, K, out_dtype=torch.float32, bias=bias, residual=res, bias_rows=brows, rows_per_bias=128,
                   alpha=0.5)
    assert relerr(out, ref) < TOL_F32
    out2 = ops.gemm(a, b, M, N, K, out=out.clone(), accumulate=True)
    assert relerr(out2, out + a.float() @ b.float().t()) < TOL_F32
    o1 = torch.empty(M, 160, device=DEV, dtype=torch.bfloat16)
    o2 = torch.empty(M, 160, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, M, N, K, out=o1, out2=o2, n_split=160, block_n=160)
    r = a.float() @ b.float().t()
    assert relerr(o1, r[:, :160]) < TOL_BF16 and relerr(o2, r[:, 160:]) < TOL_BF16


@pytest.mark.parametrize("M,N,K,ldr", [(65536, 640, 640, 640), (16384, 1280, 1280, 1280), (40000, 1000, 512, 1024), (300, 640, 1280, 640),
                                          (33000, 200, 256, 208), (70000, 328, 1536, 328)])
def test_gemm_residual_prefetched_by_tma_over_many_tiles(ops, M, N, K, ldr):
    """Short reductions take the epilogue whose residual rows are requested by TMA one tile ahead (K <= 1536): several tiles
    per CTA, a ragged last column tile that has FEWER 64-column chunks than the next tile of the same CTA (N = 640 -> 256 +
    256 + 128: the slots the short tile never consumed must still be requested), rows past M, a padded residual pitch."""
    a, b = mk(M, K), mk(N, K)
    bias = torch.randn(N, device=DEV)
    resbuf = mk(M, ldr)
    res = resbuf[:, :N]
    out = ops.gemm(a, b, M, N, K, bias=bias, residual=res)
    ref = a.float() @ b.float().t() + bias + res.float()
    assert relerr(out, ref) < TOL_BF16
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (256, 320, 512), (1280, 1280, 4096), (200, 136, 328)])
def test_gemm_a_col_major(ops, M, N, K):
    from uwudiff_b200._lib import A_COL

    at, b = mk(K, M), mk(N, K)
    out = ops.gemm(at, b, M, N, K, a_layout=A_COL, out_dtype=torch.float32)
    assert relerr(out, at.float().t() @ b.float().t()) < TOL_F32


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 128, 0), (256, 320, 512, 160), (1024, 1280, 1280, 0), (200, 136, 328, 0)])
def test_gemm_b_kn(ops, M, N, K, bn):
    from uwudiff_b200._lib import B_KN

    a, bt = mk(M, K), mk(K, N)
    out = ops.gemm(a, bt, M, N, K, b_layout=B_KN, out_dtype=torch.float32, block_n=bn)
    assert relerr(out, a.float() @ bt.float()) < TOL_F32


def test_gemm_weight_gradient_form(ops):
    from uwudiff_b200._lib import A_COL, B_KN

    Mtok, Co, Ci = 4096, 320, 640
    dy, x = mk(Mtok, Co), mk(Mtok, Ci)
    out = ops.gemm(dy, x, Co, Ci, Mtok, a_layout=A_COL, b_layout=B_KN, out_dtype=torch.float32)
    assert relerr(out, dy.float().t() @ x.float()) < TOL_F32


@pytest.mark.parametrize("Mtok,Co,Ci", [(4096, 320, 640), (16384, 1280, 1280), (8192, 640, 640), (1232, 128, 2048), (333 * 8, 200, 136)])
def test_gemm_stream_k_weight_gradient(ops, Mtok, Co, Ci):
    """Stream-K: the (tile, k-block) space is cut evenly over the SMs; partial tiles meet through fp32 atomics."""
    from uwudiff_b200._lib import A_COL, B_KN

    dy, x = mk(Mtok, Co), mk(Mtok, Ci)
    ref = dy.float().t() @ x.float()
    out = torch.full((Co, Ci), 7.0, device=DEV)  # stale contents must be cleared by the library
    ops.gemm(dy, x, Co, Ci, Mtok, a_layout=A_COL, b_layout=B_KN, out=out, stream_k=1)
    assert relerr(out, ref) < TOL_F32
    ops.gemm(dy, x, Co, Ci, Mtok, a_layout=A_COL, b_layout=B_KN, out=out, stream_k=1, accumulate=True)
    assert relerr(out, 2 * ref) < TOL_F32
    auto = ops.gemm(dy, x, Co, Ci, Mtok, a_layout=A_COL, b_layout=B_KN, out_dtype=torch.float32)  # stream_k=-1: heuristic
    assert relerr(auto, ref) < TOL_F32


def test_gemm_stream_k_row_major(ops):
    M, N, K = 300, 200, 4096 + 64
    a, b = mk(M, K), mk(N, K)
    out = ops.gemm(a, b, M, N, K, out_dtype=torch.float32, stream_k=1)
    assert relerr(out, a.float() @ b.float().t()) < TOL_F32


def test_gemm_rejects_cpu_and_bad_dtype(ops):
    from uwudiff_b200._lib import UwuError

    with pytest.raises(UwuError):
        ops.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16), 128, 128, 64)
    with pytest.raises(AssertionError):
        ops.gemm(torch.zeros(128, 64, device=DEV), torch.zeros(128, 64, device=DEV), 128, 128, 64)


# ------------------------------------------------------------------------------------------------------------------
# implicit-GEMM convolution
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Nimg,H,W,C1,C2,Co", [
    (2, 16, 16, 64, 0, 64), (2, 32, 32, 128, 0, 320), (1, 128, 128, 64, 0, 64), (4, 8, 8, 64, 64, 128),
    (2, 64, 64, 320, 640, 320), (3, 16, 16, 64, 0, 16)])
def test_conv3x3_matches_conv2d(ops, Nimg, H, W, C1, C2, Co):
    x = mk(Nimg, H, W, C1)
    x2 = mk(Nimg, H, W, C2) if C2 else None
    w = mk(Co, C1 + C2, 3, 3) * 0.2
    bias = torch.randn(Co, device=DEV)
    xin = torch.cat([x, x2], -1) if C2 else x
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    wp = w.permute(0, 2, 3, 1).reshape(Co, 9 * (C1 + C2)).contiguous()
    out = ops.conv3x3_nhwc(x, wp, x2=x2, bias=bias, out_dtype=torch.float32)
    assert relerr(out, ref) < TOL_F32


def test_conv_stride2_through_phase_planes(ops):
    N, H, W, Ci, Co = 2, 16, 16, 64, 128
    x = mk(N * H * W, Ci)
    wp = (torch.randn(Co, 3, 3, Ci, device=DEV) * 0.05).reshape(Co, 9 * Ci).to(torch.bfloat16).contiguous()
    planes = ops.phase_split2(x, N, H, W, Ci).reshape(4 * N, H // 2, W // 2, Ci)
    taps = [((((ky - 1) & 1) * 2 + ((kx - 1) & 1)) * N, -1 if ky == 0 else 0, -1 if kx == 0 else 0)
            for ky in range(3) for kx in range(3)]
    y = ops.conv3x3_nhwc(planes, wp, taps=taps, n_out_img=N, out_dtype=torch.float32)
    ref = F.conv2d(x.float().reshape(N, H, W, Ci).permute(0, 3, 1, 2), wp.float().reshape(Co, 3, 3, Ci).permute(0, 3, 1, 2),
                   stride=2, padding=1)
    assert relerr(y, ref.permute(0, 2, 3, 1).reshape(-1, Co)) < TOL_F32


# ------------------------------------------------------------------------------------------------------------------
# flash attention forward / backward (d = 64)
# ------------------------------------------------------------------------------------------------------------------
ATTN_SHAPES = [(1, 1, 128, 128), (1, 1, 256, 128), (1, 1, 128, 256), (2, 3, 256, 256), (2, 2, 1024, 1024), (2, 2, 64, 64),
               (2, 5, 256, 77), (1, 2, 320, 200),
               # short-key (cross-attention) kernels: Lk <= 128 with Lq_pad >= 2 * padded Lk; > 1024 queries = atomic dK/dV path
               (2, 3, 1024, 77), (1, 2, 1100, 77), (1, 2, 2500, 100), (1, 1, 300, 3), (1, 2, 333, 128)]


def attn_ref(q, k, v, B, heads, Lq, Lk, dout=None, d=64):
    def split(t, L):
        return t.float().reshape(B, L, heads, d).transpose(1, 2).detach().requires_grad_(True)

    qf, kf, vf = split(q, Lq), split(k, Lk), split(v, Lk)
    s = (qf @ kf.transpose(-1, -2)) * d ** -0.5
    o = torch.softmax(s, dim=-1) @ vf
    lse = torch.logsumexp(s, dim=-1)
    o2 = o.transpose(1, 2).reshape(B * Lq, heads * d)
    if dout is None:
        return o2.detach(), lse.detach()
    o2.backward(dout.float())
    back = lambda t, L: t.grad.transpose(1, 2).reshape(B * L, heads * d)  # noqa: E731
    return o2.detach(), lse.detach(), back(qf, Lq), back(kf, Lk), back(vf, Lk)


@pytest.mark.parametrize("B,heads,Lq,Lk", ATTN_SHAPES)
def test_attention_forward(ops, B, heads, Lq, Lk):
    C = heads * 64
    q, k, v = mk(B * Lq, C, s=1.0), mk(B * Lk, C, s=1.0), mk(B * Lk, C, s=1.0)
    o, lse = ops.attn_fwd(q, k, v, B, heads, Lq, Lk)
    ro, rlse = attn_ref(q, k, v, B, heads, Lq, Lk)
    assert relerr(o, ro) < TOL_ATTN
    Lp = (Lq + 127) // 128 * 128
    assert relerr(lse.view(B, heads, Lp)[:, :, :Lq], rlse) < 1e-4


def test_attention_forward_strided_qkv(ops):
    B, heads, L = 2, 2, 256
    C = heads * 64
    qkv = mk(B * L, 3 * C, s=1.0)
    o, _ = ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads, L, L)
    ro, _ = attn_ref(qkv[:, :C].contiguous(), qkv[:, C:2 * C].contiguous(), qkv[:, 2 * C:].contiguous(), B, heads, L, L)
    assert relerr(o, ro) < TOL_ATTN


@pytest.mark.parametrize("B,heads,Lq,Lk", ATTN_SHAPES)
def test_attention_backward(ops, B, heads, Lq, Lk):
    C = heads * 64
    q, k, v, do = mk(B * Lq, C, s=1.0), mk(B * Lk, C, s=1.0), mk(B * Lk, C, s=1.0), mk(B * Lq, C, s=1.0)
    o, lse = ops.attn_fwd(q, k, v, B, heads, Lq, Lk)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, B, heads, Lq, Lk)
    _, _, rdq, rdk, rdv = attn_ref(q, k, v, B, heads, Lq, Lk, do)
    assert relerr(dq, rdq) < TOL_ATTN and relerr(dk, rdk) < TOL_ATTN and relerr(dv, rdv) < TOL_ATTN


@pytest.mark.parametrize("B,heads,Lq,Lk", [(3, 10, 1024, 1024), (1, 40, 600, 600), (2, 20, 640, 640), (10, 20, 256, 77), (5, 31, 384, 384)])
def test_attention_backward_persistent_ctas_walk_several_items(ops, B, heads, Lq, Lk):
    """More (key tile, head, batch) items than SMs: every CTA of the persistent backward kernel walks two or more items — the
    operand ring, the S/dP and P/dS hand-shakes and the dQ drain run through the item boundaries, the K|V buffers alternate,
    the dK/dV accumulators are drained and re-armed (ragged last key / query tiles included)."""
    C = heads * 64
    q, k, v, do = mk(B * Lq, C, s=1.0), mk(B * Lk, C, s=1.0), mk(B * Lk, C, s=1.0), mk(B * Lq, C, s=1.0)
    o, lse = ops.attn_fwd(q, k, v, B, heads, Lq, Lk)
    _, _, rdq, rdk, rdv = attn_ref(q, k, v, B, heads, Lq, Lk, do)
    for _ in range(2):  # twice: the second launch starts from whatever the first left in TMEM / shared memory
        dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, B, heads, Lq, Lk)
        assert relerr(dq, rdq) < TOL_ATTN and relerr(dk, rdk) < TOL_ATTN and relerr(dv, rdv) < TOL_ATTN


# head dims other than 64 (SD-1.5: 40 / 80 / 160, DiT-XL/2: 72) run on the mma.sync kernels of attn_any.cu
ATTN_ANY_SHAPES = [(2, 8, 256, 77, 40), (1, 3, 1300, 77, 40), (2, 8, 256, 256, 40), (1, 8, 1024, 1024, 40), (2, 8, 256, 77, 80), (2, 4, 64, 64, 160), (1, 2, 200, 77, 160),
                   (3, 16, 256, 256, 72), (1, 3, 130, 70, 72), (1, 2, 320, 200, 128), (2, 2, 96, 96, 32), (1, 1, 64, 64, 8)]


@pytest.mark.parametrize("B,heads,Lq,Lk,d", ATTN_ANY_SHAPES)
def test_attention_any_head_dim(ops, B, heads, Lq, Lk, d):
    C = heads * d
    q, k, v, do = mk(B * Lq, C, s=1.0), mk(B * Lk, C, s=1.0), mk(B * Lk, C, s=1.0), mk(B * Lq, C, s=1.0)
    o, lse = ops.attn_fwd(q, k, v, B, heads, Lq, Lk, head_dim=d)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, B, heads, Lq, Lk, head_dim=d)
    ro, rlse, rdq, rdk, rdv = attn_ref(q, k, v, B, heads, Lq, Lk, do, d=d)
    assert relerr(o, ro) < TOL_ATTN
    Lp = (Lq + 127) // 128 * 128
    assert relerr(lse.view(B, heads, Lp)[:, :, :Lq], rlse) < 1e-4
    assert relerr(dq, rdq) < TOL_ATTN and relerr(dk, rdk) < TOL_ATTN and relerr(dv, rdv) < TOL_ATTN


def test_attention_any_strided_qkv(ops):
    B, heads, L, d = 2, 8, 192, 40
    C = heads * d
    qkv = mk(B * L, 3 * C, s=1.0)
    do = mk(B * L, C, s=1.0)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    o, lse = ops.attn_fwd(q, k, v, B, heads, L, L, head_dim=d)
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(q, k, v, o, do, lse, B, heads, L, L, head_dim=d, dq=dqkv[:, :C], dk=dqkv[:, C:2 * C], dv=dqkv[:, 2 * C:])
    ro, _, rdq, rdk, rdv = attn_ref(q.contiguous(), k.contiguous(), v.contiguous(), B, heads, L, L, do, d=d)
    assert relerr(o, ro) < TOL_ATTN
    assert relerr(dqkv, torch.cat([rdq, rdk, rdv], dim=1)) < TOL_ATTN


@pytest.mark.parametrize("short_fwd,short_bwd,Lq", [("1", "1", 512), ("0", "0", 512), ("1", "0", 512), ("1", "1", 200), ("0", "1", 1500)])
def test_attention_short_key_paths(ops, monkeypatch, short_fwd, short_bwd, Lq):
    """Cross-attention shapes: the short-key forward (default) and opt-in backward, the tcgen05 kernels, and their mixes
    (both use the same natural-log lse) against the fp32 reference, with strided (fused KV) operands."""
    monkeypatch.setenv("UWU_ATTN_SHORT", short_fwd)
    monkeypatch.setenv("UWU_ATTN_SHORT_BWD", short_bwd)
    B, heads, Lk = 2, 4, 77
    C = heads * 64
    q, kv, do = mk(B * Lq, C, s=1.0), mk(B * Lk, 2 * C, s=1.0), mk(B * Lq, C, s=1.0)
    k, v = kv[:, :C], kv[:, C:]
    o, lse = ops.attn_fwd(q, k, v, B, heads, Lq, Lk)
    dkv = torch.empty_like(kv)
    dq, _, _ = ops.attn_bwd(q, k, v, o, do, lse, B, heads, Lq, Lk, dk=dkv[:, :C], dv=dkv[:, C:])
    ro, rlse, rdq, rdk, rdv = attn_ref(q, k.contiguous(), v.contiguous(), B, heads, Lq, Lk, do)
    assert relerr(o, ro) < TOL_ATTN and relerr(dq, rdq) < TOL_ATTN
    assert relerr(dkv, torch.cat([rdk, rdv], dim=1)) < TOL_ATTN
    assert relerr(lse.view(B, heads, -1)[:, :, :Lq], rlse) < 1e-4


def test_attention_rejects_bad_head_dim(ops):
    q = mk(128, 36, s=1.0)
    with pytest.raises((ValueError, RuntimeError, NotImplementedError)):
        ops.attn_fwd(q, q, q, 1, 1, 128, 128, head_dim=36)


# ------------------------------------------------------------------------------------------------------------------
# GroupNorm(+SiLU) / LayerNorm forward + backward
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,H,W,C,G,silu,eps", [
    (2, 16, 16, 64, 32, True, 1e-5), (3, 8, 8, 320, 32, True, 1e-5), (2, 32, 32, 640, 32, False, 1e-6),
    (2, 16, 16, 1920, 32, True, 1e-5), (1, 64, 64, 960, 32, True, 1e-5)])
def test_groupnorm_silu(ops, N, H, W, C, G, silu, eps):
    HW = H * W
    x, dy, dres = mk(N * HW, C, s=1.0), mk(N * HW, C, s=1.0), mk(N * HW, C, s=1.0)
    gamma = torch.randn(C, device=DEV) * 0.5 + 1
    beta = torch.randn(C, device=DEV) * 0.5
    y, stats = ops.groupnorm_fwd(x, N, HW, C, G, eps, gamma, beta, silu)
    xr = x.float().reshape(N, HW, C).permute(0, 2, 1).detach().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.group_norm(xr, G, gr, br, eps)
    yr = F.silu(yr) if silu else yr
    yr2 = yr.permute(0, 2, 1).reshape(N * HW, C)
    assert relerr(y, yr2.detach()) < TOL_BF16
    yr2.backward(dy.float())
    dgamma, dbeta = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dx = ops.groupnorm_bwd(x, dy, N, HW, C, G, gamma, beta, stats, silu, dres=dres, dgamma=dgamma, dbeta=dbeta)
    assert relerr(dx, xr.grad.permute(0, 2, 1).reshape(N * HW, C) + dres.float()) < TOL_BF16
    assert relerr(dgamma, gr.grad) < 1e-3 and relerr(dbeta, br.grad) < 1e-3
    # frozen affine (no parameter gradients): the one-launch backward; must agree with the three-kernel path bit for bit
    # up to the summation order of the per-group sums, also when called repeatedly (self-resetting image barrier)
    want = xr.grad.permute(0, 2, 1).reshape(N * HW, C) + dres.float()
    for _ in range(3):
        dx2 = ops.groupnorm_bwd(x, dy, N, HW, C, G, gamma, beta, stats, silu, dres=dres)
        assert relerr(dx2, want) < TOL_BF16
        y2, stats2 = ops.groupnorm_fwd(x, N, HW, C, G, eps, gamma, beta, silu)
        assert torch.equal(y2, y) and torch.equal(stats2, stats)


def test_groupnorm_one_launch_matches_three_kernel_path(ops, monkeypatch):
    """Step shapes (16 images): one-launch forward / backward against the three-kernel path; the statistics may differ only by
    the chunking of the fp32 sums."""
    N, HW, C, G = 16, 1024, 1280, 32
    x, dy = mk(N * HW, C, s=1.0), mk(N * HW, C, s=1.0)
    gamma, beta = torch.randn(C, device=DEV) * 0.5 + 1, torch.randn(C, device=DEV) * 0.5
    y1, st1 = ops.groupnorm_fwd(x, N, HW, C, G, 1e-5, gamma, beta, True)
    dx1 = ops.groupnorm_bwd(x, dy, N, HW, C, G, gamma, beta, st1, True)
    monkeypatch.setattr(ops, "_GN_FUSED", False)
    y0, st0 = ops.groupnorm_fwd(x, N, HW, C, G, 1e-5, gamma, beta, True)
    dx0 = ops.groupnorm_bwd(x, dy, N, HW, C, G, gamma, beta, st0, True)
    assert relerr(st1, st0) < 1e-5 and relerr(y1, y0) < 1e-2 and relerr(dx1, dx0) < 1e-2
    assert (y1 != y0).float().mean() < 1e-2  # only bf16 rounding flips from the last-bit differences of the statistics


@pytest.mark.parametrize("M,C", [(256, 64), (1000, 640), (4096, 1280), (77, 128), (16384, 1280), (333, 320)])
def test_layernorm(ops, M, C):
    x, dy, dres = mk(M, C, s=1.0), mk(M, C, s=1.0), mk(M, C, s=1.0)
    gamma = torch.randn(C, device=DEV) * 0.5 + 1
    beta = torch.randn(C, device=DEV) * 0.5
    y, stats = ops.layernorm_fwd(x, gamma, beta, 1e-5)
    xr = x.float().detach().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), gr, br, 1e-5)
    assert relerr(y, yr.detach()) < TOL_BF16
    yr.backward(dy.float())
    dgamma, dbeta = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dx = ops.layernorm_bwd(x, dy, gamma, stats, dres=dres, dgamma=dgamma, dbeta=dbeta, accumulate=True)
    assert relerr(dx, xr.grad + dres.float()) < TOL_BF16
    assert relerr(dgamma, gr.grad) < 1e-3 and relerr(dbeta, br.grad) < 1e-3
    # accumulate=True adds onto existing gradients
    ops.layernorm_bwd(x, dy, gamma, stats, dres=dres, dgamma=dgamma, dbeta=dbeta, accumulate=True)
    assert relerr(dgamma, 2 * gr.grad) < 1e-3 and relerr(dbeta, 2 * br.grad) < 1e-3


def test_layernorm_adaln_modulate(ops):
    """DiT adaLN: y = LN(x) * (1 + scale[b]) + shift[b] with one (scale, shift) row per `rows_per_mod` tokens."""
    B, L, C = 3, 64, 256
    x = mk(B * L, C, s=1.0)
    scale, shift = torch.randn(B, C, device=DEV) * 0.3, torch.randn(B, C, device=DEV) * 0.3
    y, _ = ops.layernorm_fwd(x, None, None, 1e-6, mod_scale=scale, mod_shift=shift, rows_per_mod=L)
    ref = F.layer_norm(x.float(), (C,), None, None, 1e-6).view(B, L, C) * (1 + scale[:, None]) + shift[:, None]
    assert relerr(y, ref.reshape(B * L, C)) < TOL_BF16


# ------------------------------------------------------------------------------------------------------------------
# elementwise / layout glue
# ------------------------------------------------------------------------------------------------------------------
def test_geglu(ops):
    M, Fh = 512, 256
    x, dout = mk(M, 2 * Fh, s=1.0), mk(M, Fh, s=1.0)
    xr = x.float().detach().requires_grad_(True)
    h, g = xr.chunk(2, dim=-1)
    yr = h * F.gelu(g)
    assert relerr(ops.geglu_fwd(x), yr.detach()) < TOL_BF16
    yr.backward(dout.float())
    assert relerr(ops.geglu_bwd(x, dout), xr.grad) < TOL_BF16


def test_elementwise_and_layout(ops):
    a, b = mk(1024, 64, s=1.0), mk(1024, 64, s=1.0)
    assert relerr(ops.elementwise(a, None, ops.EW_SILU), F.silu(a.float())) < TOL_BF16
    ar = a.float().detach().requires_grad_(True)
    F.silu(ar).backward(b.float())
    assert relerr(ops.elementwise(b, a, ops.EW_SILU_BWD), ar.grad) < TOL_BF16
    assert relerr(ops.elementwise(a, b, ops.EW_ADD), a.float() + b.float()) < TOL_BF16
    x4 = torch.randn(2, 4, 8, 16, device=DEV)
    nh = ops.nchw_to_nhwc(x4, 64)
    ref = torch.zeros(2 * 8 * 16, 64, device=DEV)
    ref[:, :4] = x4.permute(0, 2, 3, 1).reshape(-1, 4)
    assert torch.equal(nh.float(), ref.bfloat16().float())
    assert torch.equal(ops.nhwc_to_nchw(nh, 2, 4, 8, 16), x4.bfloat16().float())
    N, H, W, C = 2, 8, 8, 64
    x = mk(N * H * W, C)
    up = ops.upsample2x(x, N, H, W, C)
    ref = F.interpolate(x.float().reshape(N, H, W, C).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(up.float(), ref.permute(0, 2, 3, 1).reshape(-1, C))
    dy = mk(N * 4 * H * W, C)
    dxr = dy.float().reshape(N, H, 2, W, 2, C).sum(dim=(2, 4)).reshape(-1, C)
    assert relerr(ops.upsample2x(dy, N, H, W, C, backward=True), dxr) < TOL_BF16
    ps = ops.phase_split2(x, N, H, W, C)
    xx = x.reshape(N, H // 2, 2, W // 2, 2, C).permute(2, 4, 0, 1, 3, 5).reshape(-1, C)
    assert torch.equal(ps, xx)
    assert torch.equal(ops.phase_split2(ps, N, H, W, C, inverse=True), x)
    xm = mk(3000, 320)
    assert relerr(ops.colsum(xm), xm.float().sum(0)) < 1e-4


def test_sincos_embedding(ops):
    t = torch.tensor([0., 1., 37., 999.], device=DEV)
    for dim in (320, 256):
        half = dim // 2
        freq = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(half, device=DEV) / half)
        a = t[:, None] * freq[None]
        ref = torch.cat([torch.cos(a), torch.sin(a)], -1)  # flip_sin_to_cos=True
        assert relerr(ops.sincos_embed(t, dim, True), ref) < TOL_BF16


# ------------------------------------------------------------------------------------------------------------------
# LyCORIS operand folding / gradient kernels, fused optimizer
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ol,ok,im,inn", [(10, 64, 10, 64), (5, 128, 5, 64), (20, 64, 32, 64), (3, 16, 2, 24)])
def test_lokr_fold_and_grads(ops, ol, ok, im, inn):
    N, K = ol * ok, im * inn
    W = torch.randn(N, K, device=DEV) * 0.1
    w1, w2 = torch.randn(ol, im, device=DEV), torch.randn(ok, inn, device=DEV) * 0.1
    dst = torch.empty(N, K, device=DEV, dtype=torch.bfloat16)
    ops.fold_lokr(W, w1, w2, dst)
    assert torch.equal(dst, (W + torch.kron(w1, w2)).bfloat16())
    ops.fold_lokr(W, None, None, dst)
    assert torch.equal(dst, W.bfloat16())
    G = torch.randn(N, K, device=DEV)
    dw1, dw2 = torch.zeros_like(w1), torch.zeros_like(w2)
    ops.lokr_grad(G, w1, w2, dw1, dw2)
    G4 = G.view(ol, ok, im, inn)
    assert relerr(dw1, torch.einsum("lkin,kn->li", G4, w2)) < 1e-4
    assert relerr(dw2, torch.einsum("lkin,li->kn", G4, w1)) < 1e-4
    ops.lokr_grad(G, w1, w2, dw1, dw2)  # accumulates
    assert relerr(dw1, 2 * torch.einsum("lkin,kn->li", G4, w2)) < 1e-4


@pytest.mark.parametrize("M,ol,ok,im,inn", [(4096, 20, 64, 20, 64), (2048, 5, 2048, 5, 256), (8192, 10, 64, 10, 64),
                                             (1024, 5, 1024, 5, 128), (1000, 4, 128, 4, 16), (1232, 20, 64, 32, 64),
                                             (640, 3, 192, 7, 48)])
def test_lokr_factored_gradients(ops, M, ol, ok, im, inn):
    """Factored adapter gradients (Z / segmented-K GEMM / grouped-N GEMM / mma.sync contraction) == the einsum over the
    full weight gradient; dY is a column slice of a wider buffer (as in the fused QKV gradient)."""
    from uwudiff_b200.lycoris import lokr_factored_grads

    N, K = ol * ok, im * inn
    wide = mk(M, N + 64, s=1.0)
    dy = wide[:, 64:]  # 128-byte offset, row stride N + 64
    x = mk(M, K, s=1.0)
    w1 = torch.randn(ol, im, device=DEV)
    w2 = torch.randn(ok, inn, device=DEV) * 0.5
    w2b = w2.to(torch.bfloat16)
    dw1, dw2 = torch.zeros_like(w1), torch.zeros_like(w2)
    lokr_factored_grads(dy, x, M, w1, w2b, dw1, dw2, 1.0)
    G4 = (dy.float().t() @ x.float()).view(ol, ok, im, inn)
    r1 = torch.einsum("lkin,kn->li", G4, w2b.float())
    r2 = torch.einsum("lkin,li->kn", G4, w1)
    # Z and V are rounded to bf16 before the second contraction -> bf16-level tolerance on the result
    assert relerr(dw1, r1) < TOL_BF16 and relerr(dw2, r2) < TOL_BF16
    lokr_factored_grads(dy, x, M, w1, w2b, dw1, dw2, 0.5)  # accumulates, scaled
    assert relerr(dw1, 1.5 * r1) < TOL_BF16 and relerr(dw2, 1.5 * r2) < TOL_BF16


@pytest.mark.parametrize("M,ol,ok,im,inn", [(4096, 5, 256, 5, 1024), (8192, 5, 128, 5, 512), (5000, 4, 64, 6, 512), (4096, 8, 256, 5, 320)])
def test_lokr_mirrored_factored_gradients(ops, M, ol, ok, im, inn):
    """dY-side factored route (out_k < in_n: the FeedForward down projection, w2 256 x 1024): U = (w1^T (x) I) dY, segmented
    token-reduction GEMM for dw2, T_j = X_j w2^T, mma.sync contraction for dw1 == the einsum over the full weight gradient."""
    from uwudiff_b200.lycoris import lokr_factored_grads_mirror

    N, K = ol * ok, im * inn
    dy, x = mk(M, N, s=1.0), mk(M, K, s=1.0)
    w1 = torch.randn(ol, im, device=DEV)
    w2 = torch.randn(ok, inn, device=DEV) * 0.5
    w2b = w2.to(torch.bfloat16)
    dw1, dw2 = torch.zeros_like(w1), torch.zeros_like(w2)
    lokr_factored_grads_mirror(dy, x, M, w1, w2b, dw1, dw2, 1.0)
    G4 = (dy.float().t() @ x.float()).view(ol, ok, im, inn)
    r1 = torch.einsum("lkin,kn->li", G4, w2b.float())
    r2 = torch.einsum("lkin,li->kn", G4, w1)
    assert relerr(dw1, r1) < TOL_BF16 and relerr(dw2, r2) < TOL_BF16, (relerr(dw1, r1), relerr(dw2, r2))
    lokr_factored_grads_mirror(dy, x, M, w1, w2b, dw1, dw2, 0.5)
    assert relerr(dw1, 1.5 * r1) < TOL_BF16 and relerr(dw2, 1.5 * r2) < TOL_BF16


@pytest.mark.parametrize("M,ol,im", [(4096, 20, 20), (16384, 20, 20), (8192 + 5, 10, 10), (1232, 20, 32), (1000, 5, 5),
                                      (70, 32, 20), (3, 20, 20), (65536, 10, 10)])
def test_lokr_fused_one_pass_gradients(ops, M, ol, im):
    """One-pass tcgen05 LoKr gradients (w2 64x64: the `Attention -> lokr, factor 64` adapters of the SDXL preset) == the
    einsum over the full weight gradient G = dY^T X; dY is a column slice of a wider buffer (fused QKV), ragged last tile,
    w1 wider / taller than square, accumulation with a multiplier."""
    N, K = ol * 64, im * 64
    wide = mk(M, N + 64, s=1.0)
    dy = wide[:, 64:]
    x = mk(M, K, s=1.0)
    w1 = torch.randn(ol, im, device=DEV)
    w2 = torch.randn(64, 64, device=DEV) * 0.5
    dw1, dw2 = torch.zeros_like(w1), torch.zeros_like(w2)
    assert ops.lokr_fused_supported(ol, 64, im, 64) and not ops.lokr_fused_supported(ol, 128, im, 64)
    ops.lokr_fused_grad(dy, x, M, w1, w2, dw1, dw2, 1.0)
    G4 = (dy.double().t() @ x.double()).view(ol, 64, im, 64)
    r1 = torch.einsum("lkin,kn->li", G4, w2.to(torch.bfloat16).double()).float()
    r2 = torch.einsum("lkin,li->kn", G4, w1.to(torch.bfloat16).double()).float()
    assert relerr(dw1, r1) < TOL_BF16 and relerr(dw2, r2) < TOL_BF16, (relerr(dw1, r1), relerr(dw2, r2))
    ops.lokr_fused_grad(dy, x, M, w1, w2, dw1, dw2, 0.5)  # accumulates, scaled
    assert relerr(dw1, 1.5 * r1) < TOL_BF16 and relerr(dw2, 1.5 * r2) < TOL_BF16


def test_lora_fold_and_grads(ops):
    N, K, r, scale = 640, 320, 4, 0.25
    W = torch.randn(N, K, device=DEV) * 0.1
    up, down = torch.randn(N, r, device=DEV) * 0.1, torch.randn(r, K, device=DEV) * 0.1
    dst = torch.empty(N, K, device=DEV, dtype=torch.bfloat16)
    ops.fold_lora(W, up, down, scale, dst)
    assert relerr(dst, W + up @ down * scale) < TOL_BF16
    G = torch.randn(N, K, device=DEV)
    dup, ddown = torch.zeros_like(up), torch.zeros_like(down)
    ops.lora_grad(G, up, down, scale, dup, ddown)
    assert relerr(dup, G @ down.t() * scale) < 1e-4 and relerr(ddown, up.t() @ G * scale) < 1e-4


@pytest.mark.parametrize("N,K,r", [(640, 320, 4), (128, 2048, 8), (1280, 1280, 16), (24, 36, 1)])
def test_loha_fold_and_grads(ops, N, K, r):
    """LoHa: dW = (w1a @ w1b) o (w2a @ w2b) * scale; factor gradients from G vs autograd (fp32)."""
    scale = 1.0 / r
    g = torch.Generator(device=DEV).manual_seed(N + K + r)
    W = torch.randn(N, K, device=DEV, generator=g) * 0.1
    f = [torch.randn(s, device=DEV, generator=g).requires_grad_(True) for s in ((N, r), (r, K), (N, r), (r, K))]
    w1a, w1b, w2a, w2b = f
    dst = torch.empty(N, K, device=DEV, dtype=torch.bfloat16)
    ops.fold_loha(W, w1a.detach(), w1b.detach(), w2a.detach(), w2b.detach(), scale, dst)
    ref = W + (w1a @ w1b) * (w2a @ w2b) * scale
    assert relerr(dst, ref.detach()) < TOL_BF16
    G = torch.randn(N, K, device=DEV, generator=g)
    ref.backward(G)
    grads = [torch.zeros_like(t) for t in f]
    ops.loha_grad(G, w1a.detach(), w1b.detach(), w2a.detach(), w2b.detach(), scale, *[grads[i] for i in (0, 1, 2, 3)])
    for got, t in zip(grads, f):
        assert relerr(got, t.grad) < 1e-4
    ops.loha_grad(G, w1a.detach(), w1b.detach(), w2a.detach(), w2b.detach(), scale, *grads)   # accumulates
    assert relerr(grads[1], 2 * w1b.grad) < 1e-4


def test_fused_adamw_and_clip_match_torch():
    from uwudiff_b200.optim import FusedAdamW

    torch.manual_seed(3)
    shapes = [(20, 20), (64, 64), (1280,), (5, 5), (2048, 256), (7,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    rs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = FusedAdamW(ps, lr=1e-2, weight_decay=0.01, betas=(0.9, 0.999), max_grad_norm=1.0)
    ref = torch.optim.AdamW(rs, lr=1e-2, weight_decay=0.01, betas=(0.9, 0.999))
    for it in range(3):
        for p, r in zip(ps, rs):
            g = torch.randn_like(p)
            p.grad = g.clone()
            r.grad = g.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_(rs, 1.0)
        opt.step()
        ref.step()
        assert abs(float(opt.last_norm[0]) - norm_ref.item()) < 1e-4 * norm_ref.item()
        for p, r in zip(ps, rs):
            assert relerr(p.detach(), r.detach()) < 1e-5


@pytest.mark.parametrize("M,C,F", [(16384, 1280, 5120), (4096, 640, 2560), (300, 64, 256), (256, 128, 512)])
def test_fused_geglu_epilogues_are_bit_identical_to_the_unfused_kernels(M, C, F):
    """GEGLU in the GEMM epilogues (CTA-pair kernel): up-projection + h * gelu(g), and the down-projection's data gradient
    + GEGLU backward, against uwu_gemm followed by uwu_geglu_fwd / uwu_geglu_bwd — same MMA order, same roundings."""
    from uwudiff_b200 import ops
    from uwudiff_b200._lib import B_KN

    if not ops.geglu_kernels_available(M, F):
        pytest.skip("fused GEGLU needs the CTA-pair kernel (UWU_GEMM_PAIR=0)")
    g = torch.Generator(device="cuda").manual_seed(M + F)
    x = (torch.randn(M, C, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w1 = (torch.randn(2 * F, C, device="cuda", generator=g) * C ** -0.5).to(torch.bfloat16)
    b1 = torch.randn(2 * F, device="cuda", generator=g) * 0.1
    p_ref = ops.gemm(x, w1, M, 2 * F, C, bias=b1)
    act_ref = ops.geglu_fwd(p_ref)
    p, act = ops.gemm_geglu_fwd(x, w1, M, F, C, b1)
    assert torch.equal(p, p_ref) and torch.equal(act, act_ref)
    # and against fp32 torch (erf GELU): the kernels use a 1.5e-7-accurate erf
    ref32 = torch.nn.functional.linear(x.float(), w1.float(), b1)
    h32, g32 = ref32.chunk(2, -1)
    a32 = h32 * torch.nn.functional.gelu(g32)
    assert ((act.float() - a32).abs().max() / a32.abs().max()).item() < 8e-3
    dy = (torch.randn(M, C, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, F, device="cuda", generator=g) * F ** -0.5).to(torch.bfloat16)  # down projection [out = C, in = F]
    d_ref = ops.gemm(dy, w2, M, F, C, b_layout=B_KN, ldb=F)
    dp_ref = ops.geglu_bwd(p_ref, d_ref)
    dp = ops.gemm_geglu_bwd(dy, w2, p, M, F, C)
    assert torch.equal(dp, dp_ref)


@pytest.mark.parametrize("M,N,K", [(512, 320, 2880), (1000, 640, 2560), (384, 960, 2624), (256, 1280, 5760)])
def test_gemm_wide_tiles_two_umma_per_k_step(M, N, K):
    """block_n = 320 as two 160-wide UMMAs sharing the A tile (CTA-pair kernel, one TMEM accumulator): the conv / long-K
    classes with N a multiple of 320.  Against fp32 torch on the same bf16 operands, with bias + residual epilogue."""
    from uwudiff_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(N + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    ref = a.float() @ b.float().t() + bias + res.float()
    out = ops.gemm(a, b, M, N, K, bias=bias, residual=res)
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < 8e-3
    # identical to the 256-wide schedule up to accumulation order (same products, fp32 accumulate): compare loosely, and
    # exactly against a forced narrow run of the same kernel family
    out256 = ops.gemm(a, b, M, N, K, bias=bias, residual=res, block_n=256)
    assert ((out.float() - out256.float()).abs().max() / ref.abs().max()).item() < 8e-3


def test_conv3x3_320_channels_takes_the_wide_tile_path():
    from uwudiff_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn(2, 32, 32, 320, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(320, 320, 3, 3, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(320, device="cuda", generator=g)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, 320)
    wp = w.permute(0, 2, 3, 1).reshape(320, 9 * 320).contiguous()
    out = ops.conv3x3_nhwc(x, wp, bias=bias)
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < 8e-3
