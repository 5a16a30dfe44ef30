"""TEST-ONLY torch emulation of `uwudiff_b200.ops` (the C-ABI wrappers), monkeypatched in by the `fake_ops` fixture.

It exists so that the HOST logic of the product (module scheduling, hand-written backward order, LyCORIS bookkeeping,
gradient bucketing) can be checked against the oracle on a machine without a GPU.  It restates each kernel's contract
(layouts, leading dimensions, epilogues, accumulate semantics) in plain torch on CPU tensors.  The product never
imports this file; on a GPU box the real kernels run and are themselves checked against torch in the `-m gpu` tests.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from uwudiff_b200 import _lib
from uwudiff_b200 import ops as real_ops

BF16 = torch.bfloat16
_launches = [0]


def _strided(t, rows, cols, ld):
    return torch.as_strided(t, (rows, cols), (ld, 1), t.storage_offset())


def gemm(a, b, M, N, K, *, a_layout=0, b_layout=0, lda=None, ldb=None, out=None, out_dtype=BF16, bias=None,
         bias_rows=None, rows_per_bias=1, residual=None, alpha=1.0, accumulate=False, block_n=0, out2=None, n_split=0,
         conv=None, a2=None, dbg=None, stream_k=-1, k_segs=0, a_seg_off=0, b_seg_off=0, grp_n=0, a_grp_koff=0):
    _launches[0] += 1
    assert a.dtype == BF16 and b.dtype == BF16
    assert k_segs <= 1 and grp_n == 0, "segmented-K / grouped-N GEMMs (factored LoKr route) are CUDA-only paths"
    if a_layout == _lib.A_ROW:
        A = _strided(a, M, K, lda if lda is not None else K).float()
    elif a_layout == _lib.A_COL:
        A = _strided(a, K, M, lda if lda is not None else M).float().t()
    else:
        nbuf, H, W, C1, C2 = conv["n_img_buf"], conv["H"], conv["W"], conv["Cin1"], conv.get("Cin2", 0)
        xs = [a.reshape(nbuf, H, W, C1).float()] + ([a2.reshape(nbuf, H, W, C2).float()] if C2 else [])
        x = torch.cat(xs, dim=-1)
        n_out = M // (H * W)
        cols = []
        for (dn, dh, dw) in conv["taps"]:
            src = torch.zeros((n_out, H, W, C1 + C2))
            n_lo, n_hi = max(0, -dn), min(n_out, nbuf - dn)
            h_lo, h_hi = max(0, -dh), min(H, H - dh)
            w_lo, w_hi = max(0, -dw), min(W, W - dw)
            if n_hi > n_lo and h_hi > h_lo and w_hi > w_lo:
                src[n_lo:n_hi, h_lo:h_hi, w_lo:w_hi] = x[n_lo + dn:n_hi + dn, h_lo + dh:h_hi + dh, w_lo + dw:w_hi + dw]
            cols.append(src.reshape(M, C1 + C2))
        A = torch.cat(cols, dim=1)
        assert A.shape[1] == K
    if b_layout == _lib.B_NK:
        Bm = _strided(b, N, K, ldb if ldb is not None else K).float()
    else:
        Bm = _strided(b, K, N, ldb if ldb is not None else N).float().t()
    r = alpha * (A @ Bm.t())
    if bias is not None:
        assert bias.dtype == torch.float32
        r = r + bias[:N]
    if bias_rows is not None:
        r = r + bias_rows.reshape(-1, N).repeat_interleave(rows_per_bias, 0)[:M]
    if residual is not None:
        assert residual.dtype == BF16
        r = r + residual.float()
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype)
    if accumulate:
        assert out.dtype == torch.float32
        out += r
    else:
        out.copy_(r.to(out.dtype))
    return out


def conv3x3_nhwc(x, w_packed, *, x2=None, taps=real_ops.TAPS_3X3, n_out_img=None, **epi):
    nbuf, H, W, C1 = x.shape
    C2 = x2.shape[-1] if x2 is not None else 0
    n_img = n_out_img if n_out_img is not None else nbuf
    K = len(taps) * (C1 + C2)
    assert w_packed.shape[1] == K
    assert C1 % 64 == 0 and C2 % 64 == 0 and (128 % W == 0 or W % 128 == 0)
    conv = dict(n_img_buf=nbuf, H=H, W=W, Cin1=C1, Cin2=C2, taps=list(taps))
    return gemm(x, w_packed, n_img * H * W, w_packed.shape[0], K, a_layout=_lib.A_CONV, conv=conv, a2=x2, **epi)


def _heads(t, B, L, h, d):
    return t.float().reshape(B, L, h, d).transpose(1, 2)


def attn_fwd(q, k, v, B, heads, Lq, Lk, scale=None, head_dim=64, out=None):
    _launches[0] += 1
    assert head_dim == 64
    scale = head_dim ** -0.5 if scale is None else scale
    qf, kf, vf = _heads(q, B, Lq, heads, 64), _heads(k, B, Lk, heads, 64), _heads(v, B, Lk, heads, 64)
    s = (qf @ kf.transpose(-1, -2)) * scale
    o = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B * Lq, heads * 64)
    if out is None:
        out = torch.empty((B * Lq, heads * 64), dtype=BF16)
    out.copy_(o.to(BF16))
    Lp = (Lq + 127) // 128 * 128
    lse = torch.zeros((B, heads, Lp))
    lse[:, :, :Lq] = torch.logsumexp(s, -1)
    return out, lse.reshape(-1)


def attn_bwd(q, k, v, o, dout, lse, B, heads, Lq, Lk, scale=None, head_dim=64, dq=None, dk=None, dv=None):
    _launches[0] += 1
    scale = head_dim ** -0.5 if scale is None else scale
    qf = _heads(q, B, Lq, heads, 64).detach().requires_grad_(True)
    kf = _heads(k, B, Lk, heads, 64).detach().requires_grad_(True)
    vf = _heads(v, B, Lk, heads, 64).detach().requires_grad_(True)
    with torch.enable_grad():
        oo = (torch.softmax((qf @ kf.transpose(-1, -2)) * scale, -1) @ vf).transpose(1, 2).reshape(B * Lq, heads * 64)
        oo.backward(dout.float())
    C = heads * 64
    back = lambda t, L: t.grad.transpose(1, 2).reshape(B * L, C).to(BF16)
    dq = torch.empty((B * Lq, C), dtype=BF16) if dq is None else dq
    dk = torch.empty((B * Lk, C), dtype=BF16) if dk is None else dk
    dv = torch.empty((B * Lk, C), dtype=BF16) if dv is None else dv
    dq.copy_(back(qf, Lq)); dk.copy_(back(kf, Lk)); dv.copy_(back(vf, Lk))
    return dq, dk, dv


def groupnorm_fwd(x, N, HW, C, G, eps, gamma, beta, silu):
    _launches[0] += 1
    xf = x.float().reshape(N, HW, C).permute(0, 2, 1)
    y = F.group_norm(xf, G, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    xg = x.float().reshape(N, HW, G, C // G)
    mean = xg.mean(dim=(1, 3))
    var = xg.var(dim=(1, 3), unbiased=False)
    stats = torch.stack([mean, torch.rsqrt(var + eps)], dim=-1)
    return y.permute(0, 2, 1).reshape(N * HW, C).to(BF16), stats


def groupnorm_bwd(x, dy, N, HW, C, G, gamma, beta, stats, silu, dres=None, dgamma=None, dbeta=None):
    _launches[0] += 1
    assert dy.is_contiguous() and (dres is None or dres.is_contiguous())
    cpg = C // G
    xg = x.float().reshape(N, HW, G, cpg)
    mean, rstd = stats[:, None, :, 0:1], stats[:, None, :, 1:2]
    xh = (xg - mean) * rstd
    ga, be = gamma.reshape(1, 1, G, cpg), beta.reshape(1, 1, G, cpg)
    dz = dy.float().reshape(N, HW, G, cpg)
    if silu:
        z = xh * ga + be
        sg = torch.sigmoid(z)
        dz = dz * (sg * (1 + z * (1 - sg)))
    gd = ga * dz
    A = gd.mean(dim=(1, 3), keepdim=True)
    Bm = (gd * xh).mean(dim=(1, 3), keepdim=True)
    dx = (rstd * (gd - A - xh * Bm)).reshape(N * HW, C)
    if dres is not None:
        dx = dx + dres.float()
    if dgamma is not None:
        dgamma += (dz * xh).sum(dim=(0, 1)).reshape(C)
    if dbeta is not None:
        dbeta += dz.sum(dim=(0, 1)).reshape(C)
    return dx.to(BF16)


def layernorm_fwd(x, gamma, beta, eps=1e-5, mod_scale=None, mod_shift=None, rows_per_mod=1, want_stats=True):
    _launches[0] += 1
    assert mod_scale is None
    xf = x.float()
    y = F.layer_norm(xf, (x.shape[1],), gamma, beta, eps)
    stats = torch.stack([xf.mean(1), torch.rsqrt(xf.var(1, unbiased=False) + eps)], dim=-1)
    return y.to(BF16), (stats if want_stats else None)


def layernorm_bwd(x, dy, gamma, stats, dres=None, dgamma=None, dbeta=None, accumulate=True):
    _launches[0] += 1
    assert dy.is_contiguous() and (dres is None or dres.is_contiguous())
    xf = x.float()
    xh = (xf - stats[:, :1]) * stats[:, 1:]
    gg = dy.float() * gamma
    dx = stats[:, 1:] * (gg - gg.mean(1, keepdim=True) - xh * (gg * xh).mean(1, keepdim=True))
    if dres is not None:
        dx = dx + dres.float()
    if dgamma is not None:
        v = (dy.float() * xh).sum(0)
        dgamma.copy_(dgamma + v if accumulate else v)
    if dbeta is not None:
        v = dy.float().sum(0)
        dbeta.copy_(dbeta + v if accumulate else v)
    return dx.to(BF16)


def geglu_fwd(x):
    _launches[0] += 1
    h, g = x.float().chunk(2, dim=-1)
    return (h * F.gelu(g)).to(BF16)


def geglu_bwd(x, dout):
    _launches[0] += 1
    xr = x.float().detach().requires_grad_(True)
    with torch.enable_grad():
        h, g = xr.chunk(2, dim=-1)
        (h * F.gelu(g)).backward(dout.float())
    return xr.grad.to(BF16)


def elementwise(x, a, mode, out=None):
    _launches[0] += 1
    assert x.is_contiguous() and (a is None or a.is_contiguous())
    xf = x.float()
    if mode == 0:
        r = F.silu(xf)
    elif mode == 1:
        af = a.float()
        s = torch.sigmoid(af)
        r = xf * (s * (1 + af * (1 - s)))
    elif mode == 2:
        r = xf + a.float()
    else:
        r = xf
    if out is None:
        return r.to(BF16)
    out.copy_(r.to(BF16))
    return out


def nchw_to_nhwc(x, cpad):
    _launches[0] += 1
    N, C, H, W = x.shape
    out = torch.zeros((N * H * W, cpad), dtype=BF16)
    out[:, :C] = x.permute(0, 2, 3, 1).reshape(-1, C).to(BF16)
    return out


def nhwc_to_nchw(x, N, Cc, H, W):
    _launches[0] += 1
    return x[:, :Cc].float().reshape(N, H, W, Cc).permute(0, 3, 1, 2).contiguous()


def upsample2x(x, N, H, W, C, backward=False):
    _launches[0] += 1
    if not backward:
        return x.reshape(N, H, 1, W, 1, C).expand(N, H, 2, W, 2, C).reshape(N * 4 * H * W, C).contiguous()
    return x.float().reshape(N, H, 2, W, 2, C).sum(dim=(2, 4)).reshape(N * H * W, C).to(BF16)


def phase_split2(x, N, H, W, C, inverse=False):
    _launches[0] += 1
    if not inverse:
        return x.reshape(N, H // 2, 2, W // 2, 2, C).permute(2, 4, 0, 1, 3, 5).reshape(N * H * W, C).contiguous()
    return x.reshape(2, 2, N, H // 2, W // 2, C).permute(2, 3, 0, 4, 1, 5).reshape(N * H * W, C).contiguous()


def colsum(x, out=None, accumulate=False):
    _launches[0] += 1
    v = x.float().sum(0)
    if out is None:
        return v
    out.copy_(out + v if accumulate else v)
    return out


def im2col3x3(x, N, H, W, C, stride=1):
    """[N*H*W, C] bf16 NHWC -> [N*Ho*Wo, 9*C] bf16, column = tap * C + c, tap = ky * 3 + kx, zero padding 1."""
    _launches[0] += 1
    img = x.reshape(N, H, W, C).permute(0, 3, 1, 2).float()
    cols = F.unfold(img, kernel_size=3, padding=1, stride=stride)            # [N, C*9, Ho*Wo], row = c * 9 + tap
    L = cols.shape[-1]
    cols = cols.view(N, C, 9, L).permute(0, 3, 2, 1).reshape(N * L, 9 * C)   # -> column = tap * C + c
    return cols.to(BF16)


def conv_pack(W, ci_p, co_p, cod_p, fwd, dgrad):
    """fwd[co, t*ci_p + ci] = W[co, ci, t]; dgrad[ci, t*cod_p + co] = W[co, ci, taps-1-t] (flipped kernel); zero padding."""
    _launches[0] += 1
    Co, Ci, kh, kw = W.shape
    Wp = torch.zeros((co_p, kh, kw, ci_p), dtype=torch.float32)
    Wp[:Co, :, :, :Ci] = W.permute(0, 2, 3, 1)
    fwd.copy_(Wp.reshape(co_p, kh * kw * ci_p))
    if dgrad is not None:
        Wd = torch.zeros((Ci, kh, kw, cod_p), dtype=torch.float32)
        Wd[:, :, :, :Co] = W.flip(2, 3).permute(1, 2, 3, 0)
        dgrad.copy_(Wd.reshape(Ci, kh * kw * cod_p))


def conv_wgrad_unpack(G, Co, Ci, Ci_pad, taps, wgrad, accumulate=True):
    """wgrad[co, ci, tap] (+)= G[co, tap * Ci_pad + ci]"""
    _launches[0] += 1
    g = G[:Co].reshape(Co, taps, Ci_pad)[:, :, :Ci].permute(0, 2, 1).reshape(wgrad.shape)
    wgrad.copy_(wgrad + g if accumulate else g)


def colsum_groups(x, groups, rows, out=None, accumulate=False):
    _launches[0] += 1
    v = x[: groups * rows].float().reshape(groups, rows, -1).sum(1)
    if out is None:
        return v
    out.copy_(out + v if accumulate else v)
    return out


def fold_lokr(W, w1, w2, dst, multiplier=1.0):
    _launches[0] += 1
    assert dst.is_contiguous() and dst.shape == W.shape
    r = W if w1 is None else W + torch.kron(w1, w2) * multiplier
    dst.copy_(r.to(BF16))
    return dst


def fold_lora(W, up, down, scale, dst):
    _launches[0] += 1
    dst.copy_((W + (up @ down) * scale).to(BF16))
    return dst


def axpy_f32(a, b, alpha, out):
    _launches[0] += 1
    out.copy_(a + alpha * b)
    return out


def lokr_grad(G, w1, w2, dw1, dw2, multiplier=1.0):
    _launches[0] += 1
    (ol, im), (ok, inn) = w1.shape, w2.shape
    G4 = G.reshape(ol, ok, im, inn)
    dw1 += torch.einsum("lkin,kn->li", G4, w2) * multiplier
    dw2 += torch.einsum("lkin,li->kn", G4, w1) * multiplier


def lora_grad(G, up, down, scale, dup, ddown):
    _launches[0] += 1
    dup += (G @ down.t()) * scale
    ddown += (up.t() @ G) * scale


def copy2d(src, dst):
    _launches[0] += 1
    dst.copy_(src.to(BF16))
    return dst


def sincos_embed(vals, dim, flip_sin_to_cos=True):
    _launches[0] += 1
    vals = vals.reshape(-1).float()
    half = dim // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    a = vals[:, None] * f[None, :]
    s, c = torch.sin(a), torch.cos(a)
    return (torch.cat([c, s], 1) if flip_sin_to_cos else torch.cat([s, c], 1)).to(BF16)


def _workspace(nfloats, device, tag="ws"):
    return torch.empty((max(nfloats, 1),), dtype=torch.float32)


def noise_fwd(x0, tables, *, target_type, pred_type, use_snr_weight, use_debiased, gamma, eps=None, timesteps=None, seed=0,
              offset=0, temb_dim=0, want_eps=True, sigmas=None, edm_sigma_data=0.0, step_dev=None):
    _launches[0] += 1
    from oracle import loss_oracle, philox

    B = x0.shape[0]
    n_per = x0[0].numel()
    T = tables["acp"].numel()
    if timesteps is None:
        timesteps = torch.from_numpy(philox.sample_timesteps(B, T, seed, offset))
    if eps is None:
        eps = torch.from_numpy(philox.normals(B, n_per, seed, offset)).reshape(x0.shape).to(x0.dtype)
    tab = loss_oracle.Tables(tables["acp"], tables["sigma_t"], tables["snr"])
    if sigmas is not None:  # rectified-flow time sampling: sigmas given, timesteps are placeholders
        sg = sigmas.to(x0).view(-1, *([1] * (x0.dim() - 1)))
        x_t = (x0 + eps.to(x0.dtype) * sg) * (1 / (sg ** 2 + 1) ** 0.5)
        tgt = loss_oracle.target(x0, eps.to(x0.dtype), timesteps, tab, target_type)
        return x_t, tgt, (eps if want_eps else None), timesteps, sigmas.float(), torch.ones((2, B)), None
    x_t = loss_oracle.noisy_latents(x0, eps.to(x0.dtype), timesteps, tab)
    tgt = loss_oracle.target(x0, eps.to(x0.dtype), timesteps, tab, target_type)
    w = loss_oracle.loss_weights(timesteps, tab, use_snr_weight=use_snr_weight, use_debiased=use_debiased, gamma=gamma,
                                 prediction_type=pred_type)
    if edm_sigma_data > 0:
        w = loss_oracle.edm_weight(timesteps, tab, edm_sigma_data)
    temb = sincos_embed(timesteps, temb_dim) if temb_dim else None
    return x_t, tgt, (eps if want_eps else None), timesteps, tab.sigma_t[timesteps], w, temb


def wmse_fwd(pred, target, w):
    _launches[0] += 1
    losses = ((pred.float() - target.float()) ** 2).flatten(1).mean(1)
    if w is not None:
        losses = w[1] * (losses * w[0])
    return losses.mean(), losses


def wmse_bwd(pred, target, w, grad=None, grad_scale=1.0, out_dtype=torch.float32):
    _launches[0] += 1
    B = pred.shape[0]
    coef = torch.full((B,), 2.0 * grad_scale / (pred[0].numel() * B))
    if w is not None:
        coef = coef * w[0] * w[1]
    if grad is not None:
        coef = coef * grad.float()
    return ((pred.float() - target.float()) * coef.view(-1, *([1] * (pred.dim() - 1)))).to(out_dtype)


def pred_convert(out, x, sigma, t, acp, pred_type, target_type, backward=False):
    """uwu_pred_convert: pred = A_b * out + C_b * x (get_prediction_for_training); backward maps d(pred) -> d(out) = A_b * g.
    The per-sample coefficients are read off the oracle's conversion by probing it with (1, 0) and (0, 1)."""
    _launches[0] += 1
    from oracle import loss_oracle

    B = out.shape[0]
    tab = loss_oracle.Tables(acp, None, None)
    one, zero = torch.ones((B, 1)), torch.zeros((B, 1))

    def conv(o, xx):
        x0, eps = loss_oracle.x0_eps_from_pred(xx, o, sigma.float(), pred_type)
        return loss_oracle.target(x0, eps, t, tab, target_type)

    A, Cc = conv(one, zero), conv(zero, one)
    shape = (B,) + (1,) * (out.dim() - 1)
    if backward:
        return out.float() * A.view(shape)
    return out.float() * A.view(shape) + x.float() * Cc.view(shape)


def _req_cuda(*ts):
    pass


def geglu_fusable(M, F, backward=False):
    return False  # the fused GEGLU epilogues exist only in the CUDA GEMM kernel


def launch_count():
    return _launches[0]


PATCHED = ["gemm", "conv3x3_nhwc", "attn_fwd", "attn_bwd", "groupnorm_fwd", "groupnorm_bwd", "layernorm_fwd", "layernorm_bwd",
           "geglu_fwd", "geglu_bwd", "elementwise", "nchw_to_nhwc", "nhwc_to_nchw", "upsample2x", "phase_split2", "colsum",
           "fold_lokr", "fold_lora", "axpy_f32", "lokr_grad", "lora_grad", "copy2d", "sincos_embed", "_workspace", "noise_fwd",
           "wmse_fwd", "wmse_bwd", "pred_convert", "geglu_fusable", "launch_count", "_req_cuda", "im2col3x3", "conv_wgrad_unpack", "colsum_groups", "conv_pack"]


def install(monkeypatch):
    import sys

    this = sys.modules[__name__]
    for name in PATCHED:
        monkeypatch.setattr(real_ops, name, getattr(this, name))
