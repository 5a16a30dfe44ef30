"""Tensor-level wrappers over the C ABI (include/uwu_b200.h).

PyTorch is used only for device memory and streams: every function here takes CUDA tensors, passes raw
pointers + the current stream to libuwu_b200.so and returns tensors it allocated for the outputs.
No function falls back to a PyTorch/CPU implementation.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (A_COL, A_CONV, A_ROW, B_KN, B_NK, UWU_BF16, UWU_F32, GemmDesc, NoiseDesc, check, lib)

_DT = {torch.float32: UWU_F32, torch.bfloat16: UWU_BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.UwuError("uwudiff_b200 kernels need CUDA tensors (no CPU fallback)")


def launch_count() -> int:
    return int(lib().uwu_launch_count())


# --------------------------------------------------------------------------------------------------
# GEMM / conv
# --------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, *, a_layout: int = A_ROW, b_layout: int = B_NK,
         lda: Optional[int] = None, ldb: Optional[int] = None, out: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.bfloat16, bias: Optional[torch.Tensor] = None,
         bias_rows: Optional[torch.Tensor] = None, rows_per_bias: int = 1, residual: Optional[torch.Tensor] = None,
         alpha: float = 1.0, accumulate: bool = False, block_n: int = 0, out2: Optional[torch.Tensor] = None,
         n_split: int = 0, conv: Optional[dict] = None, a2: Optional[torch.Tensor] = None,
         dbg: Optional[dict] = None) -> torch.Tensor:
    """out[M,N] = alpha * A·Bᵀ + bias + bias_rows + residual (see uwu_gemm in include/uwu_b200.h)."""
    _req_cuda(a, b, out, bias, bias_rows, residual, out2, a2)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, "GEMM operands must be bf16"
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    d = GemmDesc()
    d.a, d.a2, d.b = _ptr(a), _ptr(a2), _ptr(b)
    d.M, d.N, d.K = M, N, K
    d.a_layout, d.b_layout = a_layout, b_layout
    if a_layout == A_ROW:
        d.lda = lda if lda is not None else K
    elif a_layout == A_COL:
        d.lda = lda if lda is not None else M
    d.ldb = ldb if ldb is not None else (K if b_layout == B_NK else N)
    if conv is not None:
        d.n_img_buf, d.H, d.W = conv["n_img_buf"], conv["H"], conv["W"]
        d.Cin1, d.Cin2 = conv["Cin1"], conv.get("Cin2", 0)
        taps = conv["taps"]
        d.ntaps = len(taps)
        for i, (dn, dh, dw) in enumerate(taps):
            d.tap_dn[i], d.tap_dh[i], d.tap_dw[i] = dn, dh, dw
    d.out, d.out2 = _ptr(out), _ptr(out2)
    d.ldo = out.stride(0) if out.dim() == 2 else N
    d.ldo2 = out2.stride(0) if out2 is not None else 0
    d.n_split = n_split
    d.out_dtype = _DT[out.dtype]
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    if bias_rows is not None:
        assert bias_rows.dtype == torch.float32
    d.bias, d.bias_rows, d.rows_per_bias = _ptr(bias), _ptr(bias_rows), rows_per_bias
    if residual is not None:
        assert residual.dtype == torch.bfloat16
        d.residual, d.ldr = _ptr(residual), residual.stride(0) if residual.dim() == 2 else N
    d.alpha, d.accumulate, d.block_n = alpha, int(accumulate), block_n
    if dbg:
        for k, v in dbg.items():
            setattr(d, "dbg_" + k, v)
    check(lib().uwu_gemm(C.byref(d), _stream()), "uwu_gemm")
    return out


TAPS_3X3 = [(0, dy - 1, dx - 1) for dy in range(3) for dx in range(3)]


def conv3x3_nhwc(x: torch.Tensor, w_packed: torch.Tensor, *, x2: Optional[torch.Tensor] = None,
                 taps: Sequence = TAPS_3X3, n_out_img: Optional[int] = None, **epi) -> torch.Tensor:
    """Implicit-GEMM convolution over NHWC bf16 input(s).

    x: [Nbuf,H,W,C1] (x2: [Nbuf,H,W,C2] channel-concatenated after x); w_packed: [Cout, ntaps*(C1+C2)] bf16 with
    K index = tap*(C1+C2) + c.  Returns [n_out_img*H*W, Cout].
    """
    nbuf, H, W, C1 = x.shape
    C2 = x2.shape[-1] if x2 is not None else 0
    n_img = n_out_img if n_out_img is not None else nbuf
    Cout = w_packed.shape[0]
    K = len(taps) * (C1 + C2)
    assert w_packed.shape[1] == K, (w_packed.shape, K)
    conv = dict(n_img_buf=nbuf, H=H, W=W, Cin1=C1, Cin2=C2, taps=list(taps))
    return gemm(x, w_packed, n_img * H * W, Cout, K, a_layout=A_CONV, b_layout=B_NK, conv=conv, a2=x2, **epi)


# --------------------------------------------------------------------------------------------------
# noising / loss
# --------------------------------------------------------------------------------------------------
def noise_fwd(x0: torch.Tensor, tables: dict, *, target_type: str, pred_type: str, use_snr_weight: bool,
              use_debiased: bool, gamma: float, eps: Optional[torch.Tensor] = None,
              timesteps: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0, temb_dim: int = 0,
              want_eps: bool = True):
    """One launch: (x_t, target, eps, t, sigma, w[2,B], temb). See uwu_noise_fwd."""
    _req_cuda(x0, eps, timesteps)
    if target_type not in _lib.TARGET_CODES:
        raise ValueError(f"Unsupported target type {target_type}")
    x0 = x0.contiguous()
    B = x0.shape[0]
    n_per = x0.numel() // B if B > 0 else 0
    dev = x0.device
    x_t = torch.empty_like(x0)
    target = torch.empty_like(x0)
    eps_out = torch.empty_like(x0) if want_eps else None
    t_out = torch.empty((B,), device=dev, dtype=torch.int64)
    sigma = torch.empty((B,), device=dev, dtype=torch.float32)
    w = torch.empty((2, B), device=dev, dtype=torch.float32)
    temb = torch.empty((B, temb_dim), device=dev, dtype=torch.bfloat16) if temb_dim else None
    d = NoiseDesc()
    d.x0 = _ptr(x0)
    if eps is not None:
        eps = eps.to(x0.dtype).contiguous()
        assert eps.shape == x0.shape
        d.eps_in = _ptr(eps)
    if timesteps is not None:
        timesteps = timesteps.to(device=dev, dtype=torch.int64).contiguous()
        d.t_in = _ptr(timesteps)
    d.seed, d.offset = seed & (2**64 - 1), offset & (2**64 - 1)
    d.acp, d.sigma_t, d.snr = _ptr(tables["acp"]), _ptr(tables["sigma_t"]), _ptr(tables["snr"])
    d.T, d.B, d.n_per = tables["acp"].numel(), B, n_per
    d.dtype = _DT[x0.dtype]
    d.target_type = _lib.TARGET_CODES[target_type]
    d.pred_type = _lib.TARGET_CODES.get(pred_type, 0)
    d.weight_flags = (_lib.WEIGHT_MIN_SNR if use_snr_weight else 0) | (_lib.WEIGHT_DEBIASED if use_debiased else 0)
    d.gamma = gamma
    d.x_t, d.target, d.eps_out = _ptr(x_t), _ptr(target), _ptr(eps_out)
    d.t_out, d.sigma_out, d.w_out = _ptr(t_out), _ptr(sigma), _ptr(w)
    d.temb_out, d.temb_dim = _ptr(temb), temb_dim
    check(lib().uwu_noise_fwd(C.byref(d), _stream()), "uwu_noise_fwd")
    return x_t, target, eps_out, t_out, sigma, w, temb


def sincos_embed(vals: torch.Tensor, dim: int, flip_sin_to_cos: bool = True) -> torch.Tensor:
    _req_cuda(vals)
    vals = vals.reshape(-1).to(torch.float32).contiguous()
    out = torch.empty((vals.numel(), dim), device=vals.device, dtype=torch.bfloat16)
    check(lib().uwu_sincos_embed(_ptr(vals), vals.numel(), dim, int(flip_sin_to_cos), _ptr(out), _stream()),
          "uwu_sincos_embed")
    return out


def wmse_fwd(pred: torch.Tensor, target: torch.Tensor, w: Optional[torch.Tensor]):
    """(loss scalar tensor, losses[B]) — fused per-sample MSE * weights -> batch mean."""
    _req_cuda(pred, target, w)
    pred, target = pred.contiguous(), target.contiguous()
    B = pred.shape[0]
    n_per = pred.numel() // max(B, 1)
    ws = torch.empty((max(int(lib().uwu_wmse_workspace_floats(B, n_per)), 1),), device=pred.device, dtype=torch.float32)
    losses = torch.empty((B,), device=pred.device, dtype=torch.float32)
    loss = torch.empty((), device=pred.device, dtype=torch.float32)
    check(lib().uwu_wmse_fwd(_ptr(pred), _DT[pred.dtype], _ptr(target), _DT[target.dtype], B, n_per, _ptr(w), _ptr(ws),
                             _ptr(losses), _ptr(loss), _stream()), "uwu_wmse_fwd")
    return loss, losses


def wmse_bwd(pred: torch.Tensor, target: torch.Tensor, w: Optional[torch.Tensor], grad: Optional[torch.Tensor] = None,
             grad_scale: float = 1.0, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    _req_cuda(pred, target, w, grad)
    pred, target = pred.contiguous(), target.contiguous()
    B = pred.shape[0]
    n_per = pred.numel() // B
    dpred = torch.empty(pred.shape, device=pred.device, dtype=out_dtype)
    if grad is not None:
        grad = grad.to(torch.float32).contiguous()
    check(lib().uwu_wmse_bwd(_ptr(pred), _DT[pred.dtype], _ptr(target), _DT[target.dtype], B, n_per, _ptr(w), _ptr(grad),
                             grad_scale, _ptr(dpred), _DT[out_dtype], _stream()), "uwu_wmse_bwd")
    return dpred
