"""ctypes binding of libuwu_b200.so (the C ABI declared in include/uwu_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuwu_b200.so")

UWU_F32, UWU_BF16 = 0, 1
TARGET_CODES = {"epsilon": 0, "v_prediction": 1, "sample": 2, "rectified_flow": 3}
WEIGHT_MIN_SNR, WEIGHT_DEBIASED, WEIGHT_EDM = 1, 2, 4
A_ROW, A_COL, A_CONV = 0, 1, 2
B_NK, B_KN = 0, 1


class UwuError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a2", C.c_void_p), ("b", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("a_layout", C.c_int32), ("b_layout", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64),
        ("n_img_buf", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cin1", C.c_int32), ("Cin2", C.c_int32), ("ntaps", C.c_int32),
        ("tap_dn", C.c_int32 * 9), ("tap_dh", C.c_int32 * 9), ("tap_dw", C.c_int32 * 9),
        ("out", C.c_void_p), ("out2", C.c_void_p),
        ("ldo", C.c_int64), ("ldo2", C.c_int64),
        ("n_split", C.c_int32), ("out_dtype", C.c_int32),
        ("bias", C.c_void_p), ("bias_rows", C.c_void_p), ("rows_per_bias", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("alpha", C.c_float), ("accumulate", C.c_int32), ("block_n", C.c_int32), ("stream_k", C.c_int32),
        ("k_segs", C.c_int32), ("a_seg_off", C.c_int32), ("b_seg_off", C.c_int32), ("grp_n", C.c_int32), ("a_grp_koff", C.c_int32),
        ("dbg_a_lbo", C.c_int32), ("dbg_a_sbo", C.c_int32), ("dbg_a_kadv", C.c_int32),
        ("dbg_b_lbo", C.c_int32), ("dbg_b_sbo", C.c_int32), ("dbg_b_kadv", C.c_int32),
        ("epi_mode", C.c_int32), ("aux", C.c_void_p), ("ld_aux", C.c_int64),
    ]


class NoiseDesc(C.Structure):
    _fields_ = [
        ("x0", C.c_void_p), ("eps_in", C.c_void_p), ("t_in", C.c_void_p),
        ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("acp", C.c_void_p), ("sigma_t", C.c_void_p), ("snr", C.c_void_p),
        ("T", C.c_int32), ("B", C.c_int32), ("n_per", C.c_int64),
        ("dtype", C.c_int32), ("target_type", C.c_int32), ("pred_type", C.c_int32),
        ("weight_flags", C.c_int32), ("gamma", C.c_float),
        ("x_t", C.c_void_p), ("target", C.c_void_p), ("eps_out", C.c_void_p),
        ("t_out", C.c_void_p), ("sigma_out", C.c_void_p), ("w_out", C.c_void_p),
        ("temb_out", C.c_void_p), ("temb_dim", C.c_int32), ("sigma_in", C.c_void_p),
        ("sigma_data", C.c_float), ("step_dev", C.c_void_p),
    ]


class FoldEntry(C.Structure):
    _fields_ = [("W", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p), ("dst", C.c_void_p), ("kind", C.c_int32),
                ("N", C.c_int32), ("K", C.c_int32), ("p0", C.c_int32), ("p1", C.c_int32), ("p2", C.c_int32),
                ("scale", C.c_float), ("chunk0", C.c_int32)]


class LokrGradEntry(C.Structure):
    _fields_ = [("G", C.c_void_p), ("w1", C.c_void_p), ("w2", C.c_void_p), ("dw1", C.c_void_p), ("dw2", C.c_void_p),
                ("ldg", C.c_int64), ("out_l", C.c_int32), ("out_k", C.c_int32), ("in_m", C.c_int32), ("in_n", C.c_int32),
                ("multiplier", C.c_float), ("vec", C.c_int32), ("target", C.c_int32), ("block0", C.c_int32)]


_lib = None

# name -> (restype, argtypes); kept in one table so tests can check every header symbol is exported
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "uwu_last_error": (C.c_char_p, []),
    "uwu_version": (C.c_int, []),
    "uwu_launch_count": (C.c_int64, []),
    "uwu_gemm": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "uwu_noise_fwd": (C.c_int, [C.POINTER(NoiseDesc), _P]),
    "uwu_sincos_embed": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "uwu_wmse_workspace_floats": (C.c_int64, [_I32, _I64]),
    "uwu_wmse_fwd": (C.c_int, [_P, _I32, _P, _I32, _I32, _I64, _P, _P, _P, _P, _P]),
    "uwu_wmse_bwd": (C.c_int, [_P, _I32, _P, _I32, _I32, _I64, _P, _P, _F, _P, _I32, _P]),
    "uwu_pred_convert": (C.c_int, [_P, _P, _I32, _P, _P, _P, _I32, _I64, _I32, _I32, _I32, _P, _P]),
    "uwu_attn_lse_floats": (C.c_int64, [_I32, _I32, _I32]),
    "uwu_attn_fwd": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I64, _I64, _I64, _I64, _F, _P]),
    "uwu_attn_fwd_masked": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I64, _I64, _I64, _I64, _F, _I32, _P, _P]),
    "uwu_attn_bwd_workspace_floats": (C.c_int64, [_I32, _I32, _I32]),
    "uwu_attn_bwd": (C.c_int, [_P] * 9 + [_I32] * 5 + [_I64] * 8 + [_F, _P, _P]),
    "uwu_groupnorm_workspace_floats": (C.c_int64, [_I32, _I32, _I32, _I32]),
    "uwu_groupnorm_fwd": (C.c_int, [_P, _I32, _I32, _I32, _I32, _F, _P, _P, _I32, _P, _P, _P, _P]),
    "uwu_groupnorm_bwd": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "uwu_layernorm_fwd": (C.c_int, [_P, _I32, _I32, _F, _P, _P, _P, _P, _I32, _P, _P, _P]),
    "uwu_layernorm_bwd_workspace_floats": (C.c_int64, [_I32, _I32]),
    "uwu_layernorm_bwd": (C.c_int, [_P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _P, _P]),
    "uwu_softmax_rows": (C.c_int, [_P, _I64, _I32, _I64, _P]),
    "uwu_geglu_fwd": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "uwu_geglu_bwd": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "uwu_elementwise": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "uwu_adaln_fwd": (C.c_int, [_P, _I64, _I32, _F, _P, _I64, _I32, _I32, _I32, _P, _P, _P]),
    "uwu_adaln_bwd": (C.c_int, [_P, _P, _I64, _I32, _P, _I64, _I32, _P, _I32, _P, _P, _P, _I64, _I32, _I32, _P]),
    "uwu_gate_residual_fwd": (C.c_int, [_P, _P, _I64, _I32, _P, _I64, _I32, _I32, _P, _P]),
    "uwu_gate_residual_bwd": (C.c_int, [_P, _P, _I64, _I32, _P, _I64, _I32, _I32, _P, _P, _I64, _I32, _P]),
    "uwu_patchify": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _I64, _P]),
    "uwu_unpatchify": (C.c_int, [_P, _I32, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "uwu_embed_gather": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "uwu_embed_scatter_add": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "uwu_nchw_to_nhwc": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "uwu_nhwc_to_nchw": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I64, _P, _P]),
    "uwu_upsample2x": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "uwu_phase_split2": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "uwu_colsum_workspace_floats": (C.c_int64, [_I64, _I32]),
    "uwu_colsum_bf16": (C.c_int, [_P, _I64, _I32, _I64, _I32, _P, _P, _P]),
    "uwu_lokr_grad_plan_blocks": (C.c_int32, [_I32, _I32, _I32, _I32, _I32, _I32]),
    "uwu_lokr_grad_batch": (C.c_int, [_P, _I32, _I32, _P]),
    "uwu_conv_pack": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "uwu_timestep_hist": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "uwu_fold_loha": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _F, _P, _P]),
    "uwu_loha_grad": (C.c_int, [_P, _I64, _P, _P, _P, _P, _I32, _I32, _I32, _F, _P, _P, _P, _P, _P]),
    "uwu_fold_lokr": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _F, _P, _P]),
    "uwu_fold_lora": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _F, _P, _P]),
    "uwu_axpy_f32": (C.c_int, [_P, _P, _F, _I32, _P, _P]),
    "uwu_lokr_grad": (C.c_int, [_P, _I64, _P, _P, _I32, _I32, _I32, _I32, _F, _P, _P, _P]),
    "uwu_lora_grad": (C.c_int, [_P, _I64, _P, _P, _I32, _I32, _I32, _F, _P, _P, _P]),
    "uwu_im2col3x3": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "uwu_conv_wgrad_unpack": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "uwu_colsum_groups_bf16": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P]),
    "uwu_fold_batch": (C.c_int, [_P, _P, _I32, _I32, _P]),
    "uwu_lokr_z": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _I32, _P, _I32, _P]),
    "uwu_lokr_dw1": (C.c_int, [_P, _P, _I64, _I64, _I32, _I32, _I32, _F, _P, _P]),
    "uwu_groupnorm_fwd_fused": (C.c_int, [_P, _I32, _I32, _I32, _I32, _F, _P, _P, _I32, _P, _P, _P, _P, _P]),
    "uwu_groupnorm_bwd_fused": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P]),
    "uwu_lokr_fused_supported": (C.c_int, [_I32, _I32, _I32, _I32]),
    "uwu_lokr_fused_grad": (C.c_int, [_P, _I64, _P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _F, _P]),
    "uwu_mt_gradnorm": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _F, _P, _P, _P]),
    "uwu_mt_adamw": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _F, _F, _F, _F, _F, _I64, _P, _P, _P]),
    "uwu_copy2d_bf16": (C.c_int, [_P, _I32, _I64, _P, _I64, _I64, _I32, _P]),
}


def lib() -> C.CDLL:
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UwuError(
                f"{LIB_PATH} not found: build it with `python -m uwudiff_b200.build` "
                "(there is no CPU fallback for the uwudiff_b200 kernels)"
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().uwu_last_error().decode("utf-8", "replace")
        if rc == -2:
            # unsupported target/prediction type: same exception type as the reference (ValueError)
            raise ValueError(msg)
        raise UwuError(f"{what} failed (rc={rc}): {msg}")
