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
    torch.cuda.synchronize()


if __name__ == "__main__":
    main(sys.argv[1])
