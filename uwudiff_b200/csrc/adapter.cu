// LyCORIS adapter bookkeeping kernels and the multi-tensor optimizer (all bandwidth-bound, fp32 parameters).
//
//   fold   : W'_bf16 = bf16(W_fp32 + dW * multiplier)   with dW = kron(w1, w2) (LoKr, full_matrix) or up @ down * alpha/r (LoRA)
//            -> the adapter delta rides inside the base GEMM's B operand: forward and dgrad are ONE tcgen05 GEMM each.
//   grads  : given G = dY^T X (fp32, produced by the tcgen05 GEMM in its "wgrad form"), contract it into the adapter
//            factors:  dw1[l,i] = sum_{k,n} G[l*ok+k, i*in+n] w2[k,n],  dw2[k,n] = sum_{l,i} G[l*ok+k, i*in+n] w1[l,i]
//                      dup[o,r] = s sum_k G[o,k] down[r,k],             ddown[r,k] = s sum_o up[o,r] G[o,k]
//   adamw  : torch.optim.AdamW semantics over a table of tensors, with Lightning's clip_grad_norm_ folded in
//            (global L2 norm -> clip coefficient on the device, no host sync).
//
// Replaces lycoris-lora's `F.linear(x, W + make_kron(w1, w2))` forward patch and its autograd-through-kron backward
// (applied at src/duwu/trainer/trainer.py:152-154), torch.optim.AdamW (configs/demo_training_lycoris.yaml:50,
// src/duwu/trainer/trainer.py:52-74) and Lightning's gradient_clip_val (configs/demo_training_lycoris.yaml:13).
#include "api_internal.h"
#include "common.cuh"

namespace uwu {

// ------------------------------------------------------------------------------------------------
// fold
// ------------------------------------------------------------------------------------------------
__global__ void fold_lokr_kernel(const float* __restrict__ W, const float* __restrict__ w1, const float* __restrict__ w2, int N,
                                 int K, int ok, int in_n, int im, float mult, __nv_bfloat16* __restrict__ dst) {
    const long long total = (long long)N * K / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        const int n = (int)(e / K), k = (int)(e - (long long)n * K);
        const float4 w = *reinterpret_cast<const float4*>(W + e);
        float v[4] = {w.x, w.y, w.z, w.w};
        if (w1 != nullptr) {
            const int l = n / ok, kk = n - l * ok;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = k + j;
                const int ii = c / in_n, nn = c - ii * in_n;
                v[j] = __fadd_rn(v[j], __fmul_rn(__fmul_rn(w1[l * im + ii], w2[kk * in_n + nn]), mult));
            }
        }
        uint2 u;
        u.x = pack_bf16(v[0], v[1]);
        u.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2*>(dst + e) = u;
    }
}

__global__ void fold_lora_kernel(const float* __restrict__ W, const float* __restrict__ up, const float* __restrict__ down, int N,
                                 int K, int r, float s, __nv_bfloat16* __restrict__ dst) {
    const long long total = (long long)N * K / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        const int n = (int)(e / K), k = (int)(e - (long long)n * K);
        const float4 w = *reinterpret_cast<const float4*>(W + e);
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < r; ++q) {
            const float u = up[n * r + q];
            const float4 dn = *reinterpret_cast<const float4*>(down + (size_t)q * K + k);
            d[0] = fmaf(u, dn.x, d[0]); d[1] = fmaf(u, dn.y, d[1]); d[2] = fmaf(u, dn.z, d[2]); d[3] = fmaf(u, dn.w, d[3]);
        }
        uint2 o;
        o.x = pack_bf16(w.x + d[0] * s, w.y + d[1] * s);
        o.y = pack_bf16(w.z + d[2] * s, w.w + d[3] * s);
        *reinterpret_cast<uint2*>(dst + e) = o;
    }
}

// LoHa delta for 4 consecutive columns of row n: (sum_q w1a[n,q] w1b[q,k]) * (sum_q w2a[n,q] w2b[q,k])
__device__ __forceinline__ void loha_delta4(const float* __restrict__ w1a, const float* __restrict__ w1b,
                                            const float* __restrict__ w2a, const float* __restrict__ w2b, int n, int k, int K, int r,
                                            float (&d)[4]) {
    float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < r; ++q) {
        const float u1 = w1a[n * r + q], u2 = w2a[n * r + q];
        const float4 b1 = *reinterpret_cast<const float4*>(w1b + (size_t)q * K + k);
        const float4 b2 = *reinterpret_cast<const float4*>(w2b + (size_t)q * K + k);
        p1[0] = fmaf(u1, b1.x, p1[0]); p1[1] = fmaf(u1, b1.y, p1[1]); p1[2] = fmaf(u1, b1.z, p1[2]); p1[3] = fmaf(u1, b1.w, p1[3]);
        p2[0] = fmaf(u2, b2.x, p2[0]); p2[1] = fmaf(u2, b2.y, p2[1]); p2[2] = fmaf(u2, b2.z, p2[2]); p2[3] = fmaf(u2, b2.w, p2[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = p1[j] * p2[j];
}

__global__ void fold_loha_kernel(const float* __restrict__ W, const float* __restrict__ w1a, const float* __restrict__ w1b,
                                 const float* __restrict__ w2a, const float* __restrict__ w2b, int N, int K, int r, float s,
                                 __nv_bfloat16* __restrict__ dst) {
    const long long total = (long long)N * K / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        const int n = (int)(e / K), k = (int)(e - (long long)n * K);
        const float4 w = *reinterpret_cast<const float4*>(W + e);
        float d[4];
        loha_delta4(w1a, w1b, w2a, w2b, n, k, K, r, d);
        uint2 o;
        o.x = pack_bf16(w.x + d[0] * s, w.y + d[1] * s);
        o.y = pack_bf16(w.z + d[2] * s, w.w + d[3] * s);
        *reinterpret_cast<uint2*>(dst + e) = o;
    }
}

// LoHa gradients from G = dL/dW_eff [N, K]: dP1 = s G o P2, dP2 = s G o P1 with P1 = w1a w1b, P2 = w2a w2b (rank r <= 16)
//   rows kernel: one warp per row n      -> dw1a[n, :] += dP1[n, :] w1b^T,  dw2a[n, :] += dP2[n, :] w2b^T
//   cols kernel: one thread per column k -> dw1b[:, k] += w1a^T dP1[:, k],  dw2b[:, k] += w2a^T dP2[:, k]  (rows split over
//                blockIdx.y, partial sums added atomically)
constexpr int LOHA_MAXR = 16;
__global__ void __launch_bounds__(256) loha_grad_rows_kernel(const float* __restrict__ G, long long ldg,
                                                             const float* __restrict__ w1a, const float* __restrict__ w1b,
                                                             const float* __restrict__ w2a, const float* __restrict__ w2b, int N,
                                                             int K, int r, float s, float* __restrict__ dw1a,
                                                             float* __restrict__ dw2a) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float u1[LOHA_MAXR], u2[LOHA_MAXR], a1[LOHA_MAXR], a2[LOHA_MAXR];
#pragma unroll
    for (int q = 0; q < LOHA_MAXR; ++q) {
        u1[q] = q < r ? w1a[n * r + q] : 0.f;
        u2[q] = q < r ? w2a[n * r + q] : 0.f;
        a1[q] = a2[q] = 0.f;
    }
    for (int k = lane; k < K; k += 32) {
        float p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int q = 0; q < LOHA_MAXR; ++q)
            if (q < r) {
                p1 = fmaf(u1[q], w1b[(size_t)q * K + k], p1);
                p2 = fmaf(u2[q], w2b[(size_t)q * K + k], p2);
            }
        const float g = G[(size_t)n * ldg + k] * s;
        const float d1 = g * p2, d2 = g * p1;
#pragma unroll
        for (int q = 0; q < LOHA_MAXR; ++q)
            if (q < r) {
                a1[q] = fmaf(d1, w1b[(size_t)q * K + k], a1[q]);
                a2[q] = fmaf(d2, w2b[(size_t)q * K + k], a2[q]);
            }
    }
#pragma unroll
    for (int q = 0; q < LOHA_MAXR; ++q)
        if (q < r) {
            const float s1 = warp_sum(a1[q]), s2 = warp_sum(a2[q]);
            if (lane == 0) {
                dw1a[n * r + q] += s1;
                dw2a[n * r + q] += s2;
            }
        }
}
__global__ void __launch_bounds__(128) loha_grad_cols_kernel(const float* __restrict__ G, long long ldg,
                                                             const float* __restrict__ w1a, const float* __restrict__ w1b,
                                                             const float* __restrict__ w2a, const float* __restrict__ w2b, int N,
                                                             int K, int r, int rows_per, float s, float* __restrict__ dw1b,
                                                             float* __restrict__ dw2b) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int n0 = blockIdx.y * rows_per, n1 = min(N, n0 + rows_per);
    float b1[LOHA_MAXR], b2[LOHA_MAXR], a1[LOHA_MAXR], a2[LOHA_MAXR];
#pragma unroll
    for (int q = 0; q < LOHA_MAXR; ++q) {
        b1[q] = q < r ? w1b[(size_t)q * K + k] : 0.f;
        b2[q] = q < r ? w2b[(size_t)q * K + k] : 0.f;
        a1[q] = a2[q] = 0.f;
    }
    for (int n = n0; n < n1; ++n) {
        float p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int q = 0; q < LOHA_MAXR; ++q)
            if (q < r) {
                p1 = fmaf(w1a[n * r + q], b1[q], p1);
                p2 = fmaf(w2a[n * r + q], b2[q], p2);
            }
        const float g = G[(size_t)n * ldg + k] * s;
        const float d1 = g * p2, d2 = g * p1;
#pragma unroll
        for (int q = 0; q < LOHA_MAXR; ++q)
            if (q < r) {
                a1[q] = fmaf(w1a[n * r + q], d1, a1[q]);
                a2[q] = fmaf(w2a[n * r + q], d2, a2[q]);
            }
    }
#pragma unroll
    for (int q = 0; q < LOHA_MAXR; ++q)
        if (q < r) {
            atomicAdd(&dw1b[(size_t)q * K + k], a1[q]);
            atomicAdd(&dw2b[(size_t)q * K + k], a2[q]);
        }
}

// All adapter folds of one step in ONE launch: block -> (entry, chunk) through a table built once on the host.
__global__ void __launch_bounds__(256) fold_batch_kernel(const uwu_fold_entry* __restrict__ entries,
                                                         const int32_t* __restrict__ chunk_entry, int chunk_elems) {
    const uwu_fold_entry e = entries[chunk_entry[blockIdx.x]];
    const long long total = (long long)e.N * e.K;
    const long long beg = (long long)(blockIdx.x - e.chunk0) * chunk_elems;
    const long long end = min(total, beg + chunk_elems);
    const int K = e.K;
    if (e.kind == 3) {  // fp32 dst = W + scale * a   (norm deltas)
        float* dst = reinterpret_cast<float*>(e.dst);
        for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) dst[i] = __fadd_rn(e.W[i], __fmul_rn(e.a[i], e.scale));
        return;
    }
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.dst);
    // 32-bit index arithmetic (N * K < 2^31 for every layer), row / column advanced incrementally instead of divided
    const int ibeg = (int)beg, iend = (int)end;
    int i4 = ibeg + threadIdx.x * 4;
    int n = i4 / K, k = i4 - n * K;
    const int step = blockDim.x * 4, dn = step / K, dk = step - dn * K;
    for (; i4 < iend; i4 += step, n += dn, k += dk) {
        if (k >= K) {
            k -= K;
            ++n;
        }
        const float4 w = *reinterpret_cast<const float4*>(e.W + i4);
        float v[4] = {w.x, w.y, w.z, w.w};
        if (e.kind == 1) {  // LoKr: + kron(w1, w2)[n, k] * scale, same rounding sequence as fold_lokr_kernel
            const int ok = e.p0, in_n = e.p1, im = e.p2;
            const int l = n / ok, kk = n - l * ok;
            if ((in_n & 3) == 0) {  // the 4 columns share one w1 element and read 4 consecutive w2 elements
                const int ii = k / in_n, nn = k - ii * in_n;
                const float w1v = e.a[l * im + ii];
                const float4 w2v = *reinterpret_cast<const float4*>(e.b + kk * in_n + nn);
                v[0] = __fadd_rn(v[0], __fmul_rn(__fmul_rn(w1v, w2v.x), e.scale));
                v[1] = __fadd_rn(v[1], __fmul_rn(__fmul_rn(w1v, w2v.y), e.scale));
                v[2] = __fadd_rn(v[2], __fmul_rn(__fmul_rn(w1v, w2v.z), e.scale));
                v[3] = __fadd_rn(v[3], __fmul_rn(__fmul_rn(w1v, w2v.w), e.scale));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = k + j;
                    const int ii = c / in_n, nn = c - ii * in_n;
                    v[j] = __fadd_rn(v[j], __fmul_rn(__fmul_rn(e.a[l * im + ii], e.b[kk * in_n + nn]), e.scale));
                }
            }
        } else if (e.kind == 2) {  // LoRA: + (up @ down)[n, k] * scale
            const int r = e.p0;
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            for (int q = 0; q < r; ++q) {
                const float u = e.a[n * r + q];
                const float4 dn = *reinterpret_cast<const float4*>(e.b + (size_t)q * K + k);
                d[0] = fmaf(u, dn.x, d[0]); d[1] = fmaf(u, dn.y, d[1]); d[2] = fmaf(u, dn.z, d[2]); d[3] = fmaf(u, dn.w, d[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = v[j] + d[j] * e.scale;
        } else if (e.kind == 4) {  // LoHa: + (w1a @ w1b) o (w2a @ w2b) * scale; w2a / w2b = a + p1 / a + p2 (one flat buffer)
            float d[4];
            loha_delta4(e.a, e.b, e.a + e.p1, e.a + e.p2, n, k, K, e.p0, d);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = v[j] + d[j] * e.scale;
        }
        uint2 u;
        u.x = pack_bf16(v[0], v[1]);
        u.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2*>(dst + i4) = u;
    }
}

// out = a + alpha * b (fp32 vectors; norm deltas gamma + w_norm * multiplier)
__global__ void axpy_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fadd_rn(a[i], __fmul_rn(b[i], alpha));
}

// ------------------------------------------------------------------------------------------------
// adapter gradients from G [N, K] fp32 (ldg)
// ------------------------------------------------------------------------------------------------
// dw1[l,i] += mult * sum_{k,n} G[l*ok+k, i*in+n] w2[k,n]  and  dw2[k,n] += mult * sum_{l,i} G[l*ok+k, i*in+n] w1[l,i]
// in ONE launch (the per-layer contractions are latency bound, so two launches cost twice):
//   blocks [0, n1): (tile = (l, i), split s) reduce rows [s*rows_per, ...) of the [ok x in] tile of G against w2 -> dw1
//   blocks [n1, ..): one thread per (k, VW consecutive n), looping over its slice of l and all i            -> dw2
// VEC: in_n % 4 == 0 and 16-byte aligned rows -> float4 loads.
template <bool VEC>
__device__ __forceinline__ void lokr_grad_body(const int bid, const float* __restrict__ G, long long ldg,
                                               const float* __restrict__ w1, const float* __restrict__ w2, int ol, int ok, int im,
                                               int in_n, int rows_per, int splits, int l_per, int lsplits, int blocks2, int n1,
                                               float mult, float* __restrict__ dw1, float* __restrict__ dw2) {
    if (bid < n1) {
        const int tile = bid / splits, sp = bid - tile * splits;
        const int l = tile / im, i = tile - l * im;
        const int k0 = sp * rows_per;
        const int k1 = min(ok, k0 + rows_per);
        const float* g = G + (size_t)l * ok * ldg + (size_t)i * in_n;
        float acc = 0.f;
        if (VEC) {
            const int nv = in_n >> 2;
            const int total = (k1 - k0) * nv;
            for (int e = threadIdx.x; e < total; e += blockDim.x) {
                const int k = k0 + e / nv, n = (e - (e / nv) * nv) << 2;
                const float4 a = *reinterpret_cast<const float4*>(g + (size_t)k * ldg + n);
                const float4 b = *reinterpret_cast<const float4*>(w2 + (size_t)k * in_n + n);
                acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
            }
        } else {
            const int total = (k1 - k0) * in_n;
            for (int e = threadIdx.x; e < total; e += blockDim.x) {
                const int k = k0 + e / in_n, n = e - (e / in_n) * in_n;
                acc = fmaf(g[(size_t)k * ldg + n], w2[(size_t)k * in_n + n], acc);
            }
        }
        __shared__ float red[8];
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
            if (splits > 1)
                atomicAdd(&dw1[tile], s * mult);
            else
                dw1[tile] += s * mult;
        }
        return;
    }
    constexpr int VW = VEC ? 4 : 1;
    const int b2 = bid - n1;
    const int bx = b2 % blocks2, by = b2 / blocks2;
    const int nv = in_n / VW;
    const int e = bx * blockDim.x + threadIdx.x;
    if (e >= ok * nv) return;
    const int k = e / nv, n = (e - k * nv) * VW;
    const int l0 = by * l_per, l1 = min(ol, l0 + l_per);
    float acc[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) acc[j] = 0.f;
    for (int l = l0; l < l1; ++l) {
        const float* g = G + ((size_t)l * ok + k) * ldg + n;
#pragma unroll 4
        for (int i = 0; i < im; ++i) {
            const float w = __ldg(w1 + l * im + i);
            if (VW == 4) {
                const float4 a = *reinterpret_cast<const float4*>(g + (size_t)i * in_n);
                acc[0] = fmaf(a.x, w, acc[0]); acc[1 % VW] = fmaf(a.y, w, acc[1 % VW]);
                acc[2 % VW] = fmaf(a.z, w, acc[2 % VW]); acc[3 % VW] = fmaf(a.w, w, acc[3 % VW]);
            } else {
                acc[0] = fmaf(g[(size_t)i * in_n], w, acc[0]);
            }
        }
    }
    float* o = dw2 + (size_t)k * in_n + n;
    if (lsplits > 1) {
#pragma unroll
        for (int j = 0; j < VW; ++j) atomicAdd(o + j, acc[j] * mult);
    } else {
#pragma unroll
        for (int j = 0; j < VW; ++j) o[j] += acc[j] * mult;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256) lokr_grad_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ w1,
                                                        const float* __restrict__ w2, int ol, int ok, int im, int in_n, int rows_per,
                                                        int splits, int l_per, int lsplits, int blocks2, int n1, float mult,
                                                        float* __restrict__ dw1, float* __restrict__ dw2) {
    pdl_trigger();
    lokr_grad_body<VEC>((int)blockIdx.x, G, ldg, w1, w2, ol, ok, im, in_n, rows_per, splits, l_per, lsplits, blocks2, n1, mult, dw1,
                        dw2);
}

// Grid decomposition of one adapter's contraction for a budget of about `target` blocks (shared by host plan and kernels)
struct LokrPlan {
    int rows_per, splits, l_per, lsplits, blocks2, n1, n2, vec;
};
__host__ __device__ inline LokrPlan lokr_plan(int ol, int ok, int im, int in_n, int vec, int target) {
    LokrPlan p;
    const int tiles = ol * im;
    int splits = (target + tiles - 1) / tiles;
    if (splits > (ok + 7) / 8) splits = (ok + 7) / 8;
    if (splits < 1) splits = 1;
    p.rows_per = (ok + splits - 1) / splits;
    p.splits = (ok + p.rows_per - 1) / p.rows_per;
    const int vw = vec ? 4 : 1;
    p.blocks2 = (ok * (in_n / vw) + 255) / 256;
    int lsplits = (target + p.blocks2 - 1) / p.blocks2;
    if (lsplits > ol) lsplits = ol;
    if (lsplits < 1) lsplits = 1;
    p.l_per = (ol + lsplits - 1) / lsplits;
    p.lsplits = (ol + p.l_per - 1) / p.l_per;
    p.n1 = tiles * p.splits;
    p.n2 = p.blocks2 * p.lsplits;
    p.vec = vec;
    return p;
}

// The contractions of MANY adapters in one launch (their G = dY^T X matrices were kept, one buffer per adapter): the
// per-layer launches are latency bound (13 us each, 630 per SDXL step); batched, the pass runs at memory bandwidth.
// block -> entry by binary search over the block prefix sums.
__global__ void __launch_bounds__(256) lokr_grad_batch_kernel(const uwu_lokr_grad_entry* __restrict__ entries, int n_entries) {
    pdl_trigger();
    int lo = 0, hi = n_entries - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (entries[mid].block0 <= (int)blockIdx.x) lo = mid;
        else hi = mid - 1;
    }
    const uwu_lokr_grad_entry e = entries[lo];
    const LokrPlan p = lokr_plan(e.out_l, e.out_k, e.in_m, e.in_n, e.vec, e.target);
    const int bid = (int)blockIdx.x - e.block0;
    if (bid >= p.n1 + p.n2) return;
    // accumulation into dw1 / dw2 is always atomic here: other entries never alias, but `+=` needs a unique writer per
    // element, which the split counts (> 1 => atomics inside the body) already guarantee
    if (e.vec)
        lokr_grad_body<true>(bid, e.G, e.ldg, e.w1, e.w2, e.out_l, e.out_k, e.in_m, e.in_n, p.rows_per, p.splits, p.l_per, p.lsplits,
                             p.blocks2, p.n1, e.multiplier, e.dw1, e.dw2);
    else
        lokr_grad_body<false>(bid, e.G, e.ldg, e.w1, e.w2, e.out_l, e.out_k, e.in_m, e.in_n, p.rows_per, p.splits, p.l_per, p.lsplits,
                              p.blocks2, p.n1, e.multiplier, e.dw1, e.dw2);
}

// ------------------------------------------------------------------------------------------------
// factored LoKr gradients (no G = dY^T X): with X [M, im, in], dY [M, ol, ok], w1 [ol, im], w2 [ok, in]
//   Z[m, l, n] = sum_i w1[l, i] X[m, i, n]            (lokr_z_kernel, bandwidth / FMA bound)
//   dw2[k, n] += sum_{m, l} dY[m, l, k] Z[m, l, n]      (tcgen05 GEMM, segmented reduction over l, stream-K)
//   V[m, l, n] = sum_k dY[m, l, k] w2[k, n]            (tcgen05 GEMM, grouped N)
//   dw1[l, i] += sum_{m, n} V[m, l, n] X[m, i, n]       (lokr_dw1_mma_kernel, warp-level mma.sync, bandwidth bound)
// FLOPs: 2 * (2 M ol ok in) instead of 2 M (ol ok)(im in) for the full weight gradient, i.e. 2/im of it.
// ------------------------------------------------------------------------------------------------
// thread = (row m, 8 consecutive n); l processed in blocks of LB so the accumulators stay in registers
template <int LB>
__global__ void __launch_bounds__(256) lokr_z_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ w1,
                                                     long long M, int ol, int im, int in_n, __nv_bfloat16* __restrict__ z, int w1_t) {
    pdl_trigger();
    extern __shared__ float sw1[];  // [ol][im]  (w1_t: w1 is stored [im][ol], i.e. the mixing matrix is its transpose)
    for (int i = threadIdx.x; i < ol * im; i += blockDim.x) sw1[i] = w1_t ? w1[(i % im) * ol + i / im] : w1[i];
    __syncthreads();
    const int nv = in_n >> 3;
    const long long total = M * nv;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long m = t / nv;
        const int n = (int)(t - m * nv) << 3;
        const __nv_bfloat16* xr = x + m * ldx + n;
        __nv_bfloat16* zr = z + m * ((long long)ol * in_n) + n;
        for (int l0 = 0; l0 < ol; l0 += LB) {
            float acc[LB][8];
#pragma unroll
            for (int l = 0; l < LB; ++l)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[l][j] = 0.f;
            for (int i = 0; i < im; ++i) {
                const uint4 u = *reinterpret_cast<const uint4*>(xr + (long long)i * in_n);
                const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
                const float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    const float w = (l0 + l < ol) ? sw1[(l0 + l) * im + i] : 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[l][j] = fmaf(w, f[j], acc[l][j]);
                }
            }
#pragma unroll
            for (int l = 0; l < LB; ++l) {
                if (l0 + l < ol) {
                    uint4 o;
                    o.x = pack_bf16(acc[l][0], acc[l][1]);
                    o.y = pack_bf16(acc[l][2], acc[l][3]);
                    o.z = pack_bf16(acc[l][4], acc[l][5]);
                    o.w = pack_bf16(acc[l][6], acc[l][7]);
                    *reinterpret_cast<uint4*>(zr + (long long)(l0 + l) * in_n) = o;
                }
            }
        }
    }
}

UWU_DEVINL void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// dw1[l, i] += mult * sum_{m, n} V[m, l, n] X[m, i, n].  Each warp owns rows m = w, w + W, ... and accumulates the
// [LT*16 x IT*8] output in mma fragments across all of them (the (m, n) pairs ARE the reduction dimension).  Operand
// fragments come straight from global memory as 16-byte loads: lane (g = lane/4, t = lane%4) reads elements 8t..8t+7 of
// a 32-wide n chunk of row g; the same permutation of the reduction index is used for both operands, so the two
// m16n8k16 products per chunk sum exactly the 32 products.
template <int LT, int IT>
__global__ void __launch_bounds__(256) lokr_dw1_mma_kernel(const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ x,
                                                           long long ldx, long long M, int ol, int im, int in_n, float mult,
                                                           float* __restrict__ dw1) {
    pdl_trigger();
    __shared__ float red[LT * 16 * IT * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    float c[LT][IT][4];
#pragma unroll
    for (int a = 0; a < LT; ++a)
#pragma unroll
        for (int b = 0; b < IT; ++b)
#pragma unroll
            for (int j = 0; j < 4; ++j) c[a][b][j] = 0.f;
    const long long ldv = (long long)ol * in_n;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (long long m = (long long)blockIdx.x * 8 + warp; m < M; m += (long long)gridDim.x * 8) {
        const __nv_bfloat16* vr = v + m * ldv;
        const __nv_bfloat16* xr = x + m * ldx;
        for (int n0 = 0; n0 < in_n; n0 += 32) {
            const int n = n0 + t * 8;
            const bool nok = n < in_n;
            uint4 av[LT][2], bv[IT];
#pragma unroll
            for (int a = 0; a < LT; ++a) {
                const int l0 = a * 16 + g, l1 = l0 + 8;
                av[a][0] = (nok && l0 < ol) ? *reinterpret_cast<const uint4*>(vr + (long long)l0 * in_n + n) : zero;
                av[a][1] = (nok && l1 < ol) ? *reinterpret_cast<const uint4*>(vr + (long long)l1 * in_n + n) : zero;
            }
#pragma unroll
            for (int b = 0; b < IT; ++b) {
                const int i = b * 8 + g;
                bv[b] = (nok && i < im) ? *reinterpret_cast<const uint4*>(xr + (long long)i * in_n + n) : zero;
            }
#pragma unroll
            for (int a = 0; a < LT; ++a)
#pragma unroll
                for (int b = 0; b < IT; ++b) {
                    mma_bf16_16816(c[a][b], av[a][0].x, av[a][1].x, av[a][0].y, av[a][1].y, bv[b].x, bv[b].y);
                    mma_bf16_16816(c[a][b], av[a][0].z, av[a][1].z, av[a][0].w, av[a][1].w, bv[b].z, bv[b].w);
                }
        }
    }
    // block reduction of the 8 warps' fragments, then one atomic per output element
    for (int i = threadIdx.x; i < LT * 16 * IT * 8; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < LT; ++a)
#pragma unroll
        for (int b = 0; b < IT; ++b) {
            const int r0 = a * 16 + g, c0 = b * 8 + t * 2;
            atomicAdd(&red[r0 * (IT * 8) + c0], c[a][b][0]);
            atomicAdd(&red[r0 * (IT * 8) + c0 + 1], c[a][b][1]);
            atomicAdd(&red[(r0 + 8) * (IT * 8) + c0], c[a][b][2]);
            atomicAdd(&red[(r0 + 8) * (IT * 8) + c0 + 1], c[a][b][3]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < LT * 16 * IT * 8; i += blockDim.x) {
        const int l = i / (IT * 8), ii = i - l * (IT * 8);
        if (l < ol && ii < im) atomicAdd(&dw1[l * im + ii], red[i] * mult);
    }
}

// dup[o, q] += s * sum_k G[o,k] down[q,k] : one warp per (o, q)
__global__ void __launch_bounds__(256) lora_dup_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ down,
                                                       int N, int K, int r, float s, float* __restrict__ dup) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= N * r) return;
    const int o = w / r, q = w - o * r;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(G[(size_t)o * ldg + k], down[(size_t)q * K + k], acc);
    acc = warp_sum(acc);
    if (lane == 0) dup[w] += acc * s;
}
// ddown[q, k] += s * sum_o up[o,q] G[o,k] : one thread per column k (all r ranks, r <= 16), rows split over blockIdx.y
// (partial sums meet through atomics: the single-block-per-column version took 119 us at 1280 x 1280)
__global__ void __launch_bounds__(128) lora_ddown_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ up,
                                                         int N, int K, int r, int rows_per, float s, float* __restrict__ ddown) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int o0 = blockIdx.y * rows_per, o1 = min(N, o0 + rows_per);
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
    for (int o = o0; o < o1; ++o) {
        const float g = G[(size_t)o * ldg + k];
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (q < r) acc[q] = fmaf(up[o * r + q], g, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q)
        if (q < r) atomicAdd(&ddown[(size_t)q * K + k], acc[q] * s);
}

// ------------------------------------------------------------------------------------------------
// multi-tensor grad norm + AdamW
// ------------------------------------------------------------------------------------------------
struct MTTables {
    const uint64_t* p;
    const uint64_t* g;
    const uint64_t* m;
    const uint64_t* v;
    const int64_t* numel;
    const int32_t* chunk_tensor;
    const int32_t* chunk_index;
    int chunk_elems;
};

__global__ void __launch_bounds__(256) mt_sqnorm_kernel(MTTables t, float* __restrict__ partial) {
    const int ti = t.chunk_tensor[blockIdx.x];
    const long long n = t.numel[ti];
    const long long beg = (long long)t.chunk_index[blockIdx.x] * t.chunk_elems;
    const long long end = min(n, beg + t.chunk_elems);
    const float* g = reinterpret_cast<const float*>(t.g[ti]);
    float acc = 0.f;
    for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) acc = fmaf(g[i], g[i], acc);
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}
// out[0] = ||g||_2, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) (1 when max_norm <= 0)
__global__ void __launch_bounds__(256) mt_norm_final_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                            float* __restrict__ out) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += (double)partial[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float norm = (float)sqrt(sh[0]);
        out[0] = norm;
        float c = 1.0f;
        if (max_norm > 0.f) c = fminf(1.0f, max_norm / (norm + 1e-6f));
        out[1] = c;
    }
}

__global__ void __launch_bounds__(256) mt_adamw_kernel(MTTables t, float lr, float beta1, float beta2, float eps, float wd,
                                                       float bc1, float bc2_sqrt, const float* __restrict__ clip,
                                                       const float* __restrict__ hyper) {
    if (hyper) {  // per-step scalars read from device memory: the launch can live in a CUDA graph while lr / step advance
        lr = hyper[0];
        bc1 = hyper[1];
        bc2_sqrt = hyper[2];
    }
    const int ti = t.chunk_tensor[blockIdx.x];
    const long long n = t.numel[ti];
    const long long beg = (long long)t.chunk_index[blockIdx.x] * t.chunk_elems;
    const long long end = min(n, beg + t.chunk_elems);
    float* p = reinterpret_cast<float*>(t.p[ti]);
    const float* g = reinterpret_cast<const float*>(t.g[ti]);
    float* m = reinterpret_cast<float*>(t.m[ti]);
    float* v = reinterpret_cast<float*>(t.v[ti]);
    const float coef = clip ? clip[1] : 1.0f;
    const float step_size = lr / bc1;
    for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.0f - lr * wd);
        const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);  // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = v[i] * beta2 + (1.0f - beta2) * gi * gi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
    }
}

// strided 2-D copy / cast into bf16: dst[r, c] = src[r, c]
template <typename T>
__global__ void copy2d_kernel(const T* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd, long long rows,
                              int cols) {
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols;
        const int c = (int)(i - r * cols);
        dst[r * ldd + c] = __float2bfloat16((float)src[r * lds + c]);
    }
}
template <typename T>
__global__ void copy2d_vec_kernel(const T* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd,
                                  long long rows, int cols8) {
    pdl_trigger();
    const long long total = rows * cols8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols8;
        const int c = (int)(i - r * cols8) * 8;
        *reinterpret_cast<uint4*>(dst + r * ldd + c) = *reinterpret_cast<const uint4*>(src + r * lds + c);
    }
}

static int grid_for(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace uwu

using namespace uwu;

extern "C" int uwu_fold_lokr(const float* W, const float* w1, const float* w2, int32_t N, int32_t K, int32_t out_l, int32_t out_k,
                             int32_t in_m, int32_t in_n, float multiplier, void* dst_bf16, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(W && dst_bf16 && N > 0 && K > 0 && K % 4 == 0, "uwu_fold_lokr: bad arguments (K must be a multiple of 4)");
    if (w1 != nullptr) {
        // shape bookkeeping must be exact: (out_l*out_k, in_m*in_n) == (N, K)
        UWU_CHECK_ARG(w2 && (long long)out_l * out_k == N && (long long)in_m * in_n == K,
                      "uwu_fold_lokr: factor shapes (%d x %d) (x) (%d x %d) do not tile the %d x %d weight", out_l, in_m, out_k,
                      in_n, N, K);
    }
    fold_lokr_kernel<<<grid_for((long long)N * K / 4, 256), 256, 0, stream>>>(W, w1, w2, N, K, out_k > 0 ? out_k : 1,
                                                                              in_n > 0 ? in_n : 1, in_m, multiplier,
                                                                              reinterpret_cast<__nv_bfloat16*>(dst_bf16));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_fold_lora(const float* W, const float* up, const float* down, int32_t N, int32_t K, int32_t r, float scale,
                             void* dst_bf16, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(W && up && down && dst_bf16 && N > 0 && K > 0 && K % 4 == 0 && r > 0, "uwu_fold_lora: bad arguments");
    fold_lora_kernel<<<grid_for((long long)N * K / 4, 256), 256, 0, stream>>>(W, up, down, N, K, r, scale,
                                                                              reinterpret_cast<__nv_bfloat16*>(dst_bf16));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_axpy_f32(const float* a, const float* b, float alpha, int32_t n, float* out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(a && b && out && n >= 0, "uwu_axpy_f32: bad arguments");
    if (n == 0) return UWU_OK;
    axpy_f32_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a, b, alpha, n, out);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_lokr_grad(const float* G, int64_t ldg, const float* w1, const float* w2, int32_t out_l, int32_t out_k,
                             int32_t in_m, int32_t in_n, float multiplier, float* dw1, float* dw2, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(G && w1 && w2 && dw1 && dw2, "uwu_lokr_grad: null pointer");
    UWU_CHECK_ARG(out_l > 0 && out_k > 0 && in_m > 0 && in_n > 0 && ldg >= (int64_t)in_m * in_n, "uwu_lokr_grad: bad shape");
    const bool vec = in_n % 4 == 0 && ldg % 4 == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(w2) & 15) == 0;
    // dw1: each (l, i) tile split over row ranges (>= 8 rows each); dw2: one thread per (k, 4 n), l range split over blocks;
    // ONE launch: blocks [0, n1) do dw1, the rest dw2
    const LokrPlan p = lokr_plan(out_l, out_k, in_m, in_n, vec ? 1 : 0, 4 * sm_count());
    if (vec)
        lokr_grad_kernel<true><<<p.n1 + p.n2, 256, 0, stream>>>(G, ldg, w1, w2, out_l, out_k, in_m, in_n, p.rows_per, p.splits, p.l_per,
                                                               p.lsplits, p.blocks2, p.n1, multiplier, dw1, dw2);
    else
        lokr_grad_kernel<false><<<p.n1 + p.n2, 256, 0, stream>>>(G, ldg, w1, w2, out_l, out_k, in_m, in_n, p.rows_per, p.splits, p.l_per,
                                                                p.lsplits, p.blocks2, p.n1, multiplier, dw1, dw2);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int32_t uwu_lokr_grad_plan_blocks(int32_t out_l, int32_t out_k, int32_t in_m, int32_t in_n, int32_t vec, int32_t target) {
    if (out_l <= 0 || out_k <= 0 || in_m <= 0 || in_n <= 0 || target <= 0) return -1;
    const LokrPlan p = lokr_plan(out_l, out_k, in_m, in_n, vec, target);
    return p.n1 + p.n2;
}

extern "C" int uwu_lokr_grad_batch(const uwu_lokr_grad_entry* entries_dev, int32_t n_entries, int32_t total_blocks, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(entries_dev && n_entries > 0 && total_blocks > 0, "uwu_lokr_grad_batch: bad arguments");
    lokr_grad_batch_kernel<<<total_blocks, 256, 0, stream>>>(entries_dev, n_entries);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_fold_loha(const float* W, const float* w1a, const float* w1b, const float* w2a, const float* w2b, int32_t N,
                             int32_t K, int32_t r, float scale, void* dst_bf16, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(W && w1a && w1b && w2a && w2b && dst_bf16 && N > 0 && K > 0 && K % 4 == 0 && r > 0, "uwu_fold_loha: bad arguments");
    UWU_CHECK_ARG(((reinterpret_cast<uintptr_t>(w1b) | reinterpret_cast<uintptr_t>(w2b) | reinterpret_cast<uintptr_t>(W)) & 15) == 0,
                  "uwu_fold_loha: W / w1b / w2b must be 16-byte aligned");
    fold_loha_kernel<<<grid_for((long long)N * K / 4, 256), 256, 0, stream>>>(W, w1a, w1b, w2a, w2b, N, K, r, scale,
                                                                             reinterpret_cast<__nv_bfloat16*>(dst_bf16));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_loha_grad(const float* G, int64_t ldg, const float* w1a, const float* w1b, const float* w2a, const float* w2b,
                             int32_t N, int32_t K, int32_t r, float scale, float* dw1a, float* dw1b, float* dw2a, float* dw2b,
                             void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(G && w1a && w1b && w2a && w2b && dw1a && dw1b && dw2a && dw2b && N > 0 && K > 0 && ldg >= K,
                  "uwu_loha_grad: bad arguments");
    UWU_CHECK_ARG(r > 0 && r <= LOHA_MAXR, "uwu_loha_grad: rank %d unsupported (1..%d)", r, LOHA_MAXR);
    loha_grad_rows_kernel<<<(N + 7) / 8, 256, 0, stream>>>(G, ldg, w1a, w1b, w2a, w2b, N, K, r, scale, dw1a, dw2a);
    UWU_CHECK_LAUNCH();
    int splits = (2 * sm_count() * 128 + K - 1) / K;
    if (splits < 1) splits = 1;
    if (splits > N) splits = N;
    const int rows_per = (N + splits - 1) / splits;
    loha_grad_cols_kernel<<<dim3((K + 127) / 128, (N + rows_per - 1) / rows_per), 128, 0, stream>>>(G, ldg, w1a, w1b, w2a, w2b, N, K, r,
                                                                                                 rows_per, scale, dw1b, dw2b);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_lora_grad(const float* G, int64_t ldg, const float* up, const float* down, int32_t N, int32_t K, int32_t r,
                             float scale, float* dup, float* ddown, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(G && up && down && dup && ddown && N > 0 && K > 0 && r > 0 && ldg >= K, "uwu_lora_grad: bad arguments");
    lora_dup_kernel<<<(N * r + 7) / 8, 256, 0, stream>>>(G, ldg, down, N, K, r, scale, dup);
    UWU_CHECK_LAUNCH();
    UWU_CHECK_ARG(r <= 16, "uwu_lora_grad: rank %d > 16 unsupported", r);
    {
        int splits = (2 * sm_count() * 128 + K - 1) / K;
        if (splits < 1) splits = 1;
        if (splits > N) splits = N;
        const int rows_per = (N + splits - 1) / splits;
        lora_ddown_kernel<<<dim3((K + 127) / 128, (N + rows_per - 1) / rows_per), 128, 0, stream>>>(G, ldg, up, N, K, r, rows_per, scale,
                                                                                                ddown);
    }
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_mt_gradnorm(const uint64_t* g_ptrs, const int64_t* numels, const int32_t* chunk_tensor,
                               const int32_t* chunk_index, int32_t n_chunks, int32_t chunk_elems, float max_norm,
                               float* partial_ws, float* out2, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(g_ptrs && numels && chunk_tensor && chunk_index && partial_ws && out2 && n_chunks > 0 && chunk_elems > 0,
                  "uwu_mt_gradnorm: bad arguments");
    MTTables t{nullptr, g_ptrs, nullptr, nullptr, numels, chunk_tensor, chunk_index, chunk_elems};
    mt_sqnorm_kernel<<<n_chunks, 256, 0, stream>>>(t, partial_ws);
    UWU_CHECK_LAUNCH();
    mt_norm_final_kernel<<<1, 256, 0, stream>>>(partial_ws, n_chunks, max_norm, out2);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_mt_adamw(const uint64_t* p_ptrs, const uint64_t* g_ptrs, const uint64_t* m_ptrs, const uint64_t* v_ptrs,
                            const int64_t* numels, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t n_chunks,
                            int32_t chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay,
                            int64_t step, const float* norm_clip, const float* hyper_dev, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(p_ptrs && g_ptrs && m_ptrs && v_ptrs && numels && chunk_tensor && chunk_index && n_chunks > 0 &&
                      (step > 0 || hyper_dev),
                  "uwu_mt_adamw: bad arguments");
    if (step <= 0) step = 1;
    MTTables t{p_ptrs, g_ptrs, m_ptrs, v_ptrs, numels, chunk_tensor, chunk_index, chunk_elems};
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    mt_adamw_kernel<<<n_chunks, 256, 0, stream>>>(t, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2),
                                                  norm_clip, hyper_dev);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_copy2d_bf16(const void* src, int32_t src_dtype, int64_t lds, void* dst_bf16, int64_t ldd, int64_t rows,
                               int32_t cols, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(rows >= 0 && cols > 0 && lds >= cols && ldd >= cols, "uwu_copy2d_bf16: bad shape");
    if (rows == 0) return UWU_OK;
    UWU_CHECK_ARG(src && dst_bf16, "uwu_copy2d_bf16: null pointer");
    auto* d = reinterpret_cast<__nv_bfloat16*>(dst_bf16);
    if (src_dtype == UWU_BF16) {
        const auto* s = reinterpret_cast<const __nv_bfloat16*>(src);
        const bool vec = cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0;
        if (vec)
            copy2d_vec_kernel<__nv_bfloat16><<<grid_for(rows * (cols / 8), 256), 256, 0, stream>>>(s, lds, d, ldd, rows, cols / 8);
        else
            copy2d_kernel<__nv_bfloat16><<<grid_for(rows * cols, 256), 256, 0, stream>>>(s, lds, d, ldd, rows, cols);
    } else {
        copy2d_kernel<float><<<grid_for(rows * cols, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(src), lds, d, ldd, rows,
                                                                             cols);
    }
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_lokr_z(const void* x, int64_t ldx, const float* w1, int64_t M, int32_t out_l, int32_t in_m, int32_t in_n,
                          void* z, int32_t w1_transposed, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(x && w1 && z && M > 0, "uwu_lokr_z: bad arguments");
    UWU_CHECK_ARG(out_l > 0 && out_l <= 64 && in_m > 0 && in_m <= 64 && in_n > 0 && in_n % 8 == 0 && ldx % 8 == 0 &&
                      ldx >= (int64_t)in_m * in_n,
                  "uwu_lokr_z: unsupported factor shape ol=%d im=%d in=%d ldx=%lld", out_l, in_m, in_n, (long long)ldx);
    UWU_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(z)) & 15) == 0, "uwu_lokr_z: 16-byte alignment");
    const long long total = M * (in_n / 8);
    const int grid = grid_for(total, 256);
    const size_t sm = (size_t)out_l * in_m * sizeof(float);
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    auto* zp = reinterpret_cast<__nv_bfloat16*>(z);
    if (out_l % 5 == 0)
        lokr_z_kernel<5><<<grid, 256, sm, stream>>>(xp, ldx, w1, M, out_l, in_m, in_n, zp, w1_transposed);
    else
        lokr_z_kernel<4><<<grid, 256, sm, stream>>>(xp, ldx, w1, M, out_l, in_m, in_n, zp, w1_transposed);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_lokr_dw1(const void* v, const void* x, int64_t ldx, int64_t M, int32_t out_l, int32_t in_m, int32_t in_n,
                            float multiplier, float* dw1, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(v && x && dw1 && M > 0, "uwu_lokr_dw1: bad arguments");
    UWU_CHECK_ARG(out_l > 0 && out_l <= 32 && in_m > 0 && in_m <= 32 && in_n > 0 && in_n % 8 == 0 && ldx % 8 == 0,
                  "uwu_lokr_dw1: unsupported factor shape ol=%d im=%d in=%d (ol, im <= 32; in %% 8 == 0)", out_l, in_m, in_n);
    UWU_CHECK_ARG(((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(x)) & 15) == 0, "uwu_lokr_dw1: 16-byte alignment");
    long long blocks = (M + 7) / 8;
    const long long cap = 2ll * sm_count();
    if (blocks > cap) blocks = cap;
    const auto* vp = reinterpret_cast<const __nv_bfloat16*>(v);
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    const int LT = (out_l + 15) / 16, IT = (in_m + 7) / 8;
#define UWU_DW1(A, B) lokr_dw1_mma_kernel<A, B><<<(int)blocks, 256, 0, stream>>>(vp, xp, ldx, M, out_l, in_m, in_n, multiplier, dw1)
    if (LT == 1) {
        if (IT == 1) UWU_DW1(1, 1); else if (IT == 2) UWU_DW1(1, 2); else if (IT == 3) UWU_DW1(1, 3); else UWU_DW1(1, 4);
    } else {
        if (IT == 1) UWU_DW1(2, 1); else if (IT == 2) UWU_DW1(2, 2); else if (IT == 3) UWU_DW1(2, 3); else UWU_DW1(2, 4);
    }
#undef UWU_DW1
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_fold_batch(const uwu_fold_entry* entries_dev, const int32_t* chunk_entry_dev, int32_t n_chunks,
                              int32_t chunk_elems, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(entries_dev && chunk_entry_dev && n_chunks > 0 && chunk_elems > 0 && chunk_elems % 1024 == 0,
                  "uwu_fold_batch: bad arguments (chunk_elems must be a multiple of 1024)");
    fold_batch_kernel<<<n_chunks, 256, 0, stream>>>(entries_dev, chunk_entry_dev, chunk_elems);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
