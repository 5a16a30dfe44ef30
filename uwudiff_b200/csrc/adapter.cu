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

// out = a + alpha * b (fp32 vectors; norm deltas gamma + w_norm * multiplier)
__global__ void axpy_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fadd_rn(a[i], __fmul_rn(b[i], alpha));
}

// ------------------------------------------------------------------------------------------------
// adapter gradients from G [N, K] fp32 (ldg)
// ------------------------------------------------------------------------------------------------
// dw1[l,i] += mult * sum_{k,n} G[l*ok+k, i*in+n] w2[k,n]
// grid (ol*im, row_splits): block (b, s) reduces rows [s*rows_per, ...) of the [ok x in] tile of G against w2.
// VEC: in_n % 4 == 0 and 16-byte aligned rows -> float4 loads.
template <bool VEC>
__global__ void __launch_bounds__(256) lokr_dw1_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ w2,
                                                       int ok, int in_n, int im, int rows_per, float mult,
                                                       float* __restrict__ dw1) {
    const int l = blockIdx.x / im, i = blockIdx.x - l * im;
    const int k0 = blockIdx.y * rows_per;
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
        if (gridDim.y > 1)
            atomicAdd(&dw1[blockIdx.x], s * mult);
        else
            dw1[blockIdx.x] += s * mult;
    }
}
// dw2[k,n] += mult * sum_{l,i} G[l*ok+k, i*in+n] w1[l,i]
// grid (ceil(ok*in/VW/128), l_splits): one thread per (k, VW consecutive n), looping over its slice of l and all i.
template <int VW>
__global__ void __launch_bounds__(128) lokr_dw2_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ w1,
                                                       int ol, int ok, int im, int in_n, int l_per, float mult,
                                                       float* __restrict__ dw2) {
    const int nv = in_n / VW;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ok * nv) return;
    const int k = e / nv, n = (e - k * nv) * VW;
    const int l0 = blockIdx.y * l_per, l1 = min(ol, l0 + l_per);
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
    if (gridDim.y > 1) {
#pragma unroll
        for (int j = 0; j < VW; ++j) atomicAdd(o + j, acc[j] * mult);
    } else {
#pragma unroll
        for (int j = 0; j < VW; ++j) o[j] += acc[j] * mult;
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
// ddown[q, k] += s * sum_o up[o,q] G[o,k] : one thread per (q, k)
__global__ void lora_ddown_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ up, int N, int K, int r,
                                  float s, float* __restrict__ ddown) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r * K) return;
    const int q = e / K, k = e - q * K;
    float acc = 0.f;
    for (int o = 0; o < N; ++o) acc = fmaf(up[o * r + q], G[(size_t)o * ldg + k], acc);
    ddown[e] += acc * s;
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
                                                       float bc1, float bc2_sqrt, const float* __restrict__ clip) {
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
    const int target = 4 * sm_count();
    {
        // enough blocks to fill the machine: split each (l, i) tile over row ranges (>= 8 rows each)
        const int tiles = out_l * in_m;
        int splits = (target + tiles - 1) / tiles;
        if (splits > (out_k + 7) / 8) splits = (out_k + 7) / 8;
        if (splits < 1) splits = 1;
        const int rows_per = (out_k + splits - 1) / splits;
        splits = (out_k + rows_per - 1) / rows_per;
        dim3 grid(tiles, splits);
        if (vec)
            lokr_dw1_kernel<true><<<grid, 256, 0, stream>>>(G, ldg, w2, out_k, in_n, in_m, rows_per, multiplier, dw1);
        else
            lokr_dw1_kernel<false><<<grid, 256, 0, stream>>>(G, ldg, w2, out_k, in_n, in_m, rows_per, multiplier, dw1);
        UWU_CHECK_LAUNCH();
    }
    {
        const int vw = vec ? 4 : 1;
        const int blocks = (out_k * (in_n / vw) + 127) / 128;
        int lsplits = (target + blocks - 1) / blocks;
        if (lsplits > out_l) lsplits = out_l;
        if (lsplits < 1) lsplits = 1;
        const int l_per = (out_l + lsplits - 1) / lsplits;
        lsplits = (out_l + l_per - 1) / l_per;
        dim3 grid(blocks, lsplits);
        if (vec)
            lokr_dw2_kernel<4><<<grid, 128, 0, stream>>>(G, ldg, w1, out_l, out_k, in_m, in_n, l_per, multiplier, dw2);
        else
            lokr_dw2_kernel<1><<<grid, 128, 0, stream>>>(G, ldg, w1, out_l, out_k, in_m, in_n, l_per, multiplier, dw2);
        UWU_CHECK_LAUNCH();
    }
    return UWU_OK;
}

extern "C" int uwu_lora_grad(const float* G, int64_t ldg, const float* up, const float* down, int32_t N, int32_t K, int32_t r,
                             float scale, float* dup, float* ddown, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(G && up && down && dup && ddown && N > 0 && K > 0 && r > 0 && ldg >= K, "uwu_lora_grad: bad arguments");
    lora_dup_kernel<<<(N * r + 7) / 8, 256, 0, stream>>>(G, ldg, down, N, K, r, scale, dup);
    UWU_CHECK_LAUNCH();
    lora_ddown_kernel<<<(r * K + 127) / 128, 128, 0, stream>>>(G, ldg, up, N, K, r, scale, ddown);
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
                            int64_t step, const float* norm_clip, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(p_ptrs && g_ptrs && m_ptrs && v_ptrs && numels && chunk_tensor && chunk_index && n_chunks > 0 && step > 0,
                  "uwu_mt_adamw: bad arguments");
    MTTables t{p_ptrs, g_ptrs, m_ptrs, v_ptrs, numels, chunk_tensor, chunk_index, chunk_elems};
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    mt_adamw_kernel<<<n_chunks, 256, 0, stream>>>(t, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2),
                                                  norm_clip);
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
