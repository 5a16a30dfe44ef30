// Bandwidth-bound normalisation kernels over channels-last bf16 activations (fp32 statistics):
//   GroupNorm(+SiLU) forward/backward on NHWC [N, HW, C], LayerNorm(+adaLN modulate) forward/backward on [M, C].
// 16-byte vector accesses, per-thread fixed channel ownership (no atomics in the streaming loops),
// deterministic two-stage reductions through caller-provided workspaces.
//
// Replaces ATen group_norm / layer_norm (+ SiLU) under diffusers ResnetBlock2D / Transformer2DModel /
// BasicTransformerBlock [third-party, restated in oracle/unet_oracle.py]; the in-tree evidence for the block
// algebra is src/duwu/modules/rope_unet.py:288-415.
#include "api_internal.h"
#include "common.cuh"

namespace uwu {

UWU_DEVINL void ld8(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
UWU_DEVINL void st8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ================================================================================================
// GroupNorm
// ================================================================================================
// thread layout: blockDim = cv * rpi  (cv = C/8 channel vectors, rpi rows per iteration);
// thread owns channel vector (tid % cv) for rows r0 + tid / cv + k * rpi.
struct GNGeom {
    int N, HW, C, G, cv, rpi, chunks, rows_per_chunk;
};

// partial sums: ws[((n * chunks + chunk) * G + g) * 2 + {0,1}] = {sum, sumsq}
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, GNGeom g, float* __restrict__ ws) {
    pdl_trigger();
    extern __shared__ float sh[];  // [2][C]
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int v = threadIdx.x % g.cv, rr = threadIdx.x / g.cv;
    const int r0 = chunk * g.rows_per_chunk;
    const int r1 = min(g.HW, r0 + g.rows_per_chunk);
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const __nv_bfloat16* base = x + ((size_t)n * g.HW) * g.C + v * 8;
#pragma unroll 4
    for (int r = r0 + rr; r < r1; r += g.rpi) {
        float f[8];
        ld8(base + (size_t)r * g.C, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j] += f[j];
            q[j] = fmaf(f[j], f[j], q[j]);
        }
    }
    // fixed-order block reduction over the rpi row groups (deterministic: no atomics)
    {
        float* mine = sh + (size_t)rr * 2 * g.C;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mine[v * 8 + j] = s[j];
            mine[g.C + v * 8 + j] = q[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) {
        float a = sh[i];
        for (int k = 1; k < g.rpi; ++k) a += sh[(size_t)k * 2 * g.C + i];
        sh[i] = a;
    }
    __syncthreads();
    const int cpg = g.C / g.G;
    for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
            a += sh[c];
            b += sh[g.C + c];
        }
        float* o = ws + (((size_t)n * g.chunks + chunk) * g.G + gi) * 2;
        o[0] = a;
        o[1] = b;
    }
}

// finalise mean / rstd per (n, g) from the chunk partials; stats[n, g, {mean, rstd}]
__global__ void gn_finalize_kernel(const float* __restrict__ ws, GNGeom g, float eps, float* __restrict__ stats) {
    pdl_trigger();
    // one warp per (n, g): lanes stride over the chunk partials
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (idx >= g.N * g.G) return;
    const int n = idx / g.G, gi = idx - n * g.G;
    float a = 0.f, b = 0.f;
    for (int c = lane; c < g.chunks; c += 32) {
        const float2 p = *reinterpret_cast<const float2*>(ws + (((size_t)n * g.chunks + c) * g.G + gi) * 2);
        a += p.x;
        b += p.y;
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
        const float cnt = (float)g.HW * (float)(g.C / g.G);
        const float mean = a / cnt;
        const float var = fmaxf(b / cnt - mean * mean, 0.f);
        stats[idx * 2] = mean;
        stats[idx * 2 + 1] = rsqrtf(var + eps);
    }
}

template <bool kSilu>
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, GNGeom g,
                                __nv_bfloat16* __restrict__ y) {
    pdl_trigger();
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int v = threadIdx.x % g.cv, rr = threadIdx.x / g.cv;
    const int cpg = g.C / g.G;
    float sc[8], sf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = v * 8 + j;
        const int gi = c / cpg;
        const float mean = stats[((size_t)n * g.G + gi) * 2], rstd = stats[((size_t)n * g.G + gi) * 2 + 1];
        sc[j] = rstd * gamma[c];
        sf[j] = beta[c] - mean * sc[j];
    }
    const int r0 = chunk * g.rows_per_chunk;
    const int r1 = min(g.HW, r0 + g.rows_per_chunk);
    const size_t off = ((size_t)n * g.HW) * g.C + v * 8;
#pragma unroll 4
    for (int r = r0 + rr; r < r1; r += g.rpi) {
        float f[8];
        ld8(x + off + (size_t)r * g.C, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float z = fmaf(f[j], sc[j], sf[j]);
            f[j] = kSilu ? silu_f(z) : z;
        }
        st8(y + off + (size_t)r * g.C, f);
    }
}

// backward pass 1: per (n, chunk, channel) sums of dz and dz*xhat  (dz = dy * silu'(z))
//   wsb[((n * chunks + chunk) * 2 + {0,1}) * C + c]
template <bool kSilu>
__global__ void gn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    const float* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, GNGeom g, float* __restrict__ wsb) {
    pdl_trigger();
    extern __shared__ float sh[];  // [2][C]
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int v = threadIdx.x % g.cv, rr = threadIdx.x / g.cv;
    const int cpg = g.C / g.G;
    float mean[8], rstd[8], ga[8], be[8], s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = v * 8 + j, gi = c / cpg;
        mean[j] = stats[((size_t)n * g.G + gi) * 2];
        rstd[j] = stats[((size_t)n * g.G + gi) * 2 + 1];
        ga[j] = gamma[c];
        be[j] = beta[c];
        s1[j] = s2[j] = 0.f;
    }
    const int r0 = chunk * g.rows_per_chunk;
    const int r1 = min(g.HW, r0 + g.rows_per_chunk);
    const size_t off = ((size_t)n * g.HW) * g.C + v * 8;
#pragma unroll 4
    for (int r = r0 + rr; r < r1; r += g.rpi) {
        float f[8], d[8];
        ld8(x + off + (size_t)r * g.C, f);
        ld8(dy + off + (size_t)r * g.C, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (f[j] - mean[j]) * rstd[j];
            float dz = d[j];
            if (kSilu) dz *= silu_grad_f(fmaf(xh, ga[j], be[j]));
            s1[j] += dz;
            s2[j] = fmaf(dz, xh, s2[j]);
        }
    }
    {
        float* mine = sh + (size_t)rr * 2 * g.C;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mine[v * 8 + j] = s1[j];
            mine[g.C + v * 8 + j] = s2[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) {
        float a = sh[i];
        for (int k = 1; k < g.rpi; ++k) a += sh[(size_t)k * 2 * g.C + i];
        sh[i] = a;
    }
    __syncthreads();
    float* o = wsb + ((size_t)n * g.chunks + chunk) * 2 * g.C;
    for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) o[i] = sh[i];
}

// backward pass 1b: reduce chunk partials -> per (n, c) sums; then per (n, g): A = sum_c gamma*S1, B = sum_c gamma*S2;
// also accumulates dgamma/dbeta (sum over n) when requested.
//   red[n, {0,1}, g] = {A/cnt, B/cnt}
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const float* __restrict__ wsb, const float* __restrict__ gamma, GNGeom g,
                                                            int gpb, float* __restrict__ red, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
    pdl_trigger();
    // block = (image n, gpb consecutive groups): its nc = gpb * cpg channels are summed over the chunk partials by
    // 256 / nc lanes per channel (fixed summation order: deterministic), then folded per group
    extern __shared__ float sh[];  // [lanes][2][nc] partial sums, then [2][nc] channel sums in place of lane 0
    const int n = blockIdx.x, g0 = blockIdx.y * gpb;
    const int cpg = g.C / g.G;
    const int ng = min(gpb, g.G - g0), nc = ng * cpg, c0 = g0 * cpg;
    const int lanes = max(1, (int)blockDim.x / nc);
    const int cl = threadIdx.x % nc, kl = threadIdx.x / nc;
    if (kl < lanes) {
        float a = 0.f, b = 0.f;
        for (int k = kl; k < g.chunks; k += lanes) {
            const float* p = wsb + ((size_t)n * g.chunks + k) * 2 * g.C + c0 + cl;
            a += p[0];
            b += p[g.C];
        }
        sh[(kl * 2) * nc + cl] = a;
        sh[(kl * 2 + 1) * nc + cl] = b;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int l = 0; l < lanes; ++l) {
            a += sh[(l * 2) * nc + c];
            b += sh[(l * 2 + 1) * nc + c];
        }
        sh[c] = a;        // lane 0 slots now hold the channel totals (each slot is read before it is overwritten: same thread)
        sh[nc + c] = b;
        if (dgamma) atomicAdd(&dgamma[c0 + c], b);
        if (dbeta) atomicAdd(&dbeta[c0 + c], a);
    }
    __syncthreads();
    const float cnt = (float)g.HW * (float)cpg;
    for (int gi = threadIdx.x; gi < ng; gi += blockDim.x) {
        float A = 0.f, B = 0.f;
        for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
            A = fmaf(gamma[c0 + c], sh[c], A);
            B = fmaf(gamma[c0 + c], sh[nc + c], B);
        }
        red[((size_t)n * 2) * g.G + g0 + gi] = A / cnt;
        red[((size_t)n * 2 + 1) * g.G + g0 + gi] = B / cnt;
    }
}

// backward pass 2: dx = rstd * (gamma*dz - A - xhat*B) (+ dres)
template <bool kSilu>
__global__ void gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    const float* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ red,
                                    const __nv_bfloat16* __restrict__ dres, GNGeom g, __nv_bfloat16* __restrict__ dx) {
    pdl_trigger();
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int v = threadIdx.x % g.cv, rr = threadIdx.x / g.cv;
    const int cpg = g.C / g.G;
    float mean[8], rstd[8], ga[8], be[8], A[8], Bc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = v * 8 + j, gi = c / cpg;
        mean[j] = stats[((size_t)n * g.G + gi) * 2];
        rstd[j] = stats[((size_t)n * g.G + gi) * 2 + 1];
        ga[j] = gamma[c];
        be[j] = beta[c];
        A[j] = red[((size_t)n * 2) * g.G + gi];
        Bc[j] = red[((size_t)n * 2 + 1) * g.G + gi];
    }
    const int r0 = chunk * g.rows_per_chunk;
    const int r1 = min(g.HW, r0 + g.rows_per_chunk);
    const size_t off = ((size_t)n * g.HW) * g.C + v * 8;
#pragma unroll 4
    for (int r = r0 + rr; r < r1; r += g.rpi) {
        float f[8], d[8], o[8];
        ld8(x + off + (size_t)r * g.C, f);
        ld8(dy + off + (size_t)r * g.C, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (f[j] - mean[j]) * rstd[j];
            float dz = d[j];
            if (kSilu) dz *= silu_grad_f(fmaf(xh, ga[j], be[j]));
            o[j] = rstd[j] * (ga[j] * dz - A[j] - xh * Bc[j]);
        }
        if (dres) {
            float e[8];
            ld8(dres + off + (size_t)r * g.C, e);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += e[j];
        }
        st8(dx + off + (size_t)r * g.C, o);
    }
}

// ------------------------------------------------------------------------------------------------
// One-launch GroupNorm: statistics pass, per-image barrier, apply pass.  All blocks of the grid are co-resident (the host sizes
// the grid from the occupancy query), so the blocks of one image can wait for each other on a self-resetting counter /
// generation pair in global memory; the second pass then re-reads the block's own rows while they are still in L2 (a 10-40 MB
// image easily fits the 126 MB L2), which removes one of the two (forward) / two of the five (backward) HBM passes of the
// three-kernel path above.
// ------------------------------------------------------------------------------------------------
constexpr int kGnBatch = 4;
UWU_DEVINL void gn_unpack8(const uint4& u, float (&v)[8]) {
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

UWU_DEVINL unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// sync[2 * n] = arrival counter, sync[2 * n + 1] = generation of image slot n (zero before the first use; both are left
// consistent for the next launch: the last arriver resets the counter before it bumps the generation)
UWU_DEVINL void gn_image_barrier(unsigned* sync, int n, unsigned blocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* cnt = sync + 2 * n;
        unsigned* gen = cnt + 1;
        const unsigned g0 = ld_acquire_u32(gen);
        __threadfence();
        const unsigned old = atomicAdd(cnt, 1u);
        if (old == blocks - 1) {
            *cnt = 0u;
            __threadfence();
            atomicAdd(gen, 1u);
        } else {
            while (ld_acquire_u32(gen) == g0) __nanosleep(40);
        }
        __threadfence();
    }
    __syncthreads();
}

template <bool kSilu>
__global__ void __launch_bounds__(256) gn_fwd_fused_kernel(const __nv_bfloat16* __restrict__ x, GNGeom g, float eps,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           __nv_bfloat16* __restrict__ y, float* __restrict__ stats,
                                                           float* __restrict__ ws, unsigned* __restrict__ sync) {
    pdl_trigger();
    extern __shared__ float sh[];  // [rpi][2][C] partials, later [G][2] statistics
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int v = threadIdx.x % g.cv, rr = threadIdx.x / g.cv;
    const int cpg = g.C / g.G;
    const int r0 = chunk * g.rows_per_chunk;
    const int r1 = min(g.HW, r0 + g.rows_per_chunk);
    const size_t off = ((size_t)n * g.HW) * g.C + v * 8;
    {
        float s[8], q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
        // kGnBatch independent 16-byte loads per thread are issued before the first one is consumed (left to the compiler,
        // the unrolled loop interleaved every load with the previous row's math).  Worth 5-10 % (110.6 -> 103.9 us at
        // 16 x 128^2 x 320): the two passes, the barrier between them and the single wave's ramp / tail add up, no pass
        // alone is short of loads in flight any more
        for (int r = r0 + rr; r < r1; r += g.rpi * kGnBatch) {
            uint4 raw[kGnBatch];
#pragma unroll
            for (int u = 0; u < kGnBatch; ++u) {
                const int ru = r + u * g.rpi;
                raw[u] = ru < r1 ? *reinterpret_cast<const uint4*>(x + off + (size_t)ru * g.C) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < kGnBatch; ++u) {
                float f[8];
                gn_unpack8(raw[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {  // (rows past the chunk are zeros: they add nothing)
                    s[j] += f[j];
                    q[j] = fmaf(f[j], f[j], q[j]);
                }
            }
        }
        float* mine = sh + (size_t)rr * 2 * g.C;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mine[v * 8 + j] = s[j];
            mine[g.C + v * 8 + j] = q[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) {
        float a = sh[i];
        for (int k = 1; k < g.rpi; ++k) a += sh[(size_t)k * 2 * g.C + i];
        sh[i] = a;
    }
    __syncthreads();
    for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
            a += sh[c];
            b += sh[g.C + c];
        }
        float* o = ws + (((size_t)n * g.chunks + chunk) * g.G + gi) * 2;
        o[0] = a;
        o[1] = b;
    }
    gn_image_barrier(sync, n, (unsigned)g.chunks);
    // every block finalises the statistics of its image (chunks * G * 2 floats from L2), block 0 publishes them
    for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int c = 0; c < g.chunks; ++c) {
            const float2 pp = __ldcg(reinterpret_cast<const float2*>(ws + (((size_t)n * g.chunks + c) * g.G + gi) * 2));
            a += pp.x;
            b += pp.y;
        }
        const float cnt = (float)g.HW * (float)cpg;
        const float mean = a / cnt;
        const float var = fmaxf(b / cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        sh[gi * 2] = mean;
        sh[gi * 2 + 1] = rstd;
        if (chunk == 0) {
            stats[((size_t)n * g.G + gi) * 2] = mean;
            stats[((size_t)n * g.G + gi) * 2 + 1] = rstd;
        }
    }
    __syncthreads();
    float sc[8], sf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = v * 8 + j, gi = c / cpg;
        sc[j] = sh[gi * 2 + 1] * gamma[c];
        sf[j] = beta[c] - sh[gi * 2] * sc[j];
    }
    for (int r = r0 + rr; r < r1; r += g.rpi * kGnBatch) {
        uint4 raw[kGnBatch];
#pragma unroll
        for (int u = 0; u < kGnBatch; ++u) {
            const int ru = r + u * g.rpi;
            raw[u] = ru < r1 ? *reinterpret_cast<const uint4*>(x + off + (size_t)ru * g.C) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kGnBatch; ++u) {
            const int ru = r + u * g.rpi;
            float f[8];
            gn_unpack8(raw[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float z = fmaf(f[j], sc[j], sf[j]);
                f[j] = kSilu ? silu_f(z) : z;
            }
            if (ru < r1) st8(y + off + (size_t)ru * g.C, f);
        }
    }
}

// backward without parameter gradients (frozen affine): per-chunk, per-group sums A = sum gamma*dz, B = sum gamma*dz*xhat
// -> image barrier -> dx
template <bool kSilu>
__global__ void __launch_bounds__(256, 3) gn_bwd_fused_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                           const float* __restrict__ stats, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const __nv_bfloat16* __restrict__ dres,
                                                           GNGeom g, __nv_bfloat16* __restrict__ dx, float* __restrict__ ws,
                                                           unsigned* __restrict__ sync) {
    pdl_trigger();
    extern __shared__ float sh[];  // [rpi][2][C] partials, later [G][2] = {A / cnt, B / cnt}
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int v = threadIdx.x % g.cv, rr = threadIdx.x / g.cv;
    const int cpg = g.C / g.G;
    float mean[8], rstd[8], ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = v * 8 + j, gi = c / cpg;
        mean[j] = stats[((size_t)n * g.G + gi) * 2];
        rstd[j] = stats[((size_t)n * g.G + gi) * 2 + 1];
        ga[j] = gamma[c];
        be[j] = beta[c];
    }
    const int r0 = chunk * g.rows_per_chunk;
    const int r1 = min(g.HW, r0 + g.rows_per_chunk);
    const size_t off = ((size_t)n * g.HW) * g.C + v * 8;
    {
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
        for (int r = r0 + rr; r < r1; r += g.rpi * kGnBatch) {
            uint4 rx[kGnBatch], rd[kGnBatch];
#pragma unroll
            for (int u = 0; u < kGnBatch; ++u) {
                const int ru = r + u * g.rpi;
                const bool ok = ru < r1;
                rx[u] = ok ? *reinterpret_cast<const uint4*>(x + off + (size_t)ru * g.C) : make_uint4(0, 0, 0, 0);
                rd[u] = ok ? *reinterpret_cast<const uint4*>(dy + off + (size_t)ru * g.C) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < kGnBatch; ++u) {
                float f[8], d[8];
                gn_unpack8(rx[u], f);
                gn_unpack8(rd[u], d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {  // (rows past the chunk carry dy = 0: they add nothing)
                    const float xh = (f[j] - mean[j]) * rstd[j];
                    float dz = d[j];
                    if (kSilu) dz *= silu_grad_f(fmaf(xh, ga[j], be[j]));
                    s1[j] += dz;
                    s2[j] = fmaf(dz, xh, s2[j]);
                }
            }
        }
        float* mine = sh + (size_t)rr * 2 * g.C;
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // gamma-weighted: only the per-group sums are needed
            mine[v * 8 + j] = s1[j] * ga[j];
            mine[g.C + v * 8 + j] = s2[j] * ga[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) {
        float a = sh[i];
        for (int k = 1; k < g.rpi; ++k) a += sh[(size_t)k * 2 * g.C + i];
        sh[i] = a;
    }
    __syncthreads();
    for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
            a += sh[c];
            b += sh[g.C + c];
        }
        float* o = ws + (((size_t)n * g.chunks + chunk) * g.G + gi) * 2;
        o[0] = a;
        o[1] = b;
    }
    gn_image_barrier(sync, n, (unsigned)g.chunks);
    for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int c = 0; c < g.chunks; ++c) {
            const float2 pp = __ldcg(reinterpret_cast<const float2*>(ws + (((size_t)n * g.chunks + c) * g.G + gi) * 2));
            a += pp.x;
            b += pp.y;
        }
        const float cnt = (float)g.HW * (float)cpg;
        sh[gi * 2] = a / cnt;
        sh[gi * 2 + 1] = b / cnt;
    }
    __syncthreads();
    float A[8], Bc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int gi = (v * 8 + j) / cpg;
        A[j] = sh[gi * 2];
        Bc[j] = sh[gi * 2 + 1];
    }
    constexpr int kB2 = 2;  // x, dy and dres of two rows in flight (six 16-byte loads per thread at 80 registers)
    for (int r = r0 + rr; r < r1; r += g.rpi * kB2) {
        uint4 rx[kB2], rd[kB2], re[kB2];
#pragma unroll
        for (int u = 0; u < kB2; ++u) {
            const int ru = r + u * g.rpi;
            const bool ok = ru < r1;
            rx[u] = ok ? *reinterpret_cast<const uint4*>(x + off + (size_t)ru * g.C) : make_uint4(0, 0, 0, 0);
            rd[u] = ok ? *reinterpret_cast<const uint4*>(dy + off + (size_t)ru * g.C) : make_uint4(0, 0, 0, 0);
            re[u] = (ok && dres) ? *reinterpret_cast<const uint4*>(dres + off + (size_t)ru * g.C) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kB2; ++u) {
            const int ru = r + u * g.rpi;
            float f[8], d[8], e[8], o[8];
            gn_unpack8(rx[u], f);
            gn_unpack8(rd[u], d);
            gn_unpack8(re[u], e);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (f[j] - mean[j]) * rstd[j];
                float dz = d[j];
                if (kSilu) dz *= silu_grad_f(fmaf(xh, ga[j], be[j]));
                o[j] = rstd[j] * (ga[j] * dz - A[j] - xh * Bc[j]) + e[j];
            }
            if (ru < r1) st8(dx + off + (size_t)ru * g.C, o);
        }
    }
}

static int gn_geom(int N, int HW, int C, int G, GNGeom* g) {
    UWU_CHECK_ARG(N > 0 && HW > 0 && C > 0 && G > 0, "groupnorm: bad shape N=%d HW=%d C=%d G=%d", N, HW, C, G);
    UWU_CHECK_ARG(C % 8 == 0 && C % G == 0, "groupnorm: C=%d must be a multiple of 8 and of G=%d", C, G);
    UWU_CHECK_ARG(C <= 8192, "groupnorm: C=%d too large", C);
    g->N = N; g->HW = HW; g->C = C; g->G = G;
    g->cv = C / 8;
    int rpi = 256 / g->cv;
    if (rpi < 1) rpi = 1;
    if (g->cv * rpi > 1024) rpi = 1;
    UWU_CHECK_ARG(g->cv <= 1024, "groupnorm: C too large");
    g->rpi = rpi;
    int chunks = (8 * sm_count() + N - 1) / N;  // ~8 resident blocks per SM keep enough 16-byte loads in flight
    int max_chunks = (HW + rpi * 4 - 1) / (rpi * 4);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > 256) chunks = 256;
    if (chunks < 1) chunks = 1;
    g->rows_per_chunk = (HW + chunks - 1) / chunks;
    g->chunks = (HW + g->rows_per_chunk - 1) / g->rows_per_chunk;
    return 0;
}

// ================================================================================================
// LayerNorm: one warp per row, the row lives in registers (one HBM read, one write); kMaxV = ceil(C / 256)
// ================================================================================================
UWU_DEVINL void unpack8(const uint4& u, float (&v)[8]) {
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
UWU_DEVINL void ldf8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// y = ((x - mean) * rstd * gamma + beta) [* (1 + mscale[b]) + mshift[b]]   ;  stats[row] = {mean, rstd}
template <int kMaxV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, int M, int C, float eps,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     const float* __restrict__ mscale, const float* __restrict__ mshift,
                                                     int rows_per_mod, __nv_bfloat16* __restrict__ y,
                                                     float* __restrict__ stats) {
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cv = C / 8;
    const float inv_c = 1.0f / (float)C;
    for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
        const __nv_bfloat16* xr = x + (size_t)row * C;
        uint4 raw[kMaxV];
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            raw[k] = v < cv ? *reinterpret_cast<const uint4*>(xr + v * 8) : make_uint4(0, 0, 0, 0);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            float f[8];
            unpack8(raw[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[j];
        }
        const float mean = warp_sum(s) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            if (lane + k * 32 < cv) {
                float f[8];
                unpack8(raw[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = f[j] - mean;
                    q = fmaf(d, d, q);
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
        if (lane == 0 && stats) {
            stats[(size_t)row * 2] = mean;
            stats[(size_t)row * 2 + 1] = rstd;
        }
        const float* ms = mscale ? mscale + (size_t)(row / rows_per_mod) * C : nullptr;
        const float* mh = mshift ? mshift + (size_t)(row / rows_per_mod) * C : nullptr;
        __nv_bfloat16* yr = y + (size_t)row * C;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            if (v < cv) {
                float f[8];
                unpack8(raw[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd;
                if (gamma) {
                    float ga[8];
                    ldf8(gamma + v * 8, ga);
                    if (beta) {
                        float be[8];
                        ldf8(beta + v * 8, be);
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], ga[j], be[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] *= ga[j];
                    }
                }
                if (ms) {
                    float a[8];
                    ldf8(ms + v * 8, a);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = f[j] * (1.0f + a[j]);
                    if (mh) {
                        ldf8(mh + v * 8, a);
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] += a[j];
                    }
                }
                st8(yr + v * 8, f);
            }
        }
    }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres), g = gamma * dy
// partial param grads: pg[blockIdx.x][0][c] = sum_rows dy*xhat, pg[blockIdx.x][1][c] = sum_rows dy
// 4 warps per block, each streaming rows with x / dy / dres held packed in registers (all loads of a row in flight
// before the first use); per-lane fp32 accumulators for the parameter gradients.
template <int kMaxV, bool kParamGrads>
__global__ void __launch_bounds__(128, (kMaxV >= 5 && kParamGrads) ? 2 : 3) ln_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                        const __nv_bfloat16* __restrict__ dy, int M, int C,
                                                        const float* __restrict__ gamma, const float* __restrict__ stats,
                                                        const __nv_bfloat16* __restrict__ dres,
                                                        __nv_bfloat16* __restrict__ dx, float* __restrict__ pg) {
    extern __shared__ float sh[];  // gamma [C] ; later [2][C] partial sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cv = C / 8;
    const float inv_c = 1.0f / (float)C;
    for (int i = threadIdx.x; i < C; i += blockDim.x) sh[i] = gamma ? gamma[i] : 1.0f;
    __syncthreads();
    float ag[kParamGrads ? kMaxV : 1][8], ab[kParamGrads ? kMaxV : 1][8];
#pragma unroll
    for (int k = 0; k < (kParamGrads ? kMaxV : 1); ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) ag[k][j] = ab[k][j] = 0.f;
    for (int row = blockIdx.x * 4 + warp; row < M; row += gridDim.x * 4) {
        const __nv_bfloat16* xr = x + (size_t)row * C;
        const __nv_bfloat16* dr = dy + (size_t)row * C;
        uint4 rx[kMaxV], rd[kMaxV], rr[kMaxV];
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            const bool ok = v < cv;
            rx[k] = ok ? *reinterpret_cast<const uint4*>(xr + v * 8) : make_uint4(0, 0, 0, 0);
            rd[k] = ok ? *reinterpret_cast<const uint4*>(dr + v * 8) : make_uint4(0, 0, 0, 0);
            rr[k] = (ok && dres) ? *reinterpret_cast<const uint4*>(dres + (size_t)row * C + v * 8) : make_uint4(0, 0, 0, 0);
        }
        const float mean = stats[(size_t)row * 2], rstd = stats[(size_t)row * 2 + 1];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            if (v < cv) {
                float f[8], d[8], ga[8];
                unpack8(rx[k], f);
                unpack8(rd[k], d);
                ldf8(sh + v * 8, ga);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = (f[j] - mean) * rstd;
                    const float gv = d[j] * ga[j];
                    s1 += gv;
                    s2 = fmaf(gv, h, s2);
                    if (kParamGrads) {
                        ag[k][j] = fmaf(d[j], h, ag[k][j]);
                        ab[k][j] += d[j];
                    }
                }
            }
        }
        s1 = warp_sum(s1) * inv_c;
        s2 = warp_sum(s2) * inv_c;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            if (v < cv) {
                float f[8], d[8], ga[8], e[8], o[8];
                unpack8(rx[k], f);
                unpack8(rd[k], d);
                unpack8(rr[k], e);
                ldf8(sh + v * 8, ga);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = (f[j] - mean) * rstd;
                    o[j] = rstd * (d[j] * ga[j] - s1 - h * s2) + e[j];
                }
                st8(dx + (size_t)row * C + v * 8, o);
            }
        }
    }
    if (kParamGrads) {
        __syncthreads();  // everyone is done reading gamma from sh
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            if (v < cv) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    atomicAdd(&sh[v * 8 + j], ag[k][j]);
                    atomicAdd(&sh[C + v * 8 + j], ab[k][j]);
                }
            }
        }
        __syncthreads();
        float* o = pg + (size_t)blockIdx.x * 2 * C;
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) o[i] = sh[i];
    }
}

// LayerNorm backward WITH parameter gradients: thread = one 8-wide column vector of the row (block = ceil(C/256) warps
// covers a whole row), R rows per step.  Column ownership makes the dgamma / dbeta accumulators 16 registers per thread
// and their reduction deterministic (no atomics); the two per-row sums are combined across the block's warps through
// shared memory, one barrier per step (double-buffered partials).
//   pg[blockIdx.x][0][c] = sum_rows dy * xhat,  pg[blockIdx.x][1][c] = sum_rows dy
template <int R>
__global__ void __launch_bounds__(192, 4) ln_bwd_cols_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                          int M, int C, const float* __restrict__ gamma,
                                                          const float* __restrict__ stats, const __nv_bfloat16* __restrict__ dres,
                                                          __nv_bfloat16* __restrict__ dx, float* __restrict__ pg) {
    pdl_trigger();
    __shared__ float part[2][R][8][2];  // [parity][row][warp][s1, s2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int cv = C >> 3;
    const int v = threadIdx.x;
    const bool active = v < cv;
    const float inv_c = 1.0f / (float)C;
    float ga[8], ag[8], ab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[j] = ab[j] = 0.f;
    if (active) {
        if (gamma) ldf8(gamma + v * 8, ga);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) ga[j] = 1.0f;
        }
    }
    int par = 0;
    for (int r0 = blockIdx.x * R; r0 < M; r0 += gridDim.x * R, par ^= 1) {
        uint4 rx[R], rd[R], rr[R];
        float mean[R], rstd[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int row = r0 + k;
            const bool ok = active && row < M;
            const size_t off = (size_t)row * C + v * 8;
            rx[k] = ok ? *reinterpret_cast<const uint4*>(x + off) : make_uint4(0, 0, 0, 0);
            rd[k] = ok ? *reinterpret_cast<const uint4*>(dy + off) : make_uint4(0, 0, 0, 0);
            rr[k] = (ok && dres) ? *reinterpret_cast<const uint4*>(dres + off) : make_uint4(0, 0, 0, 0);
            mean[k] = row < M ? stats[(size_t)row * 2] : 0.f;
            rstd[k] = row < M ? stats[(size_t)row * 2 + 1] : 0.f;
        }
        float s1[R], s2[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float f[8], d[8];
            unpack8(rx[k], f);
            unpack8(rd[k], d);
            float a = 0.f, b = 0.f;
            if (active && r0 + k < M) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = (f[j] - mean[k]) * rstd[k];
                    const float gv = d[j] * ga[j];
                    a += gv;
                    b = fmaf(gv, h, b);
                    ag[j] = fmaf(d[j], h, ag[j]);
                    ab[j] += d[j];
                }
            }
            s1[k] = warp_sum(a);
            s2[k] = warp_sum(b);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                part[par][k][warp][0] = s1[k];
                part[par][k][warp][1] = s2[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float a = 0.f, b = 0.f;
            for (int w = 0; w < nw; ++w) {
                a += part[par][k][w][0];
                b += part[par][k][w][1];
            }
            a *= inv_c;
            b *= inv_c;
            if (active && r0 + k < M) {
                float f[8], d[8], e[8], o[8];
                unpack8(rx[k], f);
                unpack8(rd[k], d);
                unpack8(rr[k], e);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = (f[j] - mean[k]) * rstd[k];
                    o[j] = rstd[k] * (d[j] * ga[j] - a - h * b) + e[j];
                }
                st8(dx + (size_t)(r0 + k) * C + v * 8, o);
            }
        }
    }
    if (active) {
        float* o = pg + (size_t)blockIdx.x * 2 * C + v * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j] = ag[j];
            o[C + j] = ab[j];
        }
    }
}

// LayerNorm backward, streaming version: the rows are staged into shared memory by 1-D bulk copies (TMA) through an S-stage
// mbarrier ring issued by one thread, so every block keeps (S - 1) stages of x / dy / dres in flight regardless of its
// register count — the register-staged kernels above drain their loads at every per-row reduction and stop near 3.5 TB/s.
// Thread = one 8-wide column vector (block = ceil(C/256) warps covers a row), R rows per stage; per-row sums through warp
// shuffles + one __syncthreads per stage; dgamma / dbeta accumulate in registers (column ownership) and leave through
// ONE 1-D bulk reduction per block (cp.reduce.async.bulk .add.f32) into the pre-zeroed / accumulating outputs.
template <int R, bool kParamGrads>
__global__ void __launch_bounds__(192) ln_bwd_stream_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                            int M, int C, const float* __restrict__ gamma,
                                                            const float* __restrict__ stats, const __nv_bfloat16* __restrict__ dres,
                                                            __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int n_stages) {
    pdl_trigger();
    extern __shared__ __align__(128) uint8_t ln_smem[];
    __shared__ __align__(8) uint64_t full[8];
    __shared__ float part[2][R][8][2];  // [parity][row][warp][s1, s2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int cv = C >> 3;
    const int v = threadIdx.x;
    const bool active = v < cv;
    const float inv_c = 1.0f / (float)C;
    const int n_arr = dres ? 3 : 2;
    const uint32_t row_bytes = (uint32_t)C * 2u, arr_bytes = R * row_bytes, stage_bytes = n_arr * arr_bytes;
    const int n_steps = (M - (int)blockIdx.x * R + (int)gridDim.x * R - 1) / ((int)gridDim.x * R);  // steps of this block

    auto issue = [&](int it) {  // thread 0
        const int r0 = ((int)blockIdx.x + it * (int)gridDim.x) * R;
        const uint32_t bytes = (uint32_t)min(R, M - r0) * row_bytes;
        const int sidx = it % n_stages;
        uint8_t* dst = ln_smem + (size_t)sidx * stage_bytes;
        mbar_expect_tx(&full[sidx], bytes * n_arr);
        bulk_load_1d(dst, x + (size_t)r0 * C, bytes, &full[sidx]);
        bulk_load_1d(dst + arr_bytes, dy + (size_t)r0 * C, bytes, &full[sidx]);
        if (dres) bulk_load_1d(dst + 2 * arr_bytes, dres + (size_t)r0 * C, bytes, &full[sidx]);
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
        for (int it = 0; it < n_stages - 1 && it < n_steps; ++it) issue(it);
    }
    float ga[8], ag[8], ab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[j] = ab[j] = 0.f, ga[j] = 1.0f;
    if (active && gamma) ldf8(gamma + v * 8, ga);
    __syncthreads();
    for (int it = 0; it < n_steps; ++it) {
        const int par = it & 1;
        const int r0 = ((int)blockIdx.x + it * (int)gridDim.x) * R;
        // the stage consumed in step it - 1 is free: every thread copied its vectors to registers before that step's barrier
        if (threadIdx.x == 0 && it + n_stages - 1 < n_steps) issue(it + n_stages - 1);
        const int sidx = it % n_stages;
        const uint8_t* st = ln_smem + (size_t)sidx * stage_bytes + (size_t)v * 16;
        float mean[R], rstd[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int row = r0 + k;
            const float2 ms = row < M ? *reinterpret_cast<const float2*>(stats + (size_t)row * 2) : make_float2(0.f, 0.f);
            mean[k] = ms.x;
            rstd[k] = ms.y;
        }
        mbar_wait(&full[sidx], (uint32_t)(it / n_stages) & 1u);
        uint4 rx[R], rd[R], rr[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const bool ok = active && r0 + k < M;
            rx[k] = ok ? *reinterpret_cast<const uint4*>(st + (size_t)k * row_bytes) : make_uint4(0, 0, 0, 0);
            rd[k] = ok ? *reinterpret_cast<const uint4*>(st + arr_bytes + (size_t)k * row_bytes) : make_uint4(0, 0, 0, 0);
            rr[k] = (ok && dres) ? *reinterpret_cast<const uint4*>(st + 2 * arr_bytes + (size_t)k * row_bytes) : make_uint4(0, 0, 0, 0);
        }
        float s1[R], s2[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float f[8], d[8];
            unpack8(rx[k], f);
            unpack8(rd[k], d);
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // inactive lanes / rows hold zeros
                const float h = (f[j] - mean[k]) * rstd[k];
                const float gv = d[j] * ga[j];
                a += gv;
                b = fmaf(gv, h, b);
                if (kParamGrads) {
                    ag[j] = fmaf(d[j], h, ag[j]);
                    ab[j] += d[j];
                }
            }
            s1[k] = warp_sum(a);
            s2[k] = warp_sum(b);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                part[par][k][warp][0] = s1[k];
                part[par][k][warp][1] = s2[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float a = 0.f, b = 0.f;
            for (int w = 0; w < nw; ++w) {
                a += part[par][k][w][0];
                b += part[par][k][w][1];
            }
            a *= inv_c;
            b *= inv_c;
            if (active && r0 + k < M) {
                float f[8], d[8], e[8], o[8];
                unpack8(rx[k], f);
                unpack8(rd[k], d);
                unpack8(rr[k], e);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = (f[j] - mean[k]) * rstd[k];
                    o[j] = rstd[k] * (d[j] * ga[j] - a - h * b) + e[j];
                }
                st8(dx + (size_t)(r0 + k) * C + v * 8, o);
            }
        }
    }
    if (kParamGrads) {
        // block partials -> shared -> one bulk reduction per output (stage 0 is free: all loads were consumed)
        __syncthreads();
        float* sh = reinterpret_cast<float*>(ln_smem);
        if (active) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sh[v * 8 + j] = ag[j];
                sh[C + v * 8 + j] = ab[j];
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (dgamma) bulk_reduce_add_f32(dgamma, sh, (uint32_t)C * 4u);
            if (dbeta) bulk_reduce_add_f32(dbeta, sh + C, (uint32_t)C * 4u);
            bulk_commit_group();
            bulk_wait_group0();
        }
    }
}

// out[c] (+)= sum_p partial[p, c]
__global__ void colsum_partials_kernel(const float* __restrict__ partial, int P, int stride, int n, int accumulate,
                                       float* __restrict__ out) {
    pdl_trigger();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float a = 0.f;
    for (int p = 0; p < P; ++p) a += partial[(size_t)p * stride + c];
    out[c] = accumulate ? out[c] + a : a;
}

// LayerNorm parameter gradients: partial [P][2][C] -> dgamma[c] += sum_p partial[p][0][c], dbeta[c] += sum_p partial[p][1][c].
// grid (ceil(2C/128), slices): each block sums its slice of p and adds atomically (outputs pre-zeroed unless accumulating).
__global__ void __launch_bounds__(128) ln_param_reduce_kernel(const float* __restrict__ partial, int P, int C, int p_per,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
    pdl_trigger();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * C) return;
    const int p0 = blockIdx.y * p_per, p1 = min(P, p0 + p_per);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int p = p0;
    for (; p + 3 < p1; p += 4) {
        a0 += partial[(size_t)p * 2 * C + c];
        a1 += partial[(size_t)(p + 1) * 2 * C + c];
        a2 += partial[(size_t)(p + 2) * 2 * C + c];
        a3 += partial[(size_t)(p + 3) * 2 * C + c];
    }
    for (; p < p1; ++p) a0 += partial[(size_t)p * 2 * C + c];
    const float a = (a0 + a1) + (a2 + a3);
    if (c < C) {
        if (dgamma) atomicAdd(&dgamma[c], a);
    } else {
        if (dbeta) atomicAdd(&dbeta[c - C], a);
    }
}

// one resident wave of 256-thread blocks: `per_sm` = blocks that fit an SM at the kernel's register count (the kernels stride
// over the rows, so a larger grid only adds a partially filled last wave)
static int ln_grid(int M, int per_sm = 8) {
    int blocks = (M + 7) / 8;
    const int cap = sm_count() * per_sm;
    return blocks < cap ? blocks : cap;
}

}  // namespace uwu

using namespace uwu;

extern "C" int64_t uwu_groupnorm_workspace_floats(int32_t N, int32_t HW, int32_t C, int32_t G) {
    GNGeom g;
    if (gn_geom(N, HW, C, G, &g)) return -1;
    // forward partials (N*chunks*G*2) or backward partials (N*chunks*2*C) + reduced (N*2*G)
    // (+ the per-group partials of the one-launch kernels, whose chunk count is bounded by 256)
    const int64_t three = (int64_t)N * g.chunks * 2 * C + (int64_t)N * 2 * G + 64;
    const int64_t one = (int64_t)N * 256 * 2 * G + 64;
    return three > one ? three : one;
}

extern "C" int uwu_groupnorm_fwd(const void* x, int32_t N, int32_t HW, int32_t C, int32_t G, float eps,
                                 const float* gamma, const float* beta, int32_t fuse_silu, void* y, float* stats,
                                 float* workspace, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    GNGeom g;
    if (gn_geom(N, HW, C, G, &g)) return UWU_ERR_INVALID;
    UWU_CHECK_ARG(x && y && stats && workspace && gamma && beta, "uwu_groupnorm_fwd: null pointer");
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
    dim3 grid(g.chunks, N);
    const int threads = g.cv * g.rpi;
    gn_stats_kernel<<<grid, threads, (size_t)g.rpi * 2 * C * sizeof(float), stream>>>(xp, g, workspace);
    UWU_CHECK_LAUNCH();
    gn_finalize_kernel<<<(N * G * 32 + 255) / 256, 256, 0, stream>>>(workspace, g, eps, stats);
    UWU_CHECK_LAUNCH();
    if (fuse_silu)
        gn_apply_kernel<true><<<grid, threads, 0, stream>>>(xp, stats, gamma, beta, g, yp);
    else
        gn_apply_kernel<false><<<grid, threads, 0, stream>>>(xp, stats, gamma, beta, g, yp);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

// grid of the one-launch kernels: all N * chunks blocks must be resident at once
template <typename K>
static int gn_fused_geom(K kernel, int N, int HW, int C, int G, GNGeom* g) {
    if (gn_geom(N, HW, C, G, g)) return -1;
    const int threads = g->cv * g->rpi;
    const size_t smem = (size_t)g->rpi * 2 * C * sizeof(float);
    if (threads > 256 || smem > 48 * 1024 || (size_t)G * 2 * sizeof(float) > smem) return 1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) return 1;
    const int cap = per_sm * sm_count();
    int chunks = cap / N;
    if (chunks < 1) return 1;  // more images than resident blocks
    const int max_chunks = (HW + g->rpi * 4 - 1) / (g->rpi * 4);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > 256) chunks = 256;
    g->rows_per_chunk = (HW + chunks - 1) / chunks;
    g->chunks = (HW + g->rows_per_chunk - 1) / g->rows_per_chunk;
    return 0;
}

/* One-launch forward; `sync` = 2 * N uint32 that are ZERO before the first call and are left zero-consistent by every call
 * (a persistent per-device buffer).  Returns 1 (nothing launched) when the shape cannot run as one resident wave. */
extern "C" int uwu_groupnorm_fwd_fused(const void* x, int32_t N, int32_t HW, int32_t C, int32_t G, float eps,
                                       const float* gamma, const float* beta, int32_t fuse_silu, void* y, float* stats,
                                       float* workspace, uint32_t* sync, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(x && y && stats && workspace && gamma && beta && sync, "uwu_groupnorm_fwd_fused: null pointer");
    GNGeom g;
    const int rc = fuse_silu ? gn_fused_geom(gn_fwd_fused_kernel<true>, N, HW, C, G, &g) : gn_fused_geom(gn_fwd_fused_kernel<false>, N, HW, C, G, &g);
    if (rc < 0) return UWU_ERR_INVALID;
    if (rc > 0) return 1;
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
    dim3 grid(g.chunks, N);
    const int threads = g.cv * g.rpi;
    const size_t smem = (size_t)g.rpi * 2 * C * sizeof(float);
    if (fuse_silu)
        gn_fwd_fused_kernel<true><<<grid, threads, smem, stream>>>(xp, g, eps, gamma, beta, yp, stats, workspace, sync);
    else
        gn_fwd_fused_kernel<false><<<grid, threads, smem, stream>>>(xp, g, eps, gamma, beta, yp, stats, workspace, sync);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

/* One-launch backward for a frozen affine (no dgamma / dbeta). */
extern "C" int uwu_groupnorm_bwd_fused(const void* x, const void* dy, int32_t N, int32_t HW, int32_t C, int32_t G,
                                       const float* gamma, const float* beta, const float* stats, int32_t fuse_silu,
                                       const void* dres, void* dx, float* workspace, uint32_t* sync, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(x && dy && dx && stats && workspace && gamma && beta && sync, "uwu_groupnorm_bwd_fused: null pointer");
    GNGeom g;
    const int rc = fuse_silu ? gn_fused_geom(gn_bwd_fused_kernel<true>, N, HW, C, G, &g) : gn_fused_geom(gn_bwd_fused_kernel<false>, N, HW, C, G, &g);
    if (rc < 0) return UWU_ERR_INVALID;
    if (rc > 0) return 1;
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    const auto* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
    const auto* rp = reinterpret_cast<const __nv_bfloat16*>(dres);
    auto* dxp = reinterpret_cast<__nv_bfloat16*>(dx);
    dim3 grid(g.chunks, N);
    const int threads = g.cv * g.rpi;
    const size_t smem = (size_t)g.rpi * 2 * C * sizeof(float);
    if (fuse_silu)
        gn_bwd_fused_kernel<true><<<grid, threads, smem, stream>>>(xp, dyp, stats, gamma, beta, rp, g, dxp, workspace, sync);
    else
        gn_bwd_fused_kernel<false><<<grid, threads, smem, stream>>>(xp, dyp, stats, gamma, beta, rp, g, dxp, workspace, sync);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_groupnorm_bwd(const void* x, const void* dy, int32_t N, int32_t HW, int32_t C, int32_t G,
                                 const float* gamma, const float* beta, const float* stats, int32_t fuse_silu,
                                 const void* dres, void* dx, float* dgamma, float* dbeta, float* workspace,
                                 void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    GNGeom g;
    if (gn_geom(N, HW, C, G, &g)) return UWU_ERR_INVALID;
    UWU_CHECK_ARG(x && dy && dx && stats && workspace && gamma && beta, "uwu_groupnorm_bwd: null pointer");
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    const auto* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
    dim3 grid(g.chunks, N);
    const int threads = g.cv * g.rpi;
    float* wsb = workspace;
    float* red = workspace + (size_t)N * g.chunks * 2 * C;
    if (fuse_silu)
        gn_bwd_stats_kernel<true><<<grid, threads, (size_t)g.rpi * 2 * C * sizeof(float), stream>>>(xp, dyp, stats, gamma, beta, g, wsb);
    else
        gn_bwd_stats_kernel<false><<<grid, threads, (size_t)g.rpi * 2 * C * sizeof(float), stream>>>(xp, dyp, stats, gamma, beta, g, wsb);
    UWU_CHECK_LAUNCH();
    {
        const int cpg = C / G;
        int gpb = 128 / cpg;  // ~128 channels per block
        if (gpb < 1) gpb = 1;
        if (gpb > G) gpb = G;
        const int nc = gpb * cpg;
        UWU_CHECK_ARG(nc <= 256, "uwu_groupnorm_bwd: more than 256 channels per group");
        const int lanes = 256 / nc > 0 ? 256 / nc : 1;
        gn_bwd_reduce_kernel<<<dim3(N, (G + gpb - 1) / gpb), 256, (size_t)lanes * 2 * nc * sizeof(float), stream>>>(wsb, gamma, g, gpb, red,
                                                                                                            dgamma, dbeta);
    }
    UWU_CHECK_LAUNCH();
    if (fuse_silu)
        gn_bwd_apply_kernel<true><<<grid, threads, 0, stream>>>(xp, dyp, stats, gamma, beta, red,
                                                                reinterpret_cast<const __nv_bfloat16*>(dres), g,
                                                                reinterpret_cast<__nv_bfloat16*>(dx));
    else
        gn_bwd_apply_kernel<false><<<grid, threads, 0, stream>>>(xp, dyp, stats, gamma, beta, red,
                                                                 reinterpret_cast<const __nv_bfloat16*>(dres), g,
                                                                 reinterpret_cast<__nv_bfloat16*>(dx));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_layernorm_fwd(const void* x, int32_t M, int32_t C, float eps, const float* gamma, const float* beta,
                                 const float* mod_scale, const float* mod_shift, int32_t rows_per_mod, void* y,
                                 float* stats, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(M >= 0 && C > 0 && C % 8 == 0, "uwu_layernorm_fwd: bad shape M=%d C=%d (C must be a multiple of 8)", M, C);
    if (M == 0) return UWU_OK;
    UWU_CHECK_ARG(x && y, "uwu_layernorm_fwd: null pointer");
    UWU_CHECK_ARG(C <= 8 * 32 * 8, "uwu_layernorm_fwd: C=%d > 2048 unsupported", C);
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
    const int rpm = rows_per_mod > 0 ? rows_per_mod : 1;
    const int cv = C / 8;
    // registers per thread: <1> 32, <2> 53, <3> 64, <5> 79, <8> 127  ->  resident 256-thread blocks per SM 8 / 4 / 4 / 3 / 2
#define UWU_LN_FWD(V) ln_fwd_kernel<V><<<ln_grid(M, (V) == 1 ? 8 : (V) <= 3 ? 4 : (V) == 5 ? 3 : 2), 256, 0, stream>>>(xp, M, C, eps, gamma, beta, mod_scale, mod_shift, rpm, yp, stats)
    if (cv <= 32) UWU_LN_FWD(1);
    else if (cv <= 64) UWU_LN_FWD(2);
    else if (cv <= 96) UWU_LN_FWD(3);
    else if (cv <= 160) UWU_LN_FWD(5);
    else UWU_LN_FWD(8);
#undef UWU_LN_FWD
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

static int ln_bwd_grid(int M) {
    int blocks = (M + 3) / 4;
    const int cap = sm_count() * 6;
    return blocks < cap ? blocks : cap;
}

extern "C" int64_t uwu_layernorm_bwd_workspace_floats(int32_t M, int32_t C) {
    if (M <= 0 || C <= 0) return 0;
    return (int64_t)ln_bwd_grid(M) * 2 * C;
}

extern "C" int uwu_layernorm_bwd(const void* x, const void* dy, int32_t M, int32_t C, const float* gamma,
                                 const float* stats, const void* dres, void* dx, float* dgamma, float* dbeta,
                                 int32_t accumulate, float* workspace, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(M > 0 && C > 0 && C % 8 == 0, "uwu_layernorm_bwd: bad shape M=%d C=%d", M, C);
    UWU_CHECK_ARG(C <= 8 * 32 * 5, "uwu_layernorm_bwd: C=%d > 1280 unsupported", C);
    UWU_CHECK_ARG(x && dy && dx && stats, "uwu_layernorm_bwd: null pointer");
    const bool want_pg = dgamma != nullptr || dbeta != nullptr;
    UWU_CHECK_ARG(!want_pg || workspace, "uwu_layernorm_bwd: workspace required for parameter gradients");
    const bool aligned16 = (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dres | (uintptr_t)dx | (uintptr_t)dgamma | (uintptr_t)dbeta) & 15) == 0;
    static const bool stream_off = getenv("UWU_LN_STREAM") && atoi(getenv("UWU_LN_STREAM")) == 0;
    if (C <= 192 * 8 && aligned16 && !stream_off) {
        // streaming kernel: R rows per stage so that a stage carries >= ~5 KB per array; stages sized to keep >= 3 blocks
        // (or 160 KB) of loads in flight per SM
        const int threads = ((C / 8 + 31) / 32) * 32;
        const int R = C >= 1024 ? 2 : (C >= 512 ? 4 : 8);
        const size_t stage = (size_t)(dres ? 3 : 2) * R * C * 2;
        int n_stages = 4;
        size_t smem = stage * n_stages;
        if (smem < (size_t)2 * C * 4) smem = (size_t)2 * C * 4;
        int per_sm = (int)((size_t)200 * 1024 / (smem + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        int grid = (M + R - 1) / R;
        if (grid > per_sm * sm_count()) grid = per_sm * sm_count();
        if (want_pg && !accumulate) {
            if (dgamma) UWU_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, stream));
            if (dbeta) UWU_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, stream));
        }
#define UWU_LN_BWD_S(RR, PG)                                                                                                  \
    do {                                                                                                                     \
        static bool attr_done = false;                                                                                       \
        if (!attr_done) {                                                                                                    \
            UWU_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_stream_kernel<RR, PG>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                200 * 1024));                                                                \
            attr_done = true;                                                                                                \
        }                                                                                                                    \
        ln_bwd_stream_kernel<RR, PG><<<grid, threads, smem, stream>>>(                                                       \
            reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), M, C, gamma, stats,       \
            reinterpret_cast<const __nv_bfloat16*>(dres), reinterpret_cast<__nv_bfloat16*>(dx), dgamma, dbeta, n_stages);    \
    } while (0)
        if (R == 2) { if (want_pg) UWU_LN_BWD_S(2, true); else UWU_LN_BWD_S(2, false); }
        else if (R == 4) { if (want_pg) UWU_LN_BWD_S(4, true); else UWU_LN_BWD_S(4, false); }
        else { if (want_pg) UWU_LN_BWD_S(8, true); else UWU_LN_BWD_S(8, false); }
#undef UWU_LN_BWD_S
        UWU_CHECK_LAUNCH();
        return UWU_OK;
    }
    if (want_pg && C <= 192 * 8) {
        const int threads = ((C / 8 + 31) / 32) * 32;
        // at most one resident wave (register-limited blocks per SM): the kernel strides over the rows, so for the 160-thread
        // blocks of C = 1280 a grid of 6 per SM ran as one full wave plus a half-empty one (62 -> 52 us)
        int per_sm = 65536 / (threads * 84);
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        int grid2 = ln_bwd_grid(M);
        if (grid2 > per_sm * sm_count()) grid2 = per_sm * sm_count();
        ln_bwd_cols_kernel<2><<<grid2, threads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                             reinterpret_cast<const __nv_bfloat16*>(dy), M, C, gamma, stats,
                                                             reinterpret_cast<const __nv_bfloat16*>(dres),
                                                             reinterpret_cast<__nv_bfloat16*>(dx), workspace);
        UWU_CHECK_LAUNCH();
        if (!accumulate) {
            if (dgamma) UWU_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, stream));
            if (dbeta) UWU_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, stream));
        }
        int slices = 16;
        if (slices > grid2) slices = grid2;
        const int p_per = (grid2 + slices - 1) / slices;
        slices = (grid2 + p_per - 1) / p_per;
        ln_param_reduce_kernel<<<dim3((2 * C + 127) / 128, slices), 128, 0, stream>>>(workspace, grid2, C, p_per, dgamma, dbeta);
        UWU_CHECK_LAUNCH();
        return UWU_OK;
    }
    const int grid = ln_bwd_grid(M);
    const int cv = C / 8;
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    const auto* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
    const auto* rp = reinterpret_cast<const __nv_bfloat16*>(dres);
    auto* dxp = reinterpret_cast<__nv_bfloat16*>(dx);
    float* pg = want_pg ? workspace : nullptr;
    const size_t sm = 2 * C * sizeof(float);
#define UWU_LN_BWD(V)                                                                                          \
    do {                                                                                                       \
        if (want_pg)                                                                                           \
            ln_bwd_kernel<V, true><<<grid, 128, sm, stream>>>(xp, dyp, M, C, gamma, stats, rp, dxp, pg);      \
        else                                                                                                   \
            ln_bwd_kernel<V, false><<<grid, 128, sm, stream>>>(xp, dyp, M, C, gamma, stats, rp, dxp, pg);     \
    } while (0)
    if (cv <= 32) UWU_LN_BWD(1);
    else if (cv <= 64) UWU_LN_BWD(2);
    else if (cv <= 96) UWU_LN_BWD(3);
    else UWU_LN_BWD(5);
#undef UWU_LN_BWD
    UWU_CHECK_LAUNCH();
    if (want_pg) {
        if (!accumulate) {
            if (dgamma) UWU_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, stream));
            if (dbeta) UWU_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, stream));
        }
        int slices = 16;
        if (slices > grid) slices = grid;
        const int p_per = (grid + slices - 1) / slices;
        slices = (grid + p_per - 1) / p_per;
        ln_param_reduce_kernel<<<dim3((2 * C + 127) / 128, slices), 128, 0, stream>>>(pg, grid, C, p_per, dgamma, dbeta);
        UWU_CHECK_LAUNCH();
    }
    return UWU_OK;
}
