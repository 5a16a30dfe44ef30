// Flash attention forward / backward for head dims above 64 (SD-1.5: 80 / 160, DiT-XL/2: 72).
//
// replaces: F.scaled_dot_product_attention in diffusers AttnProcessor2_0 (in-tree copy of the flow:
//           /root/reference/src/duwu/modules/rope_unet.py:76-175, SDPA call :151) for the UNets whose heads are not 64 wide.
//
// The d = 64 heads of the SDXL path run on the tcgen05 / TMEM kernels of attn.cu.  The shapes here carry a few percent of
// their models' FLOPs (SD-1.5: 8 heads of 40..160; DiT: 16 heads of 72, L = 256) and their widths do not tile the 128-byte
// swizzle atoms those kernels are built on, so they run on warp-level mma.sync.m16n8k16 tiles instead:
//   * head dim d (multiple of 8, 64 < d <= 160; narrower heads run zero-padded on the tcgen05 kernels) is zero-padded in
//     shared memory to DP in {80, 128, 160};
//   * forward:   block = 64 query rows (16 per warp), streams 64-key tiles, online softmax in registers, P stays in
//                registers as the A operand of P.V;
//   * backward:  two deterministic passes (no atomics).  dK/dV pass: block = 64 keys, computes S^T = K Q^T and dP^T = V dO^T
//                so P^T / dS^T come out directly in A-operand layout; dQ pass: block = 64 query rows, same shape as forward.
// Layouts are those of uwu_attn_fwd / uwu_attn_bwd (include/uwu_b200.h).
#include <cstdlib>
#include "api_internal.h"
#include "common.cuh"

namespace uwu {
namespace {

constexpr float LOG2E = 1.4426950408889634f;

constexpr int TQ = 64;   // query rows per block (forward, dQ pass)
constexpr int TK = 64;   // keys per tile / per block (dK,dV pass)
constexpr int NT = 128;  // threads per block: 4 warps x 16 rows
constexpr float LN2F = 0.6931471805599453f;

struct AnyArgs {
    const __nv_bfloat16 *q, *k, *v, *o, *dout;
    __nv_bfloat16 *out, *dq, *dk, *dv;
    float* lse;         // forward: written (natural log)
    const float* lse2;  // backward: lse * log2(e)
    const float* delta;
    long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
    int B, heads, Lq, Lk, Lq_pad, d;
    float scale, scale_log2;
    int causal;               // masked short forward only: key j is visible to query i iff j <= i
    const int32_t* key_mask;  // masked short forward only: optional [B, Lk], 0 = padding key (never attended)
};

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rows [row0, row0 + rows) of a [L, ld] matrix (already offset to batch / head) -> s[rows][DP + 8]; zero outside L x d
template <int DP>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, long long ld, int row0, int L, int d,
                                          int rows) {
    constexpr int LDS = DP + 8;
    constexpr int CH = DP / 8;
    for (int i = threadIdx.x; i < rows * CH; i += blockDim.x) {
        const int r = i / CH, c = i - r * CH;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row0 + r < L && c * 8 < d) v = __ldg(reinterpret_cast<const uint4*>(g + (long long)(row0 + r) * ld + c * 8));
        *reinterpret_cast<uint4*>(s + r * LDS + c * 8) = v;
    }
}

// same, through cp.async (16-byte chunks, zero-filled outside L x d); the caller commits / waits the group
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int DP>
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* s, const __nv_bfloat16* g, long long ld, int row0, int L, int d,
                                                int rows) {
    constexpr int LDS = DP + 8;
    constexpr int CH = DP / 8;
    const uint32_t sbase = smem_addr(s);
    for (int i = threadIdx.x; i < rows * CH; i += blockDim.x) {
        const int r = i / CH, c = i - r * CH;
        const bool ok = row0 + r < L && c * 8 < d;
        const __nv_bfloat16* src = ok ? g + (long long)(row0 + r) * ld + c * 8 : g;
        cp_async16(sbase + (uint32_t)((r * LDS + c * 8) * 2), src, ok ? 16 : 0);
    }
}

// acc[NTL][4] (16 rows x 8*NTL cols) += A[16 x DP] * B[8*NTL x DP]^T, both row-major in shared memory with stride DP + 8.
// a_addr: this lane's ldmatrix address into the 16-row A slab; b_base: byte address of row 0 of B.
template <int DP, int NTL>
__device__ __forceinline__ void gemm_nt(float (&acc)[NTL][4], uint32_t a_addr, uint32_t b_base, int lane) {
    constexpr int LDS = DP + 8;
    const uint32_t b_lane = b_base + (uint32_t)((((lane & 7) + ((lane >> 4) << 3)) * LDS + (((lane >> 3) & 1) << 3)) * 2);
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks) {
        uint32_t af[4];
        ldsm4(af, a_addr + ks * 32);
#pragma unroll
        for (int np = 0; np < NTL / 2; ++np) {
            uint32_t bf[4];
            ldsm4(bf, b_lane + (uint32_t)((np * 16 * LDS + ks * 16) * 2));
            mma16816(acc[2 * np], af, bf[0], bf[1]);
            mma16816(acc[2 * np + 1], af, bf[2], bf[3]);
        }
    }
}

// acc[DP/8][4] (16 rows x DP cols) += P[16 x 8*NTL] (registers, C-fragment layout) * B[8*NTL x DP] (row-major in smem)
template <int DP, int NTL>
__device__ __forceinline__ void gemm_pn(float (&acc)[DP / 8][4], const float (&p)[NTL][4], uint32_t b_base, int lane) {
    constexpr int LDS = DP + 8;
    const uint32_t b_lane = b_base + (uint32_t)(((lane & 15) * LDS + ((lane >> 4) << 3)) * 2);
#pragma unroll
    for (int kk = 0; kk < NTL / 2; ++kk) {
        uint32_t pa[4];
        pa[0] = pack2(p[2 * kk][0], p[2 * kk][1]);
        pa[1] = pack2(p[2 * kk][2], p[2 * kk][3]);
        pa[2] = pack2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        pa[3] = pack2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < DP / 16; ++dp) {
            uint32_t bf[4];
            ldsm4t(bf, b_lane + (uint32_t)((kk * 16 * LDS + dp * 16) * 2));
            mma16816(acc[2 * dp], pa, bf[0], bf[1]);
            mma16816(acc[2 * dp + 1], pa, bf[2], bf[3]);
        }
    }
}

// store a 16 x DP C-fragment slab as bf16 (cols < d, rows < L)
template <int DP>
__device__ __forceinline__ void store_slab(const float (&acc)[DP / 8][4], float mul0, float mul1, __nv_bfloat16* g,
                                           long long ld, int row, int L, int d, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) {
        const int col = nt * 8 + 2 * t;
        if (col < d) {
            if (row < L)
                *reinterpret_cast<__nv_bfloat162*>(g + (long long)row * ld + col) =
                    __floats2bfloat162_rn(acc[nt][0] * mul0, acc[nt][1] * mul0);
            if (row + 8 < L)
                *reinterpret_cast<__nv_bfloat162*>(g + (long long)(row + 8) * ld + col) =
                    __floats2bfloat162_rn(acc[nt][2] * mul1, acc[nt][3] * mul1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(NT) attn_any_fwd_kernel(const AnyArgs a) {
    constexpr int LDS = DP + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* Ks = Qs + TQ * LDS;        // two stages
    __nv_bfloat16* Vs = Ks + 2 * TK * LDS;    // two stages
    const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* qg = a.q + (long long)b * a.Lq * a.ldq + (long long)h * a.d;
    const __nv_bfloat16* kg = a.k + (long long)b * a.Lk * a.ldk + (long long)h * a.d;
    const __nv_bfloat16* vg = a.v + (long long)b * a.Lk * a.ldv + (long long)h * a.d;
    load_tile_async<DP>(Qs, qg, a.ldq, q0, a.Lq, a.d, TQ);
    load_tile_async<DP>(Ks, kg, a.ldk, 0, a.Lk, a.d, TK);
    load_tile_async<DP>(Vs, vg, a.ldv, 0, a.Lk, a.d, TK);
    cp_async_commit();
    const uint32_t q_addr = smem_addr(Qs) + (uint32_t)(((warp * 16 + (lane & 15)) * LDS + ((lane >> 4) << 3)) * 2);

    float o[DP / 8][4];
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    int buf = 0;
    for (int k0 = 0; k0 < a.Lk; k0 += TK, buf ^= 1) {
        if (k0 + TK < a.Lk) {  // prefetch the next key tile into the other stage while this one is consumed
            load_tile_async<DP>(Ks + (buf ^ 1) * TK * LDS, kg, a.ldk, k0 + TK, a.Lk, a.d, TK);
            load_tile_async<DP>(Vs + (buf ^ 1) * TK * LDS, vg, a.ldv, k0 + TK, a.Lk, a.d, TK);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t k_base = smem_addr(Ks + buf * TK * LDS), v_base = smem_addr(Vs + buf * TK * LDS);
        float s[TK / 8][4];
#pragma unroll
        for (int i = 0; i < TK / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        gemm_nt<DP, TK / 8>(s, q_addr, k_base, lane);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < TK / 8; ++nt) {
            const int col = k0 + nt * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v = (col + (e & 1) < a.Lk) ? s[nt][e] * a.scale_log2 : -INFINITY;
                s[nt][e] = v;
            }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile has at least one valid key
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < TK / 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - mn0);
            s[nt][1] = exp2f(s[nt][1] - mn0);
            s[nt][2] = exp2f(s[nt][2] - mn1);
            s[nt][3] = exp2f(s[nt][3] - mn1);
            rs0 += s[nt][0] + s[nt][1];
            rs1 += s[nt][2] + s[nt][3];
        }
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
#pragma unroll
        for (int i = 0; i < DP / 8; ++i) {
            o[i][0] *= c0;
            o[i][1] *= c0;
            o[i][2] *= c1;
            o[i][3] *= c1;
        }
        gemm_pn<DP, TK / 8>(o, s, v_base, lane);
        __syncthreads();  // every warp is done with this stage before the prefetch after next overwrites it
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const int row = q0 + warp * 16 + g;
    __nv_bfloat16* og = a.out + (long long)b * a.Lq * a.ldo + (long long)h * a.d;
    store_slab<DP>(o, 1.f / l0, 1.f / l1, og, a.ldo, row, a.Lq, a.d, lane);
    if (t == 0) {
        float* lp = a.lse + ((size_t)b * a.heads + h) * a.Lq_pad;
        if (row < a.Lq) lp[row] = (m0 + log2f(l0)) * LN2F;
        if (row + 8 < a.Lq) lp[row + 8] = (m1 + log2f(l1)) * LN2F;
    }
}

// ------------------------------------------------------------------------------------------------
// delta[b,h,q] = sum_d O[q,d] dO[q,d];  lse2 = lse * log2(e)
__global__ void attn_any_prep_kernel(const AnyArgs a, float* __restrict__ lse2, float* __restrict__ delta) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)a.B * a.Lq * a.heads;
    if (idx >= total) return;
    const int h = (int)(idx % a.heads);
    const long long bq = idx / a.heads;
    const int q = (int)(bq % a.Lq);
    const int b = (int)(bq / a.Lq);
    const __nv_bfloat16* op = a.o + bq * a.ldo + (long long)h * a.d;
    const __nv_bfloat16* dp = a.dout + bq * a.lddo + (long long)h * a.d;
    float acc = 0.f;
    for (int c = 0; c < a.d; c += 8) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(op + c));
        const uint4 y = __ldg(reinterpret_cast<const uint4*>(dp + c));
        const float2 x0 = unpack_bf16(x.x), x1 = unpack_bf16(x.y), x2 = unpack_bf16(x.z), x3 = unpack_bf16(x.w);
        const float2 y0 = unpack_bf16(y.x), y1 = unpack_bf16(y.y), y2 = unpack_bf16(y.z), y3 = unpack_bf16(y.w);
        acc += x0.x * y0.x + x0.y * y0.y + x1.x * y1.x + x1.y * y1.y + x2.x * y2.x + x2.y * y2.y + x3.x * y3.x + x3.y * y3.y;
    }
    const size_t oi = ((size_t)b * a.heads + h) * a.Lq_pad + q;
    delta[oi] = acc;
    lse2[oi] = a.lse[oi] * LOG2E;
}

// ------------------------------------------------------------------------------------------------
// dK, dV for one 64-key tile; QT query rows per inner step.
template <int DP, int QT>
__global__ void __launch_bounds__(NT) attn_any_bwd_kv_kernel(const AnyArgs a) {
    constexpr int LDS = DP + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* Vs = Ks + TK * LDS;
    __nv_bfloat16* Qs = Vs + TK * LDS;        // two stages
    __nv_bfloat16* dOs = Qs + 2 * QT * LDS;   // two stages
    float* lses = reinterpret_cast<float*>(dOs + 2 * QT * LDS);  // [2][QT]
    float* dls = lses + 2 * QT;                                  // [2][QT]
    const int k0 = blockIdx.x * TK, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* qg = a.q + (long long)b * a.Lq * a.ldq + (long long)h * a.d;
    const __nv_bfloat16* kg = a.k + (long long)b * a.Lk * a.ldk + (long long)h * a.d;
    const __nv_bfloat16* vg = a.v + (long long)b * a.Lk * a.ldv + (long long)h * a.d;
    const __nv_bfloat16* dog = a.dout + (long long)b * a.Lq * a.lddo + (long long)h * a.d;
    const float* lse2 = a.lse2 + ((size_t)b * a.heads + h) * a.Lq_pad;
    const float* delta = a.delta + ((size_t)b * a.heads + h) * a.Lq_pad;
    auto stage_in = [&](int q0, int st) {
        load_tile_async<DP>(Qs + st * QT * LDS, qg, a.ldq, q0, a.Lq, a.d, QT);
        load_tile_async<DP>(dOs + st * QT * LDS, dog, a.lddo, q0, a.Lq, a.d, QT);
        if (threadIdx.x < QT) {
            const bool ok = q0 + threadIdx.x < a.Lq;
            lses[st * QT + threadIdx.x] = ok ? lse2[q0 + threadIdx.x] : INFINITY;  // exp2(s - inf) = 0 masks padded queries
            dls[st * QT + threadIdx.x] = ok ? delta[q0 + threadIdx.x] : 0.f;
        }
    };
    load_tile_async<DP>(Ks, kg, a.ldk, k0, a.Lk, a.d, TK);
    load_tile_async<DP>(Vs, vg, a.ldv, k0, a.Lk, a.d, TK);
    stage_in(0, 0);
    cp_async_commit();
    const uint32_t a_off = (uint32_t)(((warp * 16 + (lane & 15)) * LDS + ((lane >> 4) << 3)) * 2);
    const uint32_t k_addr = smem_addr(Ks) + a_off, v_addr = smem_addr(Vs) + a_off;
    const int key = k0 + warp * 16 + g;
    const bool kv0 = key < a.Lk, kv1 = key + 8 < a.Lk;

    float dk[DP / 8][4], dv[DP / 8][4];
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    int buf = 0;
    for (int q0 = 0; q0 < a.Lq; q0 += QT, buf ^= 1) {
        if (q0 + QT < a.Lq) {
            stage_in(q0 + QT, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t q_base = smem_addr(Qs + buf * QT * LDS), do_base = smem_addr(dOs + buf * QT * LDS);
        const float* lsb = lses + buf * QT;
        const float* dlb = dls + buf * QT;
        float st[QT / 8][4], dpt[QT / 8][4];
#pragma unroll
        for (int i = 0; i < QT / 8; ++i) {
            st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
            dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
        }
        gemm_nt<DP, QT / 8>(st, k_addr, q_base, lane);     // S^T  = K Q^T
        gemm_nt<DP, QT / 8>(dpt, v_addr, do_base, lane);   // dP^T = V dO^T
#pragma unroll
        for (int nt = 0; nt < QT / 8; ++nt) {
            const int qi = nt * 8 + 2 * t;
            const float ls0 = lsb[qi], ls1 = lsb[qi + 1], d0 = dlb[qi], d1 = dlb[qi + 1];
            const float p0 = kv0 ? exp2f(st[nt][0] * a.scale_log2 - ls0) : 0.f;
            const float p1 = kv0 ? exp2f(st[nt][1] * a.scale_log2 - ls1) : 0.f;
            const float p2 = kv1 ? exp2f(st[nt][2] * a.scale_log2 - ls0) : 0.f;
            const float p3 = kv1 ? exp2f(st[nt][3] * a.scale_log2 - ls1) : 0.f;
            st[nt][0] = p0;
            st[nt][1] = p1;
            st[nt][2] = p2;
            st[nt][3] = p3;
            dpt[nt][0] = p0 * (dpt[nt][0] - d0) * a.scale;
            dpt[nt][1] = p1 * (dpt[nt][1] - d1) * a.scale;
            dpt[nt][2] = p2 * (dpt[nt][2] - d0) * a.scale;
            dpt[nt][3] = p3 * (dpt[nt][3] - d1) * a.scale;
        }
        gemm_pn<DP, QT / 8>(dv, st, do_base, lane);   // dV += P^T dO
        gemm_pn<DP, QT / 8>(dk, dpt, q_base, lane);   // dK += dS^T Q
        __syncthreads();
    }
    __nv_bfloat16* dkg = a.dk + (long long)b * a.Lk * a.lddk + (long long)h * a.d;
    __nv_bfloat16* dvg = a.dv + (long long)b * a.Lk * a.lddv + (long long)h * a.d;
    store_slab<DP>(dk, 1.f, 1.f, dkg, a.lddk, key, a.Lk, a.d, lane);
    store_slab<DP>(dv, 1.f, 1.f, dvg, a.lddv, key, a.Lk, a.d, lane);
}

// ------------------------------------------------------------------------------------------------
// dQ for one 64-row query tile.
template <int DP>
__global__ void __launch_bounds__(NT) attn_any_bwd_q_kernel(const AnyArgs a) {
    constexpr int LDS = DP + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* dOs = Qs + TQ * LDS;
    __nv_bfloat16* Ks = dOs + TQ * LDS;       // two stages
    __nv_bfloat16* Vs = Ks + 2 * TK * LDS;    // two stages
    const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* qg = a.q + (long long)b * a.Lq * a.ldq + (long long)h * a.d;
    const __nv_bfloat16* kg = a.k + (long long)b * a.Lk * a.ldk + (long long)h * a.d;
    const __nv_bfloat16* vg = a.v + (long long)b * a.Lk * a.ldv + (long long)h * a.d;
    const __nv_bfloat16* dog = a.dout + (long long)b * a.Lq * a.lddo + (long long)h * a.d;
    load_tile_async<DP>(Qs, qg, a.ldq, q0, a.Lq, a.d, TQ);
    load_tile_async<DP>(dOs, dog, a.lddo, q0, a.Lq, a.d, TQ);
    load_tile_async<DP>(Ks, kg, a.ldk, 0, a.Lk, a.d, TK);
    load_tile_async<DP>(Vs, vg, a.ldv, 0, a.Lk, a.d, TK);
    cp_async_commit();
    const uint32_t a_off = (uint32_t)(((warp * 16 + (lane & 15)) * LDS + ((lane >> 4) << 3)) * 2);
    const uint32_t q_addr = smem_addr(Qs) + a_off, do_addr = smem_addr(dOs) + a_off;
    const int row = q0 + warp * 16 + g;
    const size_t sidx = ((size_t)b * a.heads + h) * a.Lq_pad;
    const float ls0 = row < a.Lq ? a.lse2[sidx + row] : INFINITY, ls1 = row + 8 < a.Lq ? a.lse2[sidx + row + 8] : INFINITY;
    const float d0 = row < a.Lq ? a.delta[sidx + row] : 0.f, d1 = row + 8 < a.Lq ? a.delta[sidx + row + 8] : 0.f;

    float dq[DP / 8][4];
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    int buf = 0;
    for (int k0 = 0; k0 < a.Lk; k0 += TK, buf ^= 1) {
        if (k0 + TK < a.Lk) {
            load_tile_async<DP>(Ks + (buf ^ 1) * TK * LDS, kg, a.ldk, k0 + TK, a.Lk, a.d, TK);
            load_tile_async<DP>(Vs + (buf ^ 1) * TK * LDS, vg, a.ldv, k0 + TK, a.Lk, a.d, TK);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t k_base = smem_addr(Ks + buf * TK * LDS), v_base = smem_addr(Vs + buf * TK * LDS);
        float s[TK / 8][4], dp[TK / 8][4];
#pragma unroll
        for (int i = 0; i < TK / 8; ++i) {
            s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
        }
        gemm_nt<DP, TK / 8>(s, q_addr, k_base, lane);     // S  = Q K^T
        gemm_nt<DP, TK / 8>(dp, do_addr, v_base, lane);   // dP = dO V^T
#pragma unroll
        for (int nt = 0; nt < TK / 8; ++nt) {
            const int col = k0 + nt * 8 + 2 * t;
            const bool c0 = col < a.Lk, c1 = col + 1 < a.Lk;
            const float p0 = c0 ? exp2f(s[nt][0] * a.scale_log2 - ls0) : 0.f;
            const float p1 = c1 ? exp2f(s[nt][1] * a.scale_log2 - ls0) : 0.f;
            const float p2 = c0 ? exp2f(s[nt][2] * a.scale_log2 - ls1) : 0.f;
            const float p3 = c1 ? exp2f(s[nt][3] * a.scale_log2 - ls1) : 0.f;
            dp[nt][0] = p0 * (dp[nt][0] - d0) * a.scale;
            dp[nt][1] = p1 * (dp[nt][1] - d0) * a.scale;
            dp[nt][2] = p2 * (dp[nt][2] - d1) * a.scale;
            dp[nt][3] = p3 * (dp[nt][3] - d1) * a.scale;
        }
        gemm_pn<DP, TK / 8>(dq, dp, k_base, lane);   // dQ += dS K
        __syncthreads();
    }
    __nv_bfloat16* dqg = a.dq + (long long)b * a.Lq * a.lddq + (long long)h * a.d;
    store_slab<DP>(dq, 1.f, 1.f, dqg, a.lddq, row, a.Lq, a.d, lane);
}


// ================================================================================================
// Short-key attention (cross-attention on the 77-token text context: Lk <= 128, head dim <= 64).
// All keys / values of a (batch, head) stay in shared memory; every warp walks its own 16-row query tiles with a private
// double-buffered cp.async pipeline, so there is no block barrier after the prologue and the softmax is a single pass.
// The tcgen05 kernels spend most of a CTA's life in prologue / epilogue on these shapes (one 128-key tile per CTA).
//   forward:   S = Q K^T -> P -> O = P V, O staged through the consumed Q buffer for coalesced 16-byte stores
//   backward:  dQ kernel (same walk: S, dP = dO V^T, delta = rowsum(P o dP), dS, dQ = dS K; writes lse2 / delta), then the
//              dK/dV kernel: one warp per 16 keys, block-level double-buffered 64-query tiles, S^T = K Q^T and dP^T = V dO^T
// ================================================================================================
constexpr int SH_DP = 64, SH_LDS = SH_DP + 8, SH_ROWS = 256;  // query rows per block: 4 warps x 4 tiles of 16

struct ShortWalk {
    __nv_bfloat16* buf[2];  // warp-private tile buffers
};

// one 16-row tile of a [L, ld] matrix -> warp-private smem [16][72] (zero outside L x d); 4 chunks of 16 bytes per lane
__device__ __forceinline__ void load_rows16_async(__nv_bfloat16* s, const __nv_bfloat16* g, long long ld, int row0, int L, int d,
                                                  int lane) {
    const uint32_t sbase = smem_addr(s);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int idx = lane + 32 * j, r = idx >> 3, c = idx & 7;
        const bool ok = row0 + r < L && c * 8 < d;
        const __nv_bfloat16* src = ok ? g + (long long)(row0 + r) * ld + c * 8 : g;
        cp_async16(sbase + (uint32_t)((r * SH_LDS + c * 8) * 2), src, ok ? 16 : 0);
    }
}
// C-fragment slab [16 x 64] fp32 -> bf16 in the warp's staging tile -> coalesced 16-byte row stores
__device__ __forceinline__ void store_rows16(const float (&acc)[8][4], float m0, float m1, __nv_bfloat16* stage,
                                             __nv_bfloat16* g, long long ld, int row0, int L, int d, int lane) {
    const int gq = lane >> 2, t = lane & 3;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<__nv_bfloat162*>(stage + gq * SH_LDS + nt * 8 + 2 * t) = __floats2bfloat162_rn(acc[nt][0] * m0, acc[nt][1] * m0);
        *reinterpret_cast<__nv_bfloat162*>(stage + (gq + 8) * SH_LDS + nt * 8 + 2 * t) =
            __floats2bfloat162_rn(acc[nt][2] * m1, acc[nt][3] * m1);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int idx = lane + 32 * j, r = idx >> 3, c = idx & 7;
        if (row0 + r < L && c * 8 < d)
            *reinterpret_cast<uint4*>(g + (long long)(row0 + r) * ld + c * 8) = *reinterpret_cast<const uint4*>(stage + r * SH_LDS + c * 8);
    }
    __syncwarp();
}

template <int NK8, bool MASKED>
__global__ void __launch_bounds__(128) attn_short_fwd_kernel(const AnyArgs a) {
    pdl_trigger();
    constexpr int LKP = NK8 * 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* Vs = Ks + LKP * SH_LDS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    __nv_bfloat16* Qw = Vs + LKP * SH_LDS + warp * 2 * 16 * SH_LDS;  // [2][16][72]
    const int chunk0 = blockIdx.x * SH_ROWS, h = blockIdx.y, b = blockIdx.z;
    const __nv_bfloat16* qg = a.q + (long long)b * a.Lq * a.ldq + (long long)h * a.d;
    const __nv_bfloat16* kg = a.k + (long long)b * a.Lk * a.ldk + (long long)h * a.d;
    const __nv_bfloat16* vg = a.v + (long long)b * a.Lk * a.ldv + (long long)h * a.d;
    __nv_bfloat16* og = a.out + (long long)b * a.Lq * a.ldo + (long long)h * a.d;
    float* lp = a.lse + ((size_t)b * a.heads + h) * a.Lq_pad;
    load_tile_async<SH_DP>(Ks, kg, a.ldk, 0, a.Lk, a.d, LKP);
    load_tile_async<SH_DP>(Vs, vg, a.ldv, 0, a.Lk, a.d, LKP);
    load_rows16_async(Qw, qg, a.ldq, chunk0 + warp * 16, a.Lq, a.d, lane);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const uint32_t k_base = smem_addr(Ks), v_base = smem_addr(Vs);
    const uint32_t a_off = (uint32_t)(((lane & 15) * SH_LDS + ((lane >> 4) << 3)) * 2);
    int buf = 0;
    for (int i = 0; i < SH_ROWS / 64; ++i, buf ^= 1) {
        const int row0 = chunk0 + (warp + 4 * i) * 16;
        if (row0 >= a.Lq) break;
        if (i + 1 < SH_ROWS / 64) load_rows16_async(Qw + (buf ^ 1) * 16 * SH_LDS, qg, a.ldq, row0 + 64, a.Lq, a.d, lane);
        cp_async_commit();
        __nv_bfloat16* qb = Qw + buf * 16 * SH_LDS;
        float s[NK8][4];
#pragma unroll
        for (int n = 0; n < NK8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        gemm_nt<SH_DP, NK8>(s, smem_addr(qb) + a_off, k_base, lane);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NK8; ++nt) {
            const int col = nt * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                bool ok = col + (e & 1) < a.Lk;
                if (MASKED) {  // CLIP text towers: causal mask + tokenizer padding mask (transformers create_causal_mask)
                    const int cj = col + (e & 1), ri = row0 + g + ((e >> 1) << 3);
                    ok = ok && (!a.causal || cj <= ri) && (a.key_mask == nullptr || a.key_mask[(size_t)b * a.Lk + cj] != 0);
                }
                s[nt][e] = ok ? s[nt][e] * a.scale_log2 : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < NK8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - mx0);
            s[nt][1] = exp2f(s[nt][1] - mx0);
            s[nt][2] = exp2f(s[nt][2] - mx1);
            s[nt][3] = exp2f(s[nt][3] - mx1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        float o[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
        gemm_pn<SH_DP, NK8>(o, s, v_base, lane);
        store_rows16(o, 1.f / l0, 1.f / l1, qb, og, a.ldo, row0, a.Lq, a.d, lane);
        if (t == 0) {
            if (row0 + g < a.Lq) lp[row0 + g] = (mx0 + log2f(l0)) * LN2F;
            if (row0 + g + 8 < a.Lq) lp[row0 + g + 8] = (mx1 + log2f(l1)) * LN2F;
        }
        cp_async_wait<0>();
        __syncwarp();
    }
}

template <int NK8>
__global__ void __launch_bounds__(128) attn_short_bwd_q_kernel(const AnyArgs a, float* __restrict__ lse2_out,
                                                               float* __restrict__ delta_out) {
    constexpr int LKP = NK8 * 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* Vs = Ks + LKP * SH_LDS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    __nv_bfloat16* Qw = Vs + LKP * SH_LDS + warp * 4 * 16 * SH_LDS;  // [2][16][72] Q then [2][16][72] dO
    __nv_bfloat16* Dw = Qw + 2 * 16 * SH_LDS;
    const int chunk0 = blockIdx.x * SH_ROWS, h = blockIdx.y, b = blockIdx.z;
    const __nv_bfloat16* qg = a.q + (long long)b * a.Lq * a.ldq + (long long)h * a.d;
    const __nv_bfloat16* kg = a.k + (long long)b * a.Lk * a.ldk + (long long)h * a.d;
    const __nv_bfloat16* vg = a.v + (long long)b * a.Lk * a.ldv + (long long)h * a.d;
    const __nv_bfloat16* dog = a.dout + (long long)b * a.Lq * a.lddo + (long long)h * a.d;
    __nv_bfloat16* dqg = a.dq + (long long)b * a.Lq * a.lddq + (long long)h * a.d;
    const size_t sidx = ((size_t)b * a.heads + h) * a.Lq_pad;
    load_tile_async<SH_DP>(Ks, kg, a.ldk, 0, a.Lk, a.d, LKP);
    load_tile_async<SH_DP>(Vs, vg, a.ldv, 0, a.Lk, a.d, LKP);
    load_rows16_async(Qw, qg, a.ldq, chunk0 + warp * 16, a.Lq, a.d, lane);
    load_rows16_async(Dw, dog, a.lddo, chunk0 + warp * 16, a.Lq, a.d, lane);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const uint32_t k_base = smem_addr(Ks), v_base = smem_addr(Vs);
    const uint32_t a_off = (uint32_t)(((lane & 15) * SH_LDS + ((lane >> 4) << 3)) * 2);
    int buf = 0;
    for (int i = 0; i < SH_ROWS / 64; ++i, buf ^= 1) {
        const int row0 = chunk0 + (warp + 4 * i) * 16;
        if (row0 >= a.Lq) break;
        if (i + 1 < SH_ROWS / 64) {
            load_rows16_async(Qw + (buf ^ 1) * 16 * SH_LDS, qg, a.ldq, row0 + 64, a.Lq, a.d, lane);
            load_rows16_async(Dw + (buf ^ 1) * 16 * SH_LDS, dog, a.lddo, row0 + 64, a.Lq, a.d, lane);
        }
        cp_async_commit();
        __nv_bfloat16* qb = Qw + buf * 16 * SH_LDS;
        __nv_bfloat16* db = Dw + buf * 16 * SH_LDS;
        const bool r0ok = row0 + g < a.Lq, r1ok = row0 + g + 8 < a.Lq;
        const float ls0 = r0ok ? a.lse[sidx + row0 + g] * LOG2E : INFINITY;
        const float ls1 = r1ok ? a.lse[sidx + row0 + g + 8] * LOG2E : INFINITY;
        float s[NK8][4], dp[NK8][4];
#pragma unroll
        for (int n = 0; n < NK8; ++n) {
            s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
            dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
        }
        gemm_nt<SH_DP, NK8>(s, smem_addr(qb) + a_off, k_base, lane);    // S  = Q K^T
        gemm_nt<SH_DP, NK8>(dp, smem_addr(db) + a_off, v_base, lane);   // dP = dO V^T
        float d0 = 0.f, d1 = 0.f;  // delta = rowsum(P o dP) = rowsum(dO o O)
#pragma unroll
        for (int nt = 0; nt < NK8; ++nt) {
            const int col = nt * 8 + 2 * t;
            const bool c0 = col < a.Lk, c1 = col + 1 < a.Lk;
            s[nt][0] = c0 ? exp2f(s[nt][0] * a.scale_log2 - ls0) : 0.f;
            s[nt][1] = c1 ? exp2f(s[nt][1] * a.scale_log2 - ls0) : 0.f;
            s[nt][2] = c0 ? exp2f(s[nt][2] * a.scale_log2 - ls1) : 0.f;
            s[nt][3] = c1 ? exp2f(s[nt][3] * a.scale_log2 - ls1) : 0.f;
            d0 += s[nt][0] * dp[nt][0] + s[nt][1] * dp[nt][1];
            d1 += s[nt][2] * dp[nt][2] + s[nt][3] * dp[nt][3];
        }
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
        d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
        d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
        d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
#pragma unroll
        for (int nt = 0; nt < NK8; ++nt) {
            dp[nt][0] = s[nt][0] * (dp[nt][0] - d0) * a.scale;
            dp[nt][1] = s[nt][1] * (dp[nt][1] - d0) * a.scale;
            dp[nt][2] = s[nt][2] * (dp[nt][2] - d1) * a.scale;
            dp[nt][3] = s[nt][3] * (dp[nt][3] - d1) * a.scale;
        }
        float dq[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
        gemm_pn<SH_DP, NK8>(dq, dp, k_base, lane);   // dQ = dS K
        store_rows16(dq, 1.f, 1.f, qb, dqg, a.lddq, row0, a.Lq, a.d, lane);
        if (t == 0) {
            if (r0ok) {
                lse2_out[sidx + row0 + g] = ls0;
                delta_out[sidx + row0 + g] = d0;
            }
            if (r1ok) {
                lse2_out[sidx + row0 + g + 8] = ls1;
                delta_out[sidx + row0 + g + 8] = d1;
            }
        }
        cp_async_wait<0>();
        __syncwarp();
    }
}

// dK / dV of the whole (batch, head) key set for queries [chunk * rows_per_chunk, ...): warp w owns keys [16 w, 16 w + 16).
// kAtomic: partial sums of several query chunks meet in an fp32 scratch [B, heads, 2, LKP, 64]; otherwise bf16 stores.
template <int NK8, bool kAtomic>
__global__ void __launch_bounds__(NK8 * 16) attn_short_bwd_kv_kernel(const AnyArgs a, int rows_per_chunk, float* __restrict__ scratch) {
    constexpr int LKP = NK8 * 8, QT = 64, DP = SH_DP, LDS = SH_LDS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* Vs = Ks + LKP * LDS;
    __nv_bfloat16* Qs = Vs + LKP * LDS;       // two stages
    __nv_bfloat16* dOs = Qs + 2 * QT * LDS;   // two stages
    float* lses = reinterpret_cast<float*>(dOs + 2 * QT * LDS);  // [2][QT]
    float* dls = lses + 2 * QT;
    const int h = blockIdx.y, b = blockIdx.z;
    const int qbeg = blockIdx.x * rows_per_chunk, qend = min(a.Lq, qbeg + rows_per_chunk);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* qg = a.q + (long long)b * a.Lq * a.ldq + (long long)h * a.d;
    const __nv_bfloat16* kg = a.k + (long long)b * a.Lk * a.ldk + (long long)h * a.d;
    const __nv_bfloat16* vg = a.v + (long long)b * a.Lk * a.ldv + (long long)h * a.d;
    const __nv_bfloat16* dog = a.dout + (long long)b * a.Lq * a.lddo + (long long)h * a.d;
    const float* lse2 = a.lse2 + ((size_t)b * a.heads + h) * a.Lq_pad;
    const float* delta = a.delta + ((size_t)b * a.heads + h) * a.Lq_pad;
    auto stage_in = [&](int q0, int st) {
        load_tile_async<DP>(Qs + st * QT * LDS, qg, a.ldq, q0, qend, a.d, QT);
        load_tile_async<DP>(dOs + st * QT * LDS, dog, a.lddo, q0, qend, a.d, QT);
        if (threadIdx.x < QT) {
            const bool ok = q0 + threadIdx.x < qend;
            lses[st * QT + threadIdx.x] = ok ? lse2[q0 + threadIdx.x] : INFINITY;
            dls[st * QT + threadIdx.x] = ok ? delta[q0 + threadIdx.x] : 0.f;
        }
    };
    load_tile_async<DP>(Ks, kg, a.ldk, 0, a.Lk, a.d, LKP);
    load_tile_async<DP>(Vs, vg, a.ldv, 0, a.Lk, a.d, LKP);
    stage_in(qbeg, 0);
    cp_async_commit();
    const uint32_t a_off = (uint32_t)(((warp * 16 + (lane & 15)) * LDS + ((lane >> 4) << 3)) * 2);
    const uint32_t k_addr = smem_addr(Ks) + a_off, v_addr = smem_addr(Vs) + a_off;
    const int key = warp * 16 + g;
    const bool kv0 = key < a.Lk, kv1 = key + 8 < a.Lk;
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    int buf = 0;
    for (int q0 = qbeg; q0 < qend; q0 += QT, buf ^= 1) {
        if (q0 + QT < qend) {
            stage_in(q0 + QT, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t q_base = smem_addr(Qs + buf * QT * LDS), do_base = smem_addr(dOs + buf * QT * LDS);
        const float* lsb = lses + buf * QT;
        const float* dlb = dls + buf * QT;
        float st[QT / 8][4], dpt[QT / 8][4];
#pragma unroll
        for (int i = 0; i < QT / 8; ++i) {
            st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
            dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
        }
        gemm_nt<DP, QT / 8>(st, k_addr, q_base, lane);     // S^T  = K Q^T
        gemm_nt<DP, QT / 8>(dpt, v_addr, do_base, lane);   // dP^T = V dO^T
#pragma unroll
        for (int nt = 0; nt < QT / 8; ++nt) {
            const int qi = nt * 8 + 2 * t;
            const float ls0 = lsb[qi], ls1 = lsb[qi + 1], d0 = dlb[qi], d1 = dlb[qi + 1];
            const float p0 = kv0 ? exp2f(st[nt][0] * a.scale_log2 - ls0) : 0.f;
            const float p1 = kv0 ? exp2f(st[nt][1] * a.scale_log2 - ls1) : 0.f;
            const float p2 = kv1 ? exp2f(st[nt][2] * a.scale_log2 - ls0) : 0.f;
            const float p3 = kv1 ? exp2f(st[nt][3] * a.scale_log2 - ls1) : 0.f;
            st[nt][0] = p0;
            st[nt][1] = p1;
            st[nt][2] = p2;
            st[nt][3] = p3;
            dpt[nt][0] = p0 * (dpt[nt][0] - d0) * a.scale;
            dpt[nt][1] = p1 * (dpt[nt][1] - d1) * a.scale;
            dpt[nt][2] = p2 * (dpt[nt][2] - d0) * a.scale;
            dpt[nt][3] = p3 * (dpt[nt][3] - d1) * a.scale;
        }
        gemm_pn<DP, QT / 8>(dv, st, do_base, lane);   // dV += P^T dO
        gemm_pn<DP, QT / 8>(dk, dpt, q_base, lane);   // dK += dS^T Q
        __syncthreads();
    }
    if (kAtomic) {
        float* sk = scratch + (((size_t)b * a.heads + h) * 2) * LKP * DP;
        float* sv = sk + LKP * DP;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = nt * 8 + 2 * t;
            if (kv0) {
                atomicAdd(sk + key * DP + col, dk[nt][0]);
                atomicAdd(sk + key * DP + col + 1, dk[nt][1]);
                atomicAdd(sv + key * DP + col, dv[nt][0]);
                atomicAdd(sv + key * DP + col + 1, dv[nt][1]);
            }
            if (kv1) {
                atomicAdd(sk + (key + 8) * DP + col, dk[nt][2]);
                atomicAdd(sk + (key + 8) * DP + col + 1, dk[nt][3]);
                atomicAdd(sv + (key + 8) * DP + col, dv[nt][2]);
                atomicAdd(sv + (key + 8) * DP + col + 1, dv[nt][3]);
            }
        }
    } else {
        __nv_bfloat16* dkg = a.dk + (long long)b * a.Lk * a.lddk + (long long)h * a.d;
        __nv_bfloat16* dvg = a.dv + (long long)b * a.Lk * a.lddv + (long long)h * a.d;
        store_slab<DP>(dk, 1.f, 1.f, dkg, a.lddk, key, a.Lk, a.d, lane);
        store_slab<DP>(dv, 1.f, 1.f, dvg, a.lddv, key, a.Lk, a.d, lane);
    }
}

// fp32 scratch [B, heads, 2, LKP, 64] -> bf16 dk / dv
__global__ void attn_short_kv_convert_kernel(const AnyArgs a, const float* __restrict__ scratch, int LKP) {
    const long long total = (long long)a.B * a.heads * a.Lk * (a.d / 8);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8 = (int)(idx % (a.d / 8));
    long long r = idx / (a.d / 8);
    const int key = (int)(r % a.Lk);
    r /= a.Lk;
    const int h = (int)(r % a.heads), b = (int)(r / a.heads);
    const float* sk = scratch + (((size_t)b * a.heads + h) * 2) * LKP * SH_DP + (size_t)key * SH_DP + c8 * 8;
    const float* sv = sk + (size_t)LKP * SH_DP;
    const float4 k0 = *reinterpret_cast<const float4*>(sk), k1 = *reinterpret_cast<const float4*>(sk + 4);
    const float4 v0 = *reinterpret_cast<const float4*>(sv), v1 = *reinterpret_cast<const float4*>(sv + 4);
    uint4 uk, uv;
    uk.x = pack2(k0.x, k0.y); uk.y = pack2(k0.z, k0.w); uk.z = pack2(k1.x, k1.y); uk.w = pack2(k1.z, k1.w);
    uv.x = pack2(v0.x, v0.y); uv.y = pack2(v0.z, v0.w); uv.z = pack2(v1.x, v1.y); uv.w = pack2(v1.z, v1.w);
    *reinterpret_cast<uint4*>(a.dk + ((long long)b * a.Lk + key) * a.lddk + (long long)h * a.d + c8 * 8) = uk;
    *reinterpret_cast<uint4*>(a.dv + ((long long)b * a.Lk + key) * a.lddv + (long long)h * a.d + c8 * 8) = uv;
}

template <int NK8, bool MASKED = false>
int launch_short_fwd(const AnyArgs& a, cudaStream_t stream) {
    constexpr int SMEM = (2 * NK8 * 8 + 4 * 2 * 16) * SH_LDS * 2;
    static bool attr = false;
    if (!attr) {
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_short_fwd_kernel<NK8, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr = true;
    }
    attn_short_fwd_kernel<NK8, MASKED><<<dim3((a.Lq + SH_ROWS - 1) / SH_ROWS, a.heads, a.B), 128, SMEM, stream>>>(a);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

template <int NK8>
int launch_short_bwd(AnyArgs& a, float* workspace, cudaStream_t stream) {
    constexpr int LKP = NK8 * 8;
    constexpr int SMEM_Q = (2 * LKP + 4 * 4 * 16) * SH_LDS * 2;
    constexpr int SMEM_KV = (2 * LKP + 4 * 64) * SH_LDS * 2 + 4 * 64 * 4;
    static bool attr = false;
    if (!attr) {
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_short_bwd_q_kernel<NK8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_Q));
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_short_bwd_kv_kernel<NK8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV));
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_short_bwd_kv_kernel<NK8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV));
        attr = true;
    }
    const long long rows = (long long)a.B * a.heads * a.Lq_pad;
    float* lse2 = workspace;
    float* delta = workspace + rows;
    float* scratch = workspace + 2 * rows;
    a.lse2 = lse2;
    a.delta = delta;
    attn_short_bwd_q_kernel<NK8><<<dim3((a.Lq + SH_ROWS - 1) / SH_ROWS, a.heads, a.B), 128, SMEM_Q, stream>>>(a, lse2, delta);
    UWU_CHECK_LAUNCH();
    // 256-query chunks keep ~9 blocks per SM in flight at the step shapes (B16 h20 L1024: 1280 blocks); their partial
    // dK / dV meet through fp32 atomics (40 KB per block)
    const int nchunk = (a.Lq + 255) / 256;
    if (nchunk == 1) {
        attn_short_bwd_kv_kernel<NK8, false><<<dim3(1, a.heads, a.B), NK8 * 16, SMEM_KV, stream>>>(a, a.Lq, nullptr);
        UWU_CHECK_LAUNCH();
    } else {
        const size_t n = (size_t)a.B * a.heads * 2 * LKP * SH_DP;
        UWU_CHECK_CUDA(cudaMemsetAsync(scratch, 0, n * sizeof(float), stream));
        const int rows_per = ((a.Lq + nchunk - 1) / nchunk + 63) / 64 * 64;
        attn_short_bwd_kv_kernel<NK8, true><<<dim3(nchunk, a.heads, a.B), NK8 * 16, SMEM_KV, stream>>>(a, rows_per, scratch);
        UWU_CHECK_LAUNCH();
        const long long total = (long long)a.B * a.heads * a.Lk * (a.d / 8);
        attn_short_kv_convert_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a, scratch, LKP);
        UWU_CHECK_LAUNCH();
    }
    return UWU_OK;
}

template <int DP>
int launch_fwd(const AnyArgs& a, cudaStream_t stream) {
    constexpr int SMEM = (TQ + 4 * TK) * (DP + 8) * 2;
    static bool attr = false;
    if (!attr) {
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_any_fwd_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr = true;
    }
    attn_any_fwd_kernel<DP><<<dim3((a.Lq + TQ - 1) / TQ, a.heads, a.B), NT, SMEM, stream>>>(a);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

template <int DP, int QT>
int launch_bwd(const AnyArgs& a, cudaStream_t stream) {
    constexpr int SMEM_KV = (2 * TK + 4 * QT) * (DP + 8) * 2 + 4 * QT * 4;
    constexpr int SMEM_Q = (2 * TQ + 4 * TK) * (DP + 8) * 2;
    static bool attr = false;
    if (!attr) {
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_any_bwd_kv_kernel<DP, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV));
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_any_bwd_q_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_Q));
        attr = true;
    }
    attn_any_bwd_kv_kernel<DP, QT><<<dim3((a.Lk + TK - 1) / TK, a.heads, a.B), NT, SMEM_KV, stream>>>(a);
    UWU_CHECK_LAUNCH();
    attn_any_bwd_q_kernel<DP><<<dim3((a.Lq + TQ - 1) / TQ, a.heads, a.B), NT, SMEM_Q, stream>>>(a);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

int check_any(const AnyArgs& a, const char* who) {
    UWU_CHECK_ARG(a.d > 0 && a.d % 8 == 0 && a.d <= 160, "%s: head_dim %d unsupported (multiple of 8, <= 160)", who, a.d);
    UWU_CHECK_ARG(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0, "%s: leading dimensions must be multiples of 8", who);
    UWU_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0,
                  "%s: q/k/v must be 16-byte aligned", who);
    return UWU_OK;
}

}  // namespace

int attn_any_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk, int d,
                 long long ldq, long long ldk, long long ldv, long long ldo, float scale, cudaStream_t stream) {
    AnyArgs a{};
    a.q = reinterpret_cast<const __nv_bfloat16*>(q);
    a.k = reinterpret_cast<const __nv_bfloat16*>(k);
    a.v = reinterpret_cast<const __nv_bfloat16*>(v);
    a.out = reinterpret_cast<__nv_bfloat16*>(o);
    a.lse = lse;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
    a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.Lq_pad = (Lq + 127) / 128 * 128; a.d = d;
    a.scale = scale; a.scale_log2 = scale * LOG2E;
    if (int rc = check_any(a, "uwu_attn_fwd")) return rc;
    if (d <= 80) return launch_fwd<80>(a, stream);
    if (d <= 128) return launch_fwd<128>(a, stream);
    return launch_fwd<160>(a, stream);
}

// Short-key path (Lk <= 128, head dim <= 64): 1 if taken, 0 if the shape does not qualify, < 0 / > 1 on error
// Forward: on by default (UWU_ATTN_SHORT=0 disables).  Backward: the two-kernel short-key backward measures slower than the
// tcgen05 kernel at the step shapes (221 vs 150 us at B16 h20 L1024 Lk77: its dK/dV pass is block-barrier bound), so it is
// opt-in (UWU_ATTN_SHORT_BWD=1) and the default backward stays on tcgen05 (same lse convention, so the two mix freely).
bool attn_short_ok(int Lq, int Lk, int d, bool backward) {
    const char* e = getenv(backward ? "UWU_ATTN_SHORT_BWD" : "UWU_ATTN_SHORT");
    const int enabled = e ? atoi(e) : (backward ? 0 : 1);
    if (!enabled || Lk > 128 || d > 64 || d % 8 != 0) return false;
    // backward scratch (2 * LKP * 64 floats per (b, h)) must fit the dQ region of the workspace (Lq_pad * 64 floats)
    const int lkp = Lk <= 80 ? 80 : 128;
    return (Lq + 127) / 128 * 128 >= 2 * lkp;
}

int attn_short_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk, int d,
                   long long ldq, long long ldk, long long ldv, long long ldo, float scale, cudaStream_t stream) {
    AnyArgs a{};
    a.q = reinterpret_cast<const __nv_bfloat16*>(q);
    a.k = reinterpret_cast<const __nv_bfloat16*>(k);
    a.v = reinterpret_cast<const __nv_bfloat16*>(v);
    a.out = reinterpret_cast<__nv_bfloat16*>(o);
    a.lse = lse;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
    a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.Lq_pad = (Lq + 127) / 128 * 128; a.d = d;
    a.scale = scale; a.scale_log2 = scale * LOG2E;
    if (int rc = check_any(a, "uwu_attn_fwd")) return rc;
    return Lk <= 80 ? launch_short_fwd<10>(a, stream) : launch_short_fwd<16>(a, stream);
}

int attn_short_fwd_masked(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk, int d,
                          long long ldq, long long ldk, long long ldv, long long ldo, float scale, int causal, const int32_t* key_mask,
                          cudaStream_t stream) {
    AnyArgs a{};
    a.q = reinterpret_cast<const __nv_bfloat16*>(q);
    a.k = reinterpret_cast<const __nv_bfloat16*>(k);
    a.v = reinterpret_cast<const __nv_bfloat16*>(v);
    a.out = reinterpret_cast<__nv_bfloat16*>(o);
    a.lse = lse;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
    a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.Lq_pad = (Lq + 127) / 128 * 128; a.d = d;
    a.scale = scale; a.scale_log2 = scale * LOG2E;
    a.causal = causal; a.key_mask = key_mask;
    if (int rc = check_any(a, "uwu_attn_fwd_masked")) return rc;
    return Lk <= 80 ? launch_short_fwd<10, true>(a, stream) : launch_short_fwd<16, true>(a, stream);
}

int attn_short_bwd(const void* q, const void* k, const void* v, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                   int B, int heads, int Lq, int Lk, int d, long long ldq, long long ldk, long long ldv, long long lddo,
                   long long lddq, long long lddk, long long lddv, float scale, float* workspace, cudaStream_t stream) {
    AnyArgs a{};
    a.q = reinterpret_cast<const __nv_bfloat16*>(q);
    a.k = reinterpret_cast<const __nv_bfloat16*>(k);
    a.v = reinterpret_cast<const __nv_bfloat16*>(v);
    a.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
    a.dq = reinterpret_cast<__nv_bfloat16*>(dq);
    a.dk = reinterpret_cast<__nv_bfloat16*>(dk);
    a.dv = reinterpret_cast<__nv_bfloat16*>(dv);
    a.lse = const_cast<float*>(lse);
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.lddo = lddo; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
    a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.Lq_pad = (Lq + 127) / 128 * 128; a.d = d;
    a.scale = scale; a.scale_log2 = scale * LOG2E;
    if (int rc = check_any(a, "uwu_attn_bwd")) return rc;
    return Lk <= 80 ? launch_short_bwd<10>(a, workspace, stream) : launch_short_bwd<16>(a, workspace, stream);
}

int attn_any_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq,
                 void* dk, void* dv, int B, int heads, int Lq, int Lk, int d, long long ldq, long long ldk, long long ldv,
                 long long ldo, long long lddo, long long lddq, long long lddk, long long lddv, float scale, float* workspace,
                 cudaStream_t stream) {
    AnyArgs a{};
    a.q = reinterpret_cast<const __nv_bfloat16*>(q);
    a.k = reinterpret_cast<const __nv_bfloat16*>(k);
    a.v = reinterpret_cast<const __nv_bfloat16*>(v);
    a.o = reinterpret_cast<const __nv_bfloat16*>(o);
    a.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
    a.dq = reinterpret_cast<__nv_bfloat16*>(dq);
    a.dk = reinterpret_cast<__nv_bfloat16*>(dk);
    a.dv = reinterpret_cast<__nv_bfloat16*>(dv);
    a.lse = const_cast<float*>(lse);
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.lddo = lddo; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
    a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.Lq_pad = (Lq + 127) / 128 * 128; a.d = d;
    a.scale = scale; a.scale_log2 = scale * LOG2E;
    if (int rc = check_any(a, "uwu_attn_bwd")) return rc;
    const long long rows = (long long)B * heads * a.Lq_pad;
    float* lse2 = workspace;
    float* delta = workspace + rows;
    a.lse2 = lse2;
    a.delta = delta;
    const long long total = (long long)B * Lq * heads;
    attn_any_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a, lse2, delta);
    UWU_CHECK_LAUNCH();
    if (d <= 80) return launch_bwd<80, 64>(a, stream);
    if (d <= 128) return launch_bwd<128, 32>(a, stream);
    return launch_bwd<160, 32>(a, stream);
}

}  // namespace uwu
