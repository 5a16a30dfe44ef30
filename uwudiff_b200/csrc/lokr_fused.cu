// LoKr adapter gradients of an attention projection in ONE pass over the activations (no G = dY^T X):
//
//   y = x (W + s * kron(w1, w2))^T,  w1 [ol, im], w2 [64, 64],  x [M, im*64], dY [M, ol*64]  (bf16, fp32 accumulation)
//   dw2[p, q] += s * sum_{t,i,j} dY[t, i*64+p] * w1[i, j] * X[t, j*64+q]
//   dw1[i, j] += s * sum_{t,p,q} dY[t, i*64+p] * w2[p, q] * X[t, j*64+q]
//
// Replaces autograd through lycoris' `make_kron(w1, w2)` weight rebuild in the forward patched in by
// src/duwu/trainer/trainer.py:152-154 (preset configs/lycoris/sdxl-diffusers.toml: Attention -> lokr, factor 64, full_matrix).
//
// Tokens are viewed as rows of 64 channels: X as [(t, j), q] (im rows per token), dY as [(t, i), p] (ol rows per token); a
// tile is T = 128 / max(ol, im) tokens, i.e. <= 128 such rows of 128 bytes, loaded by 3-D TMA boxes {64, im|ol, T} into
// 128-byte-swizzled shared memory.  Per tile four tcgen05 products, all with the tile rows as M, N or K:
//   (1) Z  = BD(w1) * Xs            [128 (t,i) x 64 q]   BD = block-diagonal copies of w1 (bf16, built once per CTA)
//   (2) V  = Xs * w2^T              [128 (t,j) x 64 p]
//   (3) D2 += dYs^T * bf16(Z)       [64 p x 64 q]        -> dw2   (A operand MN-major; M = 128: rows 64..127 are unused)
//   (4) Tm += dYs * bf16(V)^T       [128 (t,i) x 128 (t',j)] -> dw1[i, j] = sum_t Tm[(t,i), (t,j)]  (diagonal blocks)
// Z and V pass TMEM -> registers -> bf16 -> shared memory through four converter warps (double-buffered), D2 and Tm stay in
// TMEM for the whole kernel; warp 0 = TMA producer, warp 1 = MMA issuer.  HBM traffic = one read of X and dY.  Measured on B200
// (profiles/r02_lokr_fused.log): 30 us per 16384 x 1280 adapter against 52 us for the G = dY^T X GEMM + contraction; the ring
// depth (2..4 stages) does not change it and the loads alone stream at 6.6 TB/s (12.6 us) — the 24 small UMMAs per 6-token
// tile (N = 64 / 128, ~100 cycles each, operand reads from shared memory) bound the kernel at ~1.2 us per tile.
#include "api_internal.h"
#include "common.cuh"

namespace uwu {

namespace {

constexpr int LF_MAX_STAGES = 8;
constexpr int LF_THREADS = 192;
// shared memory (offsets from a 1024-aligned base): the stage ring first, so that the unused second M-atom of product (3)
// (16 KB after a stage's dYs tile) always lies inside the allocation
constexpr uint32_t LF_STAGE0 = 0;                     // n_stages dYs tiles, then n_stages Xs tiles (each rounded up to 8 rows of 128 B)
constexpr uint32_t LF_RING = 122880;                  // 4 x (15 + 15) KB (w1 20x20 / 10x10) or 3 x (16 + 16) KB
constexpr uint32_t LF_ZV = LF_RING;                   // 2 x (Zs 16 KB | Vs 16 KB); Zs[0] doubles as w1 / dw1 scratch
constexpr uint32_t LF_BD1 = LF_ZV + 2 * 32768;        // 2 K-atoms x 16 KB
constexpr uint32_t LF_W2 = LF_BD1 + 32768;            // 8 KB
constexpr uint32_t LF_SMEM = LF_W2 + 8192 + 1024;     // 230400 B (+ ~200 B static) of the 232448 an SM has
constexpr uint32_t LF_COL_D1 = 0, LF_COL_D3 = 128, LF_COL_D2 = 256, LF_COL_T = 320;  // TMEM columns (512 allocated)

struct LokrFusedArgs {
    CUtensorMap tmX, tmDY;
    const float* w1;
    const float* w2;
    float* dw1;
    float* dw2;
    int M, ol, im, T, n_tiles;
    int n_stages, y_bytes, x_bytes;  // ring: n_stages dYs tiles of y_bytes, then n_stages Xs tiles of x_bytes
    float scale;
};

__global__ void __launch_bounds__(LF_THREADS, 1) lokr_fused_kernel(const __grid_constant__ LokrFusedArgs p) {
    pdl_trigger();  // dependents may be scheduled as this grid's CTAs retire; they wait for its completion before touching memory
    extern __shared__ uint8_t lf_raw[];
    // two rings with their own barriers: an Xs tile is released as soon as products (1) and (2) have read it, one tile-period
    // before its dYs tile
    __shared__ __align__(8) uint64_t xfull[LF_MAX_STAGES], xempty[LF_MAX_STAGES], yfull[LF_MAX_STAGES], yempty[LF_MAX_STAGES];
    __shared__ __align__(8) uint64_t zv_full[2], zv_ready[2], acc_done[2], fin;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sm0 = (smem_u32(lf_raw) + 1023u) & ~1023u;
    uint8_t* sm = lf_raw + (sm0 - smem_u32(lf_raw));
    float* const w1_s = reinterpret_cast<float*>(sm + LF_ZV + 2 * 32768 - 4096);  // prologue only: tail of Vs[1]
    float* const dw1_s = reinterpret_cast<float*>(sm + LF_ZV);  // epilogue only (after the last product)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int rows_y = p.T * p.ol, rows_x = p.T * p.im;
    const int n_stages = p.n_stages;
    const uint32_t y_bytes = (uint32_t)p.y_bytes, x_bytes = (uint32_t)p.x_bytes;
    const uint32_t x_ring = LF_STAGE0 + (uint32_t)n_stages * y_bytes;

    // ---- nothing in this part depends on the previous kernel (it overlaps its tail under PDL) ----
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(&xfull[i], 1);
            mbar_init(&xempty[i], 1);
            mbar_init(&yfull[i], 1);
            mbar_init(&yempty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&zv_full[i], 1);
            mbar_init(&zv_ready[i], 4);
            mbar_init(&acc_done[i], 1);
        }
        mbar_init(&fin, 1);
        fence_barrier_init();
        tma_prefetch_desc(&p.tmX);
        tma_prefetch_desc(&p.tmDY);
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    {
        // the products run over 128 rows / k: rows past the T tokens of a tile must hold finite values (they meet zero rows
        // or columns of the other operand) -> the whole ring and the block-diagonal operand start as zeros
        const uint4 z = make_uint4(0, 0, 0, 0);
        // (+ 1 KB: the rows past the last Xs tile of the ring are the first rows of Zs[0])
        for (int id = threadIdx.x; id < (int)((LF_RING + 1024) >> 4); id += LF_THREADS) st_shared_v4(sm0 + LF_STAGE0 + ((uint32_t)id << 4), z);
        for (int id = threadIdx.x; id < 2 * 128 * 8; id += LF_THREADS) st_shared_v4(sm0 + LF_BD1 + ((uint32_t)id << 4), z);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    pdl_wait();
    if (warp != 0) {
        // ---- constant operands (built while warp 0 already streams tiles): w2 (K-major), block-diagonal copies of w1 ----
        const int tid = threadIdx.x - 32;
        for (int id = tid; id < p.ol * p.im; id += LF_THREADS - 32) w1_s[id] = p.w1[id];
        for (int id = tid; id < 64 * 8; id += LF_THREADS - 32) {
            const int r = id >> 3, cc = id & 7;
            const float4 a = *reinterpret_cast<const float4*>(p.w2 + r * 64 + cc * 8), b = *reinterpret_cast<const float4*>(p.w2 + r * 64 + cc * 8 + 4);
            uint4 u;
            u.x = pack_bf16(a.x, a.y); u.y = pack_bf16(a.z, a.w); u.z = pack_bf16(b.x, b.y); u.w = pack_bf16(b.z, b.w);
            st_shared_v4(sm0 + LF_W2 + r * 128 + ((cc ^ (r & 7)) << 4), u);
        }
        named_bar_sync(2, LF_THREADS - 32);
        // element (r = t*ol + i, c = t*im + j) = w1[i, j]
        for (int id = tid; id < rows_y * p.im; id += LF_THREADS - 32) {
            const int r = id / p.im, j = id - r * p.im;
            const int t = r / p.ol, i = r - t * p.ol;
            const int c = t * p.im + j;
            const uint32_t off = (uint32_t)((c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
            const __nv_bfloat16 v = __float2bfloat16(w1_s[i * p.im + j]);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(sm0 + LF_BD1 + off), "h"(*reinterpret_cast<const unsigned short*>(&v)) : "memory");
        }
        fence_proxy_async();
        named_bar_sync(2, LF_THREADS - 32);
    }

    if (warp == 0) {
        // ===================== TMA producer =====================
        for (int n = 0; n < n_my; ++n) {
            const int stage = n % n_stages;
            const uint32_t par = (uint32_t)((n / n_stages) - 1) & 1u;
            const int t0 = ((int)blockIdx.x + n * (int)gridDim.x) * p.T;
            if (n >= n_stages) mbar_wait(&xempty[stage], par);
            if (elect_one()) {
                mbar_expect_tx(&xfull[stage], (uint32_t)rows_x * 128u);
                tma_load_3d(sm + x_ring + stage * x_bytes, &p.tmX, &xfull[stage], 0, 0, t0);
            }
            __syncwarp();
            if (n >= n_stages) mbar_wait(&yempty[stage], par);
            if (elect_one()) {
                mbar_expect_tx(&yfull[stage], (uint32_t)rows_y * 128u);
                tma_load_3d(sm + LF_STAGE0 + stage * y_bytes, &p.tmDY, &yfull[stage], 0, 0, t0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_z = make_idesc_bf16(128, 64, 0, 1);   // BD1 (K-major) x Xs (MN-major)
        const uint32_t idesc_v = make_idesc_bf16(128, 64, 0, 0);   // Xs (K-major) x w2 (K-major)
        const uint32_t idesc_w2 = make_idesc_bf16(128, 64, 1, 1);  // dYs^T (MN-major) x Zs (MN-major)
        const uint32_t idesc_t = make_idesc_bf16(128, 128, 0, 0);  // dYs (K-major) x Vs (K-major)
        const uint64_t d_bd1 = make_smem_desc(sm0 + LF_BD1, 0, 1024);
        const uint64_t d_w2 = make_smem_desc(sm0 + LF_W2, 0, 1024);
        auto issue_zv = [&](int n) {
            const int stage = n % n_stages, b = n & 1;
            mbar_wait(&xfull[stage], (uint32_t)(n / n_stages) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t xs = sm0 + x_ring + stage * x_bytes;
                const uint64_t d_x_mn = make_smem_desc(xs, 8192, 1024), d_x_k = make_smem_desc(xs, 0, 1024);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tmem + LF_COL_D1 + b * 64, d_bd1 + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4), d_x_mn + (uint64_t)((k * 2048) >> 4),
                              idesc_z, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem + LF_COL_D3 + b * 64, d_x_k + (uint64_t)((k * 32) >> 4), d_w2 + (uint64_t)((k * 32) >> 4), idesc_v, k > 0 ? 1u : 0u);
                umma_commit(&xempty[stage]);
                umma_commit(&zv_full[b]);
            }
            __syncwarp();
        };
        if (n_my > 0) issue_zv(0);
        for (int n = 0; n < n_my; ++n) {
            if (n + 1 < n_my) issue_zv(n + 1);
            const int stage = n % n_stages, b = n & 1;
            mbar_wait(&yfull[stage], (uint32_t)(n / n_stages) & 1u);
            mbar_wait(&zv_ready[b], (uint32_t)(n >> 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ys = sm0 + LF_STAGE0 + stage * y_bytes;
                const uint32_t zs = sm0 + LF_ZV + b * 32768, vs = zs + 16384;
                const uint64_t d_y_mn = make_smem_desc(ys, 16384, 1024), d_y_k = make_smem_desc(ys, 0, 1024);
                const uint64_t d_z = make_smem_desc(zs, 8192, 1024), d_v = make_smem_desc(vs, 0, 1024);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tmem + LF_COL_D2, d_y_mn + (uint64_t)((k * 2048) >> 4), d_z + (uint64_t)((k * 2048) >> 4), idesc_w2,
                              (n > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem + LF_COL_T, d_y_k + (uint64_t)((k * 32) >> 4), d_v + (uint64_t)((k * 32) >> 4), idesc_t, (n > 0 || k > 0) ? 1u : 0u);
                umma_commit(&yempty[stage]);
                umma_commit(&acc_done[b]);
                if (n == n_my - 1) umma_commit(&fin);
            }
            __syncwarp();
        }
    } else {
        // ===================== converters (TMEM -> bf16 -> shared) and final reduction =====================
        const int q = warp & 3;  // TMEM sub-partition of this warp
        const int r = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        for (int n = 0; n < n_my; ++n) {
            const int b = n & 1;
            mbar_wait(&zv_full[b], (uint32_t)(n >> 1) & 1u);
            if (n >= 2) mbar_wait(&acc_done[b], (uint32_t)((n >> 1) - 1) & 1u);
            tc_fence_after();
            const uint32_t zs = sm0 + LF_ZV + b * 32768 + r * 128;
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const uint32_t col = (which == 0 ? LF_COL_D1 : LF_COL_D3) + b * 64;
                const uint32_t dst = zs + which * 16384;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_off + col + h * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 u;
                        u.x = pack_bf16(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1]));
                        u.y = pack_bf16(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3]));
                        u.z = pack_bf16(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5]));
                        u.w = pack_bf16(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7]));
                        st_shared_v4(dst + (((h * 4 + c) ^ (r & 7)) << 4), u);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&zv_ready[b]);
        }
        if (n_my > 0) {
            mbar_wait(&fin, 0);
            tc_fence_after();
            // dw1: row r = (t, i) of Tm holds its diagonal block in columns [t*im, t*im + im): every warp reads those columns
            // for each token its 32 rows touch (tcgen05.ld takes a warp-uniform column), the owning lanes park them in shared
            // memory (Zs is free now) as S[r][j]; the sum over tokens follows after the barrier
            {
                const int t = r / p.ol;
                const int t_lo = (q * 32) / p.ol, t_hi = min(p.T - 1, (q * 32 + 31) / p.ol);
                for (int tt = t_lo; tt <= t_hi; ++tt) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_off + LF_COL_T + tt * p.im, v);
                    tmem_ld_wait();
                    if (tt == t && r < rows_y) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < p.im) dw1_s[r * p.im + j] = __uint_as_float(v[j]);
                    }
                }
            }
            // dw2: rows 0..63 of D2
            if (r < 64) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_off + LF_COL_D2 + h * 32, v);
                    tmem_ld_wait();
                    float* dst = p.dw2 + r * 64 + h * 32;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 4), "f"(__uint_as_float(v[c * 4]) * p.scale),
                                     "f"(__uint_as_float(v[c * 4 + 1]) * p.scale), "f"(__uint_as_float(v[c * 4 + 2]) * p.scale),
                                     "f"(__uint_as_float(v[c * 4 + 3]) * p.scale)
                                     : "memory");
                }
            }
            tc_fence_before();
            named_bar_sync(1, 128);
            for (int id = threadIdx.x - 64; id < p.ol * p.im; id += 128) {
                const int i = id / p.im, j = id - i * p.im;
                float a = 0.f;
                for (int tt = 0; tt < p.T; ++tt) a += dw1_s[(tt * p.ol + i) * p.im + j];
                atomicAdd(p.dw1 + id, a * p.scale);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace

}  // namespace uwu

using namespace uwu;

extern "C" int uwu_lokr_fused_supported(int32_t out_l, int32_t out_k, int32_t in_m, int32_t in_n) {
    return out_k == 64 && in_n == 64 && out_l >= 1 && in_m >= 1 && out_l <= 32 && in_m <= 32;
}

extern "C" int uwu_lokr_fused_grad(const void* x, int64_t ldx, const void* dy, int64_t ldy, int64_t M, int32_t out_l, int32_t in_m,
                                   const float* w1, const float* w2, float* dw1, float* dw2, float scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(x && dy && w1 && w2 && dw1 && dw2 && M > 0, "uwu_lokr_fused_grad: bad arguments");
    UWU_CHECK_ARG(uwu_lokr_fused_supported(out_l, 64, in_m, 64), "uwu_lokr_fused_grad: w1 %dx%d unsupported (w2 must be 64x64, w1 <= 32x32)",
                  out_l, in_m);
    UWU_CHECK_ARG(ldx >= (int64_t)in_m * 64 && ldy >= (int64_t)out_l * 64 && ldx % 8 == 0 && ldy % 8 == 0 && M < (1ll << 31),
                  "uwu_lokr_fused_grad: bad leading dimensions");
    UWU_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)w2 & 15) == 0 && ((uintptr_t)dw2 & 15) == 0 &&
                      ((uintptr_t)dw1 & 15) == 0,
                  "uwu_lokr_fused_grad: pointers must be 16-byte aligned");
    LokrFusedArgs p;
    p.w1 = w1; p.w2 = w2; p.dw1 = dw1; p.dw2 = dw2;
    p.M = (int)M; p.ol = out_l; p.im = in_m; p.scale = scale;
    p.T = 128 / (out_l > in_m ? out_l : in_m);
    p.n_tiles = (int)((M + p.T - 1) / p.T);
    p.y_bytes = ((p.T * out_l + 7) / 8) * 1024;
    p.x_bytes = ((p.T * in_m + 7) / 8) * 1024;
    p.n_stages = (int)(LF_RING / (uint32_t)(p.y_bytes + p.x_bytes));
    if (p.n_stages > LF_MAX_STAGES) p.n_stages = LF_MAX_STAGES;
    {
        uint64_t dims[3] = {64, (uint64_t)in_m, (uint64_t)M};
        uint64_t str[2] = {128, (uint64_t)ldx * 2};
        uint32_t box[3] = {64, (uint32_t)in_m, (uint32_t)p.T};
        if (encode_tmap_bf16(&p.tmX, x, 3, dims, str, box, 1)) return UWU_ERR_INVALID;
    }
    {
        uint64_t dims[3] = {64, (uint64_t)out_l, (uint64_t)M};
        uint64_t str[2] = {128, (uint64_t)ldy * 2};
        uint32_t box[3] = {64, (uint32_t)out_l, (uint32_t)p.T};
        if (encode_tmap_bf16(&p.tmDY, dy, 3, dims, str, box, 1)) return UWU_ERR_INVALID;
    }
    static bool attr_done = false;
    if (!attr_done) {
        UWU_CHECK_CUDA(cudaFuncSetAttribute(lokr_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LF_SMEM));
        attr_done = true;
    }
    int grid = sm_count();
    if (grid > p.n_tiles) grid = p.n_tiles;
    UWU_CHECK_CUDA(launch_pdl(lokr_fused_kernel, dim3(grid), dim3(LF_THREADS), LF_SMEM, stream, p));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
