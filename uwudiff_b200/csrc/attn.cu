// Flash-style attention forward and backward on tcgen05 / TMEM / TMA for sm_100a (head_dim 64, bf16, fp32 softmax).
//
//   forward : O = softmax(scale * Q K^T) V, LSE      one CTA = 2 x 128 query rows of one (batch, head), ping-ponged
//   backward: dK, dV                                  pass 1: one CTA = 128 keys of one (batch, head), loops over query tiles
//             dQ                                      pass 2: one CTA = 128 queries, loops over key tiles (S / dP recomputed)
//
// Q/K/V/O live in token-major activations [B*L, ld] with head h in columns [64h, 64h+64) (exactly how the fused
// QKV projection GEMM writes them), so no permutes are needed: 3-D tensor maps {64h.., l, b} pick the head slice.
//
// forward roles : warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax of query tile 0, warps 6-9 of tile 1.
//   S = Q K^T lands in TMEM (128 lanes x 128 fp32 columns per tile); each softmax thread owns one query row
//   (tcgen05.ld 32x32b), writes P as bf16 into 128B-swizzled shared memory, the P V product goes to a second TMEM
//   region and is folded into fp32 register accumulators with the running-max rescale.
// backward roles: warp 0 TMA, warp 1 MMA, warps 2-9 compute.  Pass 1, per query tile: S^T = K Q^T and dP^T = V dO^T in
//   TMEM (thread = key row), P^T = exp2(S^T*c - lse), dS^T = P^T o (dP^T - delta) written to swizzled smem as bf16, then
//   dV += P^T dO and dK += dS^T Q (TMEM accumulators over the whole loop).  Pass 2 mirrors it with the roles of queries
//   and keys swapped (thread = query row) and accumulates dQ += dS K.  No cross-CTA reduction, no fp32 scratch.
//
// Replaces F.scaled_dot_product_attention under diffusers' AttnProcessor2_0 (the in-tree copy of that flow is
// src/duwu/modules/rope_unet.py:76-175; SDPA call at :151) for attn1 (self) and attn2 (cross, Lk = 77).
#include <cstdlib>
#include "api_internal.h"
#include "common.cuh"

namespace uwu {

static constexpr int AT_TILE = 128 * 64 * 2;  // one [128 x 64] bf16 operand tile, 16 KiB
static constexpr float LOG2E = 1.4426950408889634f;

UWU_DEVINL uint8_t* align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// K-major SW128 descriptors (row = M/N index, 64 bf16 = 128 B per row): LBO unused, SBO = 8 rows * 128 B
UWU_DEVINL uint64_t desc_kmajor(uint32_t saddr) { return make_smem_desc(saddr, 0, 1024); }
// MN-major SW128 descriptors (row = K index, 64 contiguous M/N elements per row)
UWU_DEVINL uint64_t desc_mnmajor(uint32_t saddr, uint32_t lbo) { return make_smem_desc(saddr, lbo, 1024); }

// ================================================================================================
// forward
// ================================================================================================
struct AttnFwdArgs {
    CUtensorMap tmQ, tmK, tmV;  // 4-D (head_dim, heads, L, B); boxes are 64 wide, columns >= head_dim read as zero
    int Lq, Lk, heads, Lq_pad;
    int hd;  // head dim (multiple of 8, <= 64): narrower heads run zero-padded to the 64-wide tiles
    float scale, scale_log2;
    __nv_bfloat16* o;
    long long ldo;
    float* lse;  // [B, heads, Lq_pad]
};

static constexpr int FWD_ST = 4;                            // K / V stages (TMA runs up to 4 key tiles ahead)
static constexpr int FWD_SQ = 0;                            // 2 query tiles
static constexpr int FWD_SK = 2 * AT_TILE;                  // FWD_ST stages
static constexpr int FWD_SV = (2 + FWD_ST) * AT_TILE;       // FWD_ST stages
static constexpr int FWD_SP = (2 + 2 * FWD_ST) * AT_TILE;   // 2 tiles x [128 x 128] bf16
static constexpr int FWD_BAR = (6 + 2 * FWD_ST) * AT_TILE;
static constexpr int FWD_SMEM = FWD_BAR + 256 + 1024;
static constexpr int FWD_THREADS = 320;

__global__ void __launch_bounds__(FWD_THREADS, 1) attn_fwd_kernel(const __grid_constant__ AttnFwdArgs p) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FWD_BAR);
    uint64_t* q_full = bars;                      // [1]
    uint64_t* k_full = bars + 1;                  // [FWD_ST]
    uint64_t* k_empty = k_full + FWD_ST;          // [FWD_ST]
    uint64_t* v_full = k_empty + FWD_ST;          // [FWD_ST]
    uint64_t* v_empty = v_full + FWD_ST;          // [FWD_ST]
    uint64_t* s_full = v_empty + FWD_ST;          // [2] per tile
    uint64_t* p_full = s_full + 2;                // [2]
    uint64_t* pv_full = p_full + 2;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
    const int ntiles = (q0 + 128 < p.Lq) ? 2 : 1;
    const int nkv = (p.Lk + 127) / 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmQ);
        tma_prefetch_desc(&p.tmK);
        tma_prefetch_desc(&p.tmV);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < FWD_ST; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], 4);
            mbar_init(&pv_full[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    // TMEM columns: S tile t at [128 t, 128 t + 128), PV tile t at [256 + 64 t, +64)

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, (uint32_t)(ntiles * AT_TILE));
            tma_load_4d(smem + FWD_SQ, &p.tmQ, q_full, 0, h, q0, b);
            if (ntiles == 2) tma_load_4d(smem + FWD_SQ + AT_TILE, &p.tmQ, q_full, 0, h, q0 + 128, b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j % FWD_ST;
                const uint32_t ph = (uint32_t)((j / FWD_ST) & 1);
                mbar_wait(&k_empty[s], ph ^ 1);
                mbar_expect_tx(&k_full[s], AT_TILE);
                tma_load_4d(smem + FWD_SK + s * AT_TILE, &p.tmK, &k_full[s], 0, h, j * 128, b);
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_expect_tx(&v_full[s], AT_TILE);
                tma_load_4d(smem + FWD_SV + s * AT_TILE, &p.tmV, &v_full[s], 0, h, j * 128, b);
            }
        }
    } else if (warp == 1) {
        // the whole warp walks the loop (warp-uniform waits and operands); one elected lane issues each group of UMMAs
        {
            const uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
            const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
            const uint32_t sq = smem_u32(smem + FWD_SQ), sk = smem_u32(smem + FWD_SK), sv = smem_u32(smem + FWD_SV),
                           sp = smem_u32(smem + FWD_SP);
            auto issue_qk = [&](int t, int s) {
                const uint64_t ad = desc_kmajor(sq + t * AT_TILE), bd = desc_kmajor(sk + s * AT_TILE);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + (uint32_t)(t * 128), ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_qk,
                              (uint32_t)(k != 0));
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            if (elect_one()) {
                for (int t = 0; t < ntiles; ++t) {
                    issue_qk(t, 0);
                    umma_commit(&s_full[t]);
                }
                umma_commit(&k_empty[0]);
            }
            __syncwarp();
            for (int j = 0; j < nkv; ++j) {
                const int s = j % FWD_ST;
                const uint32_t ph = (uint32_t)((j / FWD_ST) & 1);
                const bool has_next = j + 1 < nkv;
                const int s1 = (j + 1) % FWD_ST;
                mbar_wait(&v_full[s], ph);
                if (has_next) mbar_wait(&k_full[s1], (uint32_t)(((j + 1) / FWD_ST) & 1));
                for (int t = 0; t < ntiles; ++t) {
                    mbar_wait(&p_full[t], (uint32_t)(j & 1));
                    tc_fence_after();
                    if (elect_one()) {
                        // S(t, j) has been consumed: the next scores go first, they are what the softmax warps wait for;
                        // nobody waits for P V (O lives in TMEM) except the P-buffer reuse guard below
                        if (has_next) {
                            issue_qk(t, s1);
                            umma_commit(&s_full[t]);
                        }
                        const uint64_t vd = desc_mnmajor(sv + s * AT_TILE, 8192);
                        const uint64_t pd = desc_kmajor(sp + t * 2 * AT_TILE);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint64_t ad = pd + (uint64_t)(((k >> 2) * AT_TILE) >> 4) + (uint64_t)((k & 3) * 2);
                            umma_bf16(tmem_base + 256u + (uint32_t)(t * 64), ad, vd + (uint64_t)(k * 128), idesc_pv,
                                      (uint32_t)((j | k) != 0));  // O accumulates in TMEM over the whole key loop
                        }
                        umma_commit(&pv_full[t]);
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    umma_commit(&v_empty[s]);
                    if (has_next) umma_commit(&k_empty[s1]);
                }
                __syncwarp();
            }
        }
    } else {
        const int t = (warp - 2) >> 2;
        if (t < ntiles) {
            const int sub = warp & 3;
            const int row = sub * 32 + lane;  // row within the 128-query tile == TMEM lane
            const uint32_t lane_addr = (uint32_t)(sub * 32) << 16;
            const uint32_t t_s = tmem_base + lane_addr + (uint32_t)(t * 128);
            const uint32_t t_pv = tmem_base + lane_addr + 256u + (uint32_t)(t * 64);
            const uint32_t prow_s = smem_u32(smem + FWD_SP + t * 2 * AT_TILE + row * 128);
            const int sw = row & 7;
            const float sl2 = p.scale_log2;
            // O lives in TMEM (the P V products accumulate there); m is the running maximum the probabilities are expressed
            // against.  It is only advanced -- and O / l rescaled -- when the true row maximum has grown by more than 2^8
            // (lazy rescaling): p <= 256 is harmless in bf16 / fp32 and the final O / l is exact for any reference maximum.
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < nkv; ++j) {
                mbar_wait(&s_full[t], (uint32_t)(j & 1));
                tc_fence_after();
                const int kv_valid = min(128, p.Lk - j * 128);
                const bool full = kv_valid == 128;
                float rs = 0.f;
                if (full) {
                    // ---- full key tile: S is read from TMEM ONCE (128 registers per thread).  TMEM reads are the scarcest
                    // resource of this kernel: 128 lanes x 128 fp32 columns = 64 KiB per tile, and the tensor-memory load path
                    // delivers on the order of 64 B / clock / SM, so every extra pass over S costs ~1000 cycles per tile ----
                    uint32_t sr[4][32];
                    tmem_ld32(t_s, sr[0]);
                    tmem_ld32(t_s + 32u, sr[1]);
                    tmem_ld32(t_s + 64u, sr[2]);
                    tmem_ld32(t_s + 96u, sr[3]);
                    tmem_ld_wait();
                    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        m0 = fmax3(m0, __uint_as_float(sr[0][i]), __uint_as_float(sr[0][i + 1]));
                        m1 = fmax3(m1, __uint_as_float(sr[1][i]), __uint_as_float(sr[1][i + 1]));
                        m2 = fmax3(m2, __uint_as_float(sr[2][i]), __uint_as_float(sr[2][i + 1]));
                        m3 = fmax3(m3, __uint_as_float(sr[3][i]), __uint_as_float(sr[3][i + 1]));
                    }
                    const float m_new = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                    if (j == 0) {
                        m = m_new;
                    } else if (__any_sync(0xffffffffu, (m_new - m) * sl2 > 8.0f)) {
                        // rare: bring O (complete up to the previous key tile) and l to the new reference maximum
                        mbar_wait(&pv_full[t], (uint32_t)((j - 1) & 1));
                        tc_fence_after();
                        const float alpha = ex2_approx((m - m_new) * sl2);
                        // 8 columns at a time: the 128 score registers stay live across this (rare) path, a 32-register
                        // temporary here would push them onto the stack in the common path
#pragma unroll 1
                        for (int c = 0; c < 8; ++c) {
                            uint32_t r[8];
                            tmem_ld8(t_pv + (uint32_t)(c * 8), r);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                            tmem_st8(t_pv + (uint32_t)(c * 8), r);
                        }
                        tmem_st_wait();
                        l *= alpha;
                        m = m_new;
                    }
                    const float msl = m * sl2;
                    // the P V product of the previous key tile must have finished reading the P buffer before it is overwritten
                    if (j > 0) mbar_wait(&pv_full[t], (uint32_t)((j - 1) & 1));
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // independent partial row sums (no serial FADD chain)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            float e[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(sr[c][q4 * 8 + i]), sl2, -msl));
                            s0 += e[0] + e[4];
                            s1 += e[1] + e[5];
                            s2 += e[2] + e[6];
                            s3 += e[3] + e[7];
                            uint4 u;
                            u.x = pack_bf16(e[0], e[1]);
                            u.y = pack_bf16(e[2], e[3]);
                            u.z = pack_bf16(e[4], e[5]);
                            u.w = pack_bf16(e[6], e[7]);
                            const int cc = (c & 1) * 4 + q4;
                            st_shared_v4(prow_s + (uint32_t)((c >> 1) * AT_TILE + ((cc ^ sw) << 4)), u);
                        }
                    }
                    rs = (s0 + s1) + (s2 + s3);
                } else {
                    // ---- ragged last key tile (Lk % 128 != 0): two masked passes over TMEM ----
                    float mx = -INFINITY;
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp) {
                        uint32_t ra[32], rb[32];
                        tmem_ld32(t_s + (uint32_t)(cp * 64), ra);
                        tmem_ld32(t_s + (uint32_t)(cp * 64 + 32), rb);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (cp * 64 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(ra[i]));
                            if (cp * 64 + 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(rb[i]));
                        }
                    }
                    const float m_new = fmaxf(m, mx);
                    if (j == 0) {
                        m = m_new;
                    } else if (__any_sync(0xffffffffu, (m_new - m) * sl2 > 8.0f)) {
                        mbar_wait(&pv_full[t], (uint32_t)((j - 1) & 1));
                        tc_fence_after();
                        const float alpha = ex2_approx((m - m_new) * sl2);
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t r[32];
                            tmem_ld32(t_pv + (uint32_t)(c * 32), r);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                            tmem_st32(t_pv + (uint32_t)(c * 32), r);
                        }
                        tmem_st_wait();
                        l *= alpha;
                        m = m_new;
                    }
                    const float msl = m * sl2;
                    if (j > 0) mbar_wait(&pv_full[t], (uint32_t)((j - 1) & 1));
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t r[32];
                        tmem_ld32(t_s + (uint32_t)(c * 32), r);
                        tmem_ld_wait();
                        float pe[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float e = ex2_approx(fmaf(__uint_as_float(r[i]), sl2, -msl));
                            pe[i] = (c * 32 + i < kv_valid) ? e : 0.f;
                            rs += pe[i];
                        }
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            uint4 u;
                            u.x = pack_bf16(pe[q4 * 8 + 0], pe[q4 * 8 + 1]);
                            u.y = pack_bf16(pe[q4 * 8 + 2], pe[q4 * 8 + 3]);
                            u.z = pack_bf16(pe[q4 * 8 + 4], pe[q4 * 8 + 5]);
                            u.w = pack_bf16(pe[q4 * 8 + 6], pe[q4 * 8 + 7]);
                            const int cc = (c & 1) * 4 + q4;
                            st_shared_v4(prow_s + (uint32_t)((c >> 1) * AT_TILE + ((cc ^ sw) << 4)), u);
                        }
                    }
                }
                l += rs;
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t]);
            }
            // all products have landed: read O once
            mbar_wait(&pv_full[t], (uint32_t)((nkv - 1) & 1));
            tc_fence_after();
            float O[64];
            {
                uint32_t r0[32], r1[32];
                tmem_ld32(t_pv, r0);
                tmem_ld32(t_pv + 32u, r1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    O[i] = __uint_as_float(r0[i]);
                    O[32 + i] = __uint_as_float(r1[i]);
                }
            }
            const int q = q0 + t * 128 + row;
            if (q < p.Lq) {
                const float inv = 1.0f / l;
                __nv_bfloat16* op = p.o + ((size_t)b * p.Lq + q) * p.ldo + h * p.hd;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 8 >= p.hd) break;
                    uint4 u;
                    u.x = pack_bf16(O[c * 8 + 0] * inv, O[c * 8 + 1] * inv);
                    u.y = pack_bf16(O[c * 8 + 2] * inv, O[c * 8 + 3] * inv);
                    u.z = pack_bf16(O[c * 8 + 4] * inv, O[c * 8 + 5] * inv);
                    u.w = pack_bf16(O[c * 8 + 6] * inv, O[c * 8 + 7] * inv);
                    *reinterpret_cast<uint4*>(op + c * 8) = u;
                }
                p.lse[((size_t)b * p.heads + h) * p.Lq_pad + q] = fmaf(m, p.scale, __logf(l));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ================================================================================================
// backward
// ================================================================================================
// delta[b,h,q] = sum_d O[q,d] dO[q,d];  lse2 = lse * log2(e)
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ dout,
                                     long long lddo, const float* __restrict__ lse, int B, int heads, int Lq, int Lq_pad, int hd,
                                     float* __restrict__ lse2, float* __restrict__ delta) {
    pdl_trigger();
    // 8 lanes per (row, head): lane j reads the j-th 16-byte chunk of the head's O and dO rows, so a warp instruction covers
    // four contiguous 128-byte head slices (one thread per slice made every load touch 32 half-used sectors)
    const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int c = threadIdx.x & 7;
    const long long total = (long long)B * Lq * heads;
    const bool ok = item < total;
    const long long it = ok ? item : total - 1;
    const int h = (int)(it % heads);
    const long long bq = it / heads;
    const int q = (int)(bq % Lq);
    const int b = (int)(bq / Lq);
    float acc = 0.f;
    if (ok && c * 8 < hd) {
        const uint4 a = *reinterpret_cast<const uint4*>(o + bq * ldo + h * hd + c * 8);
        const uint4 d = *reinterpret_cast<const uint4*>(dout + bq * lddo + h * hd + c * 8);
        const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
        const float2 d0 = unpack_bf16(d.x), d1 = unpack_bf16(d.y), d2 = unpack_bf16(d.z), d3 = unpack_bf16(d.w);
        acc = a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y + a2.x * d2.x + a2.y * d2.y + a3.x * d3.x + a3.y * d3.y;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (ok && c == 0) {
        const size_t oi = ((size_t)b * heads + h) * Lq_pad + q;
        delta[oi] = acc;
        lse2[oi] = lse[oi] * LOG2E;
    }
}

// dq[b*Lq+q, 64h+d] = scale * acc[b,h,q/128, d/4, q%128, d%4]   (single-pass backward: fp32 dQ scratch -> bf16)
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, int B, int heads, int Lq, int nqt, float scale, int hd,
                                           __nv_bfloat16* __restrict__ dq, long long lddq) {
    pdl_trigger();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 8-wide d chunk per thread
    const long long total = (long long)B * Lq * heads * 8;
    if (idx >= total) return;
    const int c8 = (int)(idx & 7);
    if (c8 * 8 >= hd) return;
    long long r = idx >> 3;
    const int h = (int)(r % heads);
    r /= heads;
    const int q = (int)(r % Lq);
    const int b = (int)(r / Lq);
    const size_t tile = (((size_t)b * heads + h) * nqt + (q >> 7)) * (16 * 128 * 4);
    const float4 f0 = *reinterpret_cast<const float4*>(acc + tile + ((size_t)(2 * c8) * 128 + (q & 127)) * 4);
    const float4 f1 = *reinterpret_cast<const float4*>(acc + tile + ((size_t)(2 * c8 + 1) * 128 + (q & 127)) * 4);
    uint4 u;
    u.x = pack_bf16(f0.x * scale, f0.y * scale);
    u.y = pack_bf16(f0.z * scale, f0.w * scale);
    u.z = pack_bf16(f1.x * scale, f1.y * scale);
    u.w = pack_bf16(f1.z * scale, f1.w * scale);
    *reinterpret_cast<uint4*>(dq + ((size_t)b * Lq + q) * lddq + h * hd + c8 * 8) = u;
}

struct AttnBwdArgs {
    CUtensorMap tmQ, tmK, tmV, tmdO;
    float* dq_acc;  // single-pass mode: fp32 [B, heads, Lq_pad/128, 16, 128, 4]
    int Lq, Lk, heads, Lq_pad;
    float scale, scale_log2;
    const float* lse2;   // [B, heads, Lq_pad]
    const float* delta;  // [B, heads, Lq_pad]
    __nv_bfloat16* dq;
    long long lddq;
    __nv_bfloat16* dk;
    long long lddk;
    __nv_bfloat16* dv;
    long long lddv;
    int hd;  // head dim (multiple of 8, <= 64)
    int gx, n_items;  // work items: gx resident tiles per (batch, head), gx * heads * B in total (persistent CTAs)
    int direct_dq;  // single-pass mode with ONE key tile (cross-attention): each dQ tile is complete in its CTA -> bf16 store,
                    // no fp32 scratch, no memset, no conversion pass
};

// Shared-memory plan: TWO buffers of two resident operand tiles (the next work item's K / V arrive while this one computes),
// dS and P tiles, kStages x two streamed operand tiles.
static constexpr int BWD_SR = 0;             // resident: buffer kvb, operand w at (2 kvb + w) tiles: K|V (dK/dV pass) or Q|dO (dQ pass)
static constexpr int BWD_SDS = 4 * AT_TILE;  // dS [128 queries x 128 keys] bf16, two [128 x 64] K-major tiles
static constexpr int BWD_SP = 6 * AT_TILE;   // P, same layout (dK/dV pass only)
// kMode 0: dK/dV pass, 1: dQ pass, 2: single pass (dK, dV and dQ partials reduced into an fp32 scratch with vector atomics)
template <int kMode>
struct BwdCfg {
    static constexpr bool kDQ = kMode == 1;
    static constexpr int kStages = kDQ ? 4 : 3;
    static constexpr int SS0 = (kDQ ? 6 : 8) * AT_TILE;  // streamed 0: Q (dK/dV pass) or K (dQ pass)
    static constexpr int SS1 = SS0 + kStages * AT_TILE;   // streamed 1: dO             or V
    static constexpr int BAR = SS1 + kStages * AT_TILE;
    static constexpr int SMEM = BAR + 256 + 1024;
};
static constexpr int BWD_THREADS = 64 + 512;  // TMA warp, MMA warp, 16 compute warps (4 per SM sub-partition)
// TMEM columns: S [0,128) dP [128,256) acc0 [256,320) acc1 [320,384) dQ partial [384,448)

// PERSISTENT CTAs over the work items (resident tile x, head, batch): one CTA per SM walks items blockIdx.x, + gridDim.x, ...
// All pipelines (operand ring, S/dP, P/dS, dQ) run through the item boundaries: the next item's resident tiles are loaded into
// the other K|V buffer, its first S / dP products are issued before the last accumulating products of the current item, and
// only the drain of the dK / dV accumulators (16 warps, ~1 us) separates two items.  With 8 query tiles per item (the
// 1024-token levels) set-up + first loads + pipeline fill + drain + relaunch were a third of a CTA's life.
//
// Two kinds of pass, neither needs a cross-CTA reduction for dK / dV:
//   kDQ = false: item = 128 keys of one (batch, head), streams the query tiles (3 stages): dV += P^T dO, dK += dS^T Q
//                (kMode 2 also forms dQ_i = dS K per query tile and reduces it into an fp32 scratch with vector atomics)
//   kDQ = true : item = 128 queries, streams the key tiles (4 stages):                     dQ += dS K
// S = Q K^T and dP = dO V^T land in TMEM with one thread per QUERY row, so lse / delta are per-thread scalars;
// P = exp2(S c - lse) and dS = P o (dP - delta) go to 128B-swizzled smem as [query][key] bf16, which the accumulating
// products read K-major (dQ) or MN-major (dV, dK: the transposes, for free).  The MMA warp issues S / dP of step g+1 before
// the accumulating products of step g, so the exp / dS math overlaps the tensor pipe.
template <int kMode>
__global__ void __launch_bounds__(BWD_THREADS, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdArgs p) {
    pdl_trigger();
    using Cfg = BwdCfg<kMode>;
    constexpr bool kDQ = kMode == 1;
    constexpr bool kFused = kMode == 2;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR);
    uint64_t* res_full = bars;        // [2] one per resident buffer
    uint64_t* res_empty = bars + 2;   // [2]
    uint64_t* st_full = bars + 4;     // [4]
    uint64_t* st_empty = bars + 8;    // [4]
    uint64_t* sdp_full = bars + 12;
    uint64_t* sdp_empty = bars + 13;
    uint64_t* pds_full = bars + 14;
    uint64_t* pds_empty = bars + 15;
    uint64_t* acc_full = bars + 16;
    uint64_t* acc_empty = bars + 17;
    uint64_t* dq_full = bars + 18;
    uint64_t* dq_empty = bars + 19;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iter = kDQ ? (p.Lk + 127) / 128 : (p.Lq + 127) / 128;
    // items of this CTA: first, stride, count; item -> (resident tile, head, batch)
    const int item0 = (int)blockIdx.x, istep = (int)gridDim.x;
    const int n_items = item0 < p.n_items ? (p.n_items - item0 + istep - 1) / istep : 0;
    const int g_total = n_items * n_iter;  // pipeline steps of this CTA
    auto decode = [&](int it, int& r0, int& h, int& b) {
        const int item = item0 + it * istep;
        const int x = item % p.gx;
        const int hb = item / p.gx;
        r0 = x * 128;
        h = hb % p.heads;
        b = hb / p.heads;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmQ);
        tma_prefetch_desc(&p.tmK);
        tma_prefetch_desc(&p.tmV);
        tma_prefetch_desc(&p.tmdO);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&res_empty[s], 1);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&st_full[s], 1);
            mbar_init(&st_empty[s], 1);
        }
        mbar_init(sdp_full, 1);
        mbar_init(sdp_empty, 16);
        mbar_init(pds_full, 16);
        mbar_init(pds_empty, 1);
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 16);
        mbar_init(dq_full, 1);
        mbar_init(dq_empty, 16);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            const CUtensorMap* tmR0 = kDQ ? &p.tmQ : &p.tmK;
            const CUtensorMap* tmR1 = kDQ ? &p.tmdO : &p.tmV;
            const CUtensorMap* tmS0 = kDQ ? &p.tmK : &p.tmQ;
            const CUtensorMap* tmS1 = kDQ ? &p.tmV : &p.tmdO;
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < n_items; ++it) {
                int r0, h, b;
                decode(it, r0, h, b);
                const int kvb = it & 1;
                // the products of item it - 2 that read this resident buffer have retired
                if (it >= 2) mbar_wait(&res_empty[kvb], (uint32_t)(((it >> 1) - 1) & 1));
                mbar_expect_tx(&res_full[kvb], 2 * AT_TILE);
                tma_load_4d(smem + BWD_SR + (2 * kvb) * AT_TILE, tmR0, &res_full[kvb], 0, h, r0, b);
                tma_load_4d(smem + BWD_SR + (2 * kvb + 1) * AT_TILE, tmR1, &res_full[kvb], 0, h, r0, b);
                for (int i = 0; i < n_iter; ++i) {
                    mbar_wait(&st_empty[s], ph ^ 1);
                    mbar_expect_tx(&st_full[s], 2 * AT_TILE);
                    tma_load_4d(smem + Cfg::SS0 + s * AT_TILE, tmS0, &st_full[s], 0, h, i * 128, b);
                    tma_load_4d(smem + Cfg::SS1 + s * AT_TILE, tmS1, &st_full[s], 0, h, i * 128, b);
                    if (++s == kStages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // the whole warp walks the loop (warp-uniform waits and operands); one elected lane issues each group of UMMAs
        if (g_total > 0) {
            const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);              // S, dP: both operands K-major
            const uint32_t idesc_acc = make_idesc_bf16(128, 64, kDQ ? 0 : 1, 1);  // dQ: A K-major; dV/dK: A MN-major; B MN-major
            const uint32_t sr = smem_u32(smem + BWD_SR), ss0 = smem_u32(smem + Cfg::SS0), ss1 = smem_u32(smem + Cfg::SS1),
                           sp = smem_u32(smem + BWD_SP), sds = smem_u32(smem + BWD_SDS);
            // S = Q K^T, dP = dO V^T with the query tile as the M operand in both passes
            auto issue_sdp = [&](int s, int kvb) {
                const uint32_t sr0 = sr + (2 * kvb) * AT_TILE, sr1 = sr0 + AT_TILE;
                const uint64_t q_d = desc_kmajor(kDQ ? sr0 : ss0 + s * AT_TILE), k_d = desc_kmajor(kDQ ? ss0 + s * AT_TILE : sr0);
                const uint64_t do_d = desc_kmajor(kDQ ? sr1 : ss1 + s * AT_TILE), v_d = desc_kmajor(kDQ ? ss1 + s * AT_TILE : sr1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + 0u, q_d + (uint64_t)(k * 2), k_d + (uint64_t)(k * 2), idesc_s, (uint32_t)(k != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + 128u, do_d + (uint64_t)(k * 2), v_d + (uint64_t)(k * 2), idesc_s, (uint32_t)(k != 0));
            };
            mbar_wait(&res_full[0], 0);
            mbar_wait(&st_full[0], 0);
            tc_fence_after();
            if (elect_one()) {
                issue_sdp(0, 0);
                umma_commit(sdp_full);
            }
            __syncwarp();
            int s = 0, s1 = (kStages > 1) ? 1 : 0;
            uint32_t ph1 = (kStages > 1) ? 0u : 1u;  // phase of stage s1's NEXT fill
            int i = 0, it = 0;
            for (int g = 0; g < g_total; ++g) {
                const int kvb = it & 1;
                const bool last = (i == n_iter - 1);
                if (g + 1 < g_total) {
                    mbar_wait(&st_full[s1], ph1);
                    const int it1 = last ? it + 1 : it;
                    if (last) mbar_wait(&res_full[it1 & 1], (uint32_t)((it1 >> 1) & 1));  // the next item's resident tiles
                    mbar_wait(sdp_empty, (uint32_t)(g & 1));  // compute warps have pulled S / dP of step g out of TMEM
                    tc_fence_after();
                    if (elect_one()) {
                        issue_sdp(s1, it1 & 1);
                        umma_commit(sdp_full);
                    }
                    __syncwarp();
                }
                mbar_wait(pds_full, (uint32_t)(g & 1));
                if (kFused) mbar_wait(dq_empty, (uint32_t)((g & 1) ^ 1));
                if (i == 0 && it > 0) mbar_wait(acc_empty, (uint32_t)((it - 1) & 1));  // previous item's accumulators drained
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sr0 = sr + (2 * kvb) * AT_TILE;
                    const uint64_t b0 = desc_mnmajor(ss0 + s * AT_TILE, 8192);  // Q (dK) or K (dQ), [rows = reduction, 64 d]
                    const uint64_t b1 = desc_mnmajor(ss1 + s * AT_TILE, 8192);  // dO (dV)
                    if (kDQ) {
                        // dQ[q, d] += sum_key dS[q, key] K[key, d]: A = dS K-major
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint64_t ad = desc_kmajor(sds + (k >> 2) * AT_TILE) + (uint64_t)((k & 3) * 2);
                            umma_bf16(tmem_base + 320u, ad, b0 + (uint64_t)(k * 128), idesc_acc, (uint32_t)((i | k) != 0));
                        }
                    } else {
                        // dV[key, d] += sum_q P[q, key] dO[q, d], dK[key, d] += sum_q dS[q, key] Q[q, d]:
                        // A = P / dS read MN-major (M = key: two 64-key blocks 16 KiB apart, 16 query rows = 2048 B per K step)
                        const uint64_t ap = desc_mnmajor(sp, 16384), ads = desc_mnmajor(sds, 16384);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16(tmem_base + 256u, ap + (uint64_t)(k * 128), b1 + (uint64_t)(k * 128), idesc_acc,
                                      (uint32_t)((i | k) != 0));
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16(tmem_base + 320u, ads + (uint64_t)(k * 128), b0 + (uint64_t)(k * 128), idesc_acc,
                                      (uint32_t)((i | k) != 0));
                        if (kFused) {
                            // dQ_i[q, d] = sum_key dS[q, key] K[key, d]: A = dS K-major, B = the resident K tile MN-major;
                            // a fresh 64-column accumulator per query tile, drained by the compute warps one step later
                            // (dq_empty was awaited by the whole warp above)
                            const uint32_t idesc_dq = make_idesc_bf16(128, 64, 0, 1);
                            const uint64_t kb = desc_mnmajor(sr0, 8192);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const uint64_t ad = desc_kmajor(sds + (k >> 2) * AT_TILE) + (uint64_t)((k & 3) * 2);
                                umma_bf16(tmem_base + 384u, ad, kb + (uint64_t)(k * 128), idesc_dq, (uint32_t)(k != 0));
                            }
                        }
                    }
                    umma_commit(&st_empty[s]);
                    umma_commit(pds_empty);
                    if (kFused) umma_commit(dq_full);
                    if (last) {
                        umma_commit(acc_full);
                        umma_commit(&res_empty[kvb]);
                    }
                }
                __syncwarp();
                s = s1;
                if (++s1 == kStages) {
                    s1 = 0;
                    ph1 ^= 1;
                }
                if (last) {
                    i = 0;
                    ++it;
                } else {
                    ++i;
                }
            }
        }
    } else {
        // 16 compute warps: warp (sub, quarter) owns TMEM lanes [32 sub, +32) (the query rows) and key columns [32 quarter, +32)
        const int cw = warp - 2;
        const int sub = warp & 3;
        const int quarter = cw >> 2;
        const int half = quarter >> 1;    // which [128 x 64] smem tile / which 32-column half of the 64-wide accumulators
        const int row = sub * 32 + lane;  // query row within the tile == TMEM lane
        const int sw = row & 7;
        const uint32_t lane_addr = (uint32_t)(sub * 32) << 16;
        const float sl2 = p.scale_log2;
        const uint64_t sl2_2 = pk2(sl2, sl2);
        const uint32_t prow_s = smem_u32(smem + BWD_SP + half * AT_TILE + row * 128);
        const uint32_t dsrow_s = smem_u32(smem + BWD_SDS + half * AT_TILE + row * 128);
        int r0 = 0, h = 0, b = 0;
        size_t bh = 0;
        // single-pass mode: drain the dQ partial of query tile `tile` (pipeline step gq; 16 columns per warp) into the fp32
        // scratch with vector atomics straight from registers: lane-contiguous 16-byte chunks, 512 B per warp instruction
        auto flush_dq = [&](int tile, int gq, size_t bhq, int hq, int bq) __attribute__((always_inline)) {
            uint32_t rq[16];
            mbar_wait(dq_full, (uint32_t)(gq & 1));
            tc_fence_after();
            tmem_ld16(tmem_base + lane_addr + 384u + (uint32_t)(quarter * 16), rq);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_empty);
            if (p.direct_dq) {
                const int qrow = tile * 128 + row;
                if (qrow < p.Lq) {
                    __nv_bfloat16* op = p.dq + ((size_t)bq * p.Lq + qrow) * p.lddq + hq * p.hd + quarter * 16;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (quarter * 16 + c * 8 >= p.hd) break;
                        uint4 u;
                        u.x = pack_bf16(__uint_as_float(rq[c * 8 + 0]) * p.scale, __uint_as_float(rq[c * 8 + 1]) * p.scale);
                        u.y = pack_bf16(__uint_as_float(rq[c * 8 + 2]) * p.scale, __uint_as_float(rq[c * 8 + 3]) * p.scale);
                        u.z = pack_bf16(__uint_as_float(rq[c * 8 + 4]) * p.scale, __uint_as_float(rq[c * 8 + 5]) * p.scale);
                        u.w = pack_bf16(__uint_as_float(rq[c * 8 + 6]) * p.scale, __uint_as_float(rq[c * 8 + 7]) * p.scale);
                        *reinterpret_cast<uint4*>(op + c * 8) = u;
                    }
                }
                return;
            }
            float* dst = p.dq_acc + (bhq * (size_t)(p.Lq_pad / 128) + tile) * (size_t)(16 * 128 * 4) + ((size_t)(quarter * 4) * 128 + row) * 4;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (size_t)c * 512),
                             "f"(__uint_as_float(rq[c * 4])), "f"(__uint_as_float(rq[c * 4 + 1])),
                             "f"(__uint_as_float(rq[c * 4 + 2])), "f"(__uint_as_float(rq[c * 4 + 3]))
                             : "memory");
        };
        int g = 0;
        for (int it = 0; it < n_items; ++it) {
            decode(it, r0, h, b);
            bh = (size_t)b * p.heads + h;
            // lse * log2(e) and delta of this thread's query row (padded rows of the scratch are zero and never written out)
            const float* lse_p = p.lse2 + bh * p.Lq_pad + row;
            const float* del_p = p.delta + bh * p.Lq_pad + row;
            float my_lse = lse_p[kDQ ? r0 : 0], my_delta = del_p[kDQ ? r0 : 0];
            for (int i = 0; i < n_iter; ++i, ++g) {
                float nx_lse = my_lse, nx_delta = my_delta;
                if (!kDQ && i + 1 < n_iter) {  // next query tile's scalars, in flight during this tile's math
                    nx_lse = lse_p[(size_t)(i + 1) * 128];
                    nx_delta = del_p[(size_t)(i + 1) * 128];
                }
                mbar_wait(sdp_full, (uint32_t)(g & 1));
                tc_fence_after();
                // pull this thread's 32 S and 32 dP values out of TMEM and hand the accumulators back to the MMA warp
                uint32_t rs1[32], rp1[32];
                tmem_ld32(tmem_base + lane_addr + (uint32_t)(quarter * 32), rs1);
                tmem_ld32(tmem_base + lane_addr + 128u + (uint32_t)(quarter * 32), rp1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sdp_empty);
                uint4 pu[kDQ ? 1 : 4], du[4];
                const uint64_t nl2 = pk2(-my_lse, -my_lse), nd2 = pk2(-my_delta, -my_delta);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint32_t pw[4], dw[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {  // two elements per step: FFMA2 / FADD2 / FMUL2
                        const int i2 = q4 * 8 + 2 * e;
                        float a0, a1, d0, d1;
                        unpk2(ffma2(pk2u(rs1[i2], rs1[i2 + 1]), sl2_2, nl2), a0, a1);
                        const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
                        unpk2(fmul2(pk2(p0, p1), fadd2(pk2u(rp1[i2], rp1[i2 + 1]), nd2)), d0, d1);
                        pw[e] = pack_bf16(p0, p1);
                        dw[e] = pack_bf16(d0, d1);
                    }
                    du[q4] = make_uint4(dw[0], dw[1], dw[2], dw[3]);
                    if (!kDQ) pu[kDQ ? 0 : q4] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                }
                // the accumulating products of the previous step have retired: their smem operands may be overwritten
                mbar_wait(pds_empty, (uint32_t)((g & 1) ^ 1));
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint32_t chunk = (uint32_t)(((((quarter & 1) << 2) + q4) ^ sw) << 4);
                    if (!kDQ) st_shared_v4(prow_s + chunk, pu[kDQ ? 0 : q4]);
                    st_shared_v4(dsrow_s + chunk, du[q4]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(pds_full);
                if (kFused && i > 0) flush_dq(i - 1, g - 1, bh, h, b);  // previous tile's dQ, off the critical path
                my_lse = nx_lse;
                my_delta = nx_delta;
            }
            if (kFused) flush_dq(n_iter - 1, g - 1, bh, h, b);
            // ---- write the accumulators (TMEM lane = key row in the dK/dV pass, query row in the dQ pass) ----
            // warp (sub, quarter) writes columns [16 quarter, +16) of its 32 rows
            mbar_wait(acc_full, (uint32_t)(it & 1));
            tc_fence_after();
            const int r = r0 + row;
            if (kDQ) {
                uint32_t v[16];
                tmem_ld16(tmem_base + lane_addr + 320u + (uint32_t)(quarter * 16), v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
                if (r < p.Lq) {
                    __nv_bfloat16* op = p.dq + ((size_t)b * p.Lq + r) * p.lddq + h * p.hd + quarter * 16;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (quarter * 16 + c * 8 >= p.hd) break;
                        uint4 u;
                        u.x = pack_bf16(__uint_as_float(v[c * 8 + 0]) * p.scale, __uint_as_float(v[c * 8 + 1]) * p.scale);
                        u.y = pack_bf16(__uint_as_float(v[c * 8 + 2]) * p.scale, __uint_as_float(v[c * 8 + 3]) * p.scale);
                        u.z = pack_bf16(__uint_as_float(v[c * 8 + 4]) * p.scale, __uint_as_float(v[c * 8 + 5]) * p.scale);
                        u.w = pack_bf16(__uint_as_float(v[c * 8 + 6]) * p.scale, __uint_as_float(v[c * 8 + 7]) * p.scale);
                        *reinterpret_cast<uint4*>(op + c * 8) = u;
                    }
                }
            } else {
                uint32_t v[2][16];
                tmem_ld16(tmem_base + lane_addr + 256u + (uint32_t)(quarter * 16), v[0]);
                tmem_ld16(tmem_base + lane_addr + 320u + (uint32_t)(quarter * 16), v[1]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);  // both accumulators are in registers: the next item may overwrite them
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    if (r < p.Lk) {
                        const float sc = which ? p.scale : 1.0f;
                        __nv_bfloat16* base = which ? p.dk + ((size_t)b * p.Lk + r) * p.lddk : p.dv + ((size_t)b * p.Lk + r) * p.lddv;
                        __nv_bfloat16* op = base + h * p.hd + quarter * 16;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            if (quarter * 16 + c * 8 >= p.hd) break;
                            uint4 u;
                            u.x = pack_bf16(__uint_as_float(v[which][c * 8 + 0]) * sc, __uint_as_float(v[which][c * 8 + 1]) * sc);
                            u.y = pack_bf16(__uint_as_float(v[which][c * 8 + 2]) * sc, __uint_as_float(v[which][c * 8 + 3]) * sc);
                            u.z = pack_bf16(__uint_as_float(v[which][c * 8 + 4]) * sc, __uint_as_float(v[which][c * 8 + 5]) * sc);
                            u.w = pack_bf16(__uint_as_float(v[which][c * 8 + 6]) * sc, __uint_as_float(v[which][c * 8 + 7]) * sc);
                            *reinterpret_cast<uint4*>(op + c * 8) = u;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int bwd_grid(int n_items) {  // persistent CTAs: one per SM (UWU_ATTN_BWD_PERSIST=0: one CTA per item, for A/B runs)
    static int persist = -1;
    if (persist < 0) {
        const char* e = getenv("UWU_ATTN_BWD_PERSIST");
        persist = e ? atoi(e) : 1;
    }
    const int sms = sm_count();
    return (!persist || n_items < sms) ? n_items : sms;
}

// (head_dim, heads, L, B) view of a [B*L, ld] activation; the box is always 64 columns wide, so heads narrower than 64 are
// zero-filled by TMA up to the tile width (out-of-bounds fill) and the 128-byte swizzled tile layout does not change.
static int make_head_map(CUtensorMap* tm, const void* base, int heads, int hd, int L, int B, long long ld, const char* what) {
    if (ld % 8 != 0 || ld < (long long)heads * hd) {
        set_error("uwu_attn: leading dimension %lld of %s must be a multiple of 8 and >= heads*head_dim", ld, what);
        return UWU_ERR_INVALID;
    }
    uint64_t dims[4] = {(uint64_t)hd, (uint64_t)heads, (uint64_t)L, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)hd * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)L};
    uint32_t box[4] = {64, 1, 128, 1};
    return encode_tmap_bf16(tm, base, 4, dims, str, box, 1);
}

// mma.sync kernels for head dims other than 64 (attn_any.cu)
int attn_any_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk, int d,
                 long long ldq, long long ldk, long long ldv, long long ldo, float scale, cudaStream_t stream);
int attn_any_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq,
                 void* dk, void* dv, int B, int heads, int Lq, int Lk, int d, long long ldq, long long ldk, long long ldv,
                 long long ldo, long long lddo, long long lddq, long long lddk, long long lddv, float scale, float* workspace,
                 cudaStream_t stream);

// short-key (cross-attention) kernels, attn_any.cu
bool attn_short_ok(int Lq, int Lk, int d, bool backward);
int attn_short_fwd_masked(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk, int d,
                          long long ldq, long long ldk, long long ldv, long long ldo, float scale, int causal, const int32_t* key_mask,
                          cudaStream_t stream);
int attn_short_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk, int d,
                   long long ldq, long long ldk, long long ldv, long long ldo, float scale, cudaStream_t stream);
int attn_short_bwd(const void* q, const void* k, const void* v, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                   int B, int heads, int Lq, int Lk, int d, long long ldq, long long ldk, long long ldv, long long lddo,
                   long long lddq, long long lddk, long long lddv, float scale, float* workspace, cudaStream_t stream);

}  // namespace uwu

using namespace uwu;

static int attn_check(int32_t B, int32_t heads, int32_t Lq, int32_t Lk, int32_t head_dim) {
    if (head_dim <= 0 || head_dim % 8 != 0 || head_dim > 160) {
        set_error("uwu_attn: head_dim %d unsupported (64 on tcgen05; other multiples of 8 up to 160 on mma.sync)", head_dim);
        return UWU_ERR_UNSUPPORTED;
    }
    UWU_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk > 0, "uwu_attn: bad shape B=%d heads=%d Lq=%d Lk=%d", B, heads, Lq, Lk);
    UWU_CHECK_ARG(B <= 65535 && heads <= 65535, "uwu_attn: batch/heads exceed grid limits");
    return UWU_OK;
}

extern "C" int64_t uwu_attn_lse_floats(int32_t B, int32_t heads, int32_t Lq) {
    if (B <= 0 || heads <= 0 || Lq <= 0) return 0;
    return (int64_t)B * heads * ((Lq + 127) / 128 * 128);
}

extern "C" int uwu_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t heads,
                            int32_t Lq, int32_t Lk, int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                            float scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = attn_check(B, heads, Lq, Lk, head_dim)) return rc;
    UWU_CHECK_ARG(q && k && v && o && lse, "uwu_attn_fwd: null pointer");
    UWU_CHECK_ARG(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "uwu_attn_fwd: output must be 16-byte aligned");
    if (head_dim > 64) return attn_any_fwd(q, k, v, o, lse, B, heads, Lq, Lk, head_dim, ldq, ldk, ldv, ldo, scale, stream);
    if (attn_short_ok(Lq, Lk, head_dim, false))
        return attn_short_fwd(q, k, v, o, lse, B, heads, Lq, Lk, head_dim, ldq, ldk, ldv, ldo, scale, stream);
    static thread_local AttnFwdArgs a;
    if (int rc = make_head_map(&a.tmQ, q, heads, head_dim, Lq, B, ldq, "q")) return rc;
    if (int rc = make_head_map(&a.tmK, k, heads, head_dim, Lk, B, ldk, "k")) return rc;
    if (int rc = make_head_map(&a.tmV, v, heads, head_dim, Lk, B, ldv, "v")) return rc;
    a.Lq = Lq; a.Lk = Lk; a.heads = heads; a.Lq_pad = (Lq + 127) / 128 * 128; a.hd = head_dim;
    a.scale = scale; a.scale_log2 = scale * LOG2E;
    a.o = reinterpret_cast<__nv_bfloat16*>(o); a.ldo = ldo; a.lse = lse;
    static bool attr_set = false;
    if (!attr_set) {
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
        attr_set = true;
    }
    dim3 grid((Lq + 255) / 256, heads, B);
    UWU_CHECK_CUDA(launch_pdl(attn_fwd_kernel, grid, dim3(FWD_THREADS), FWD_SMEM, stream, a));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_attn_fwd_masked(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t heads,
                                   int32_t Lq, int32_t Lk, int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                   float scale, int32_t causal, const int32_t* key_mask, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = attn_check(B, heads, Lq, Lk, head_dim)) return rc;
    UWU_CHECK_ARG(q && k && v && o && lse, "uwu_attn_fwd_masked: null pointer");
    UWU_CHECK_ARG(Lk <= 128 && head_dim <= 64 && head_dim % 8 == 0,
                  "uwu_attn_fwd_masked: built for text-encoder sequences (Lk <= 128, head_dim <= 64), got Lk=%d d=%d", Lk, head_dim);
    UWU_CHECK_ARG(!causal || Lq == Lk, "uwu_attn_fwd_masked: the causal mask needs Lq == Lk");
    UWU_CHECK_ARG(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "uwu_attn_fwd_masked: output must be 16-byte aligned");
    return attn_short_fwd_masked(q, k, v, o, lse, B, heads, Lq, Lk, head_dim, ldq, ldk, ldv, ldo, scale, causal, key_mask, stream);
}

extern "C" int64_t uwu_attn_bwd_workspace_floats(int32_t B, int32_t heads, int32_t Lq) {
    if (B <= 0 || heads <= 0 || Lq <= 0) return 0;
    const int64_t rows = (int64_t)B * heads * ((Lq + 127) / 128 * 128);
    return rows * 2 + rows * 64;  // lse * log2(e), delta, fp32 dQ scratch (single-pass mode)
}

// 0 = two passes (dK/dV, then dQ; no atomics), 1 = single pass with an fp32 dQ scratch.  Default from UWU_ATTN_BWD_MODE.
static int attn_bwd_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("UWU_ATTN_BWD_MODE");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

extern "C" int uwu_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                            void* dq, void* dk, void* dv, int32_t B, int32_t heads, int32_t Lq, int32_t Lk,
                            int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo,
                            int64_t lddq, int64_t lddk, int64_t lddv, float scale, float* workspace, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = attn_check(B, heads, Lq, Lk, head_dim)) return rc;
    UWU_CHECK_ARG(q && k && v && o && dout && lse && dq && dk && dv && workspace, "uwu_attn_bwd: null pointer");
    UWU_CHECK_ARG(ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
                  "uwu_attn_bwd: leading dimensions must be multiples of 8");
    UWU_CHECK_ARG(((reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dq) |
                    reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv) |
                    reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
                  "uwu_attn_bwd: pointers must be 16-byte aligned");
    if (head_dim > 64)
        return attn_any_bwd(q, k, v, o, dout, lse, dq, dk, dv, B, heads, Lq, Lk, head_dim, ldq, ldk, ldv, ldo, lddo, lddq, lddk,
                            lddv, scale, workspace, stream);
    if (attn_short_ok(Lq, Lk, head_dim, true))
        return attn_short_bwd(q, k, v, dout, lse, dq, dk, dv, B, heads, Lq, Lk, head_dim, ldq, ldk, ldv, lddo, lddq, lddk, lddv,
                              scale, workspace, stream);
    const int Lq_pad = (Lq + 127) / 128 * 128;
    const int64_t rows = (int64_t)B * heads * Lq_pad;
    float* lse2 = workspace;
    float* delta = workspace + rows;
    static thread_local AttnBwdArgs a;
    if (int rc = make_head_map(&a.tmQ, q, heads, head_dim, Lq, B, ldq, "q")) return rc;
    if (int rc = make_head_map(&a.tmK, k, heads, head_dim, Lk, B, ldk, "k")) return rc;
    if (int rc = make_head_map(&a.tmV, v, heads, head_dim, Lk, B, ldv, "v")) return rc;
    if (int rc = make_head_map(&a.tmdO, dout, heads, head_dim, Lq, B, lddo, "dout")) return rc;
    a.Lq = Lq; a.Lk = Lk; a.heads = heads; a.Lq_pad = Lq_pad; a.hd = head_dim;
    a.scale = scale; a.scale_log2 = scale * LOG2E;
    a.lse2 = lse2; a.delta = delta;
    a.dq = reinterpret_cast<__nv_bfloat16*>(dq); a.lddq = lddq;
    a.dk = reinterpret_cast<__nv_bfloat16*>(dk); a.lddk = lddk;
    a.dv = reinterpret_cast<__nv_bfloat16*>(dv); a.lddv = lddv;
    if (Lq_pad != Lq)  // padded query rows of the scratch are read (and masked) by the kernels: keep them finite
        UWU_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)(rows * 2) * sizeof(float), stream));
    {
        const long long total = (long long)B * Lq * heads;
        attn_bwd_prep_kernel<<<(unsigned)((total * 8 + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(o), ldo, reinterpret_cast<const __nv_bfloat16*>(dout), lddo, lse, B, heads,
            Lq, Lq_pad, head_dim, lse2, delta);
        UWU_CHECK_LAUNCH();
    }
    static bool attr_set = false;
    if (!attr_set) {
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg<0>::SMEM));
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg<1>::SMEM));
        UWU_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg<2>::SMEM));
        attr_set = true;
    }
    a.direct_dq = 0;
    if (attn_bwd_mode() == 1 && Lk <= 128) {
        a.dq_acc = workspace + 2 * rows;
        a.direct_dq = 1;
        a.gx = 1;
        a.n_items = heads * B;
        UWU_CHECK_CUDA(launch_pdl(attn_bwd_kernel<2>, dim3(bwd_grid(a.n_items)), dim3(BWD_THREADS), BwdCfg<2>::SMEM, stream, a));
        UWU_CHECK_LAUNCH();
    } else if (attn_bwd_mode() == 1) {
        a.dq_acc = workspace + 2 * rows;
        UWU_CHECK_CUDA(cudaMemsetAsync(a.dq_acc, 0, (size_t)(rows * 64) * sizeof(float), stream));
        a.gx = (Lk + 127) / 128;
        a.n_items = a.gx * heads * B;
        UWU_CHECK_CUDA(launch_pdl(attn_bwd_kernel<2>, dim3(bwd_grid(a.n_items)), dim3(BWD_THREADS), BwdCfg<2>::SMEM, stream,
                                  a));  // dK, dV, dQ partials
        UWU_CHECK_LAUNCH();
        const long long total = (long long)B * Lq * heads * 8;
        attn_bwd_dq_convert_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            a.dq_acc, B, heads, Lq, Lq_pad / 128, scale, head_dim, reinterpret_cast<__nv_bfloat16*>(dq), lddq);
        UWU_CHECK_LAUNCH();
    } else {
        a.gx = (Lk + 127) / 128;
        a.n_items = a.gx * heads * B;
        attn_bwd_kernel<0><<<dim3(bwd_grid(a.n_items)), BWD_THREADS, BwdCfg<0>::SMEM, stream>>>(a);  // dK, dV
        UWU_CHECK_LAUNCH();
        a.gx = (Lq + 127) / 128;
        a.n_items = a.gx * heads * B;
        attn_bwd_kernel<1><<<dim3(bwd_grid(a.n_items)), BWD_THREADS, BwdCfg<1>::SMEM, stream>>>(a);  // dQ
        UWU_CHECK_LAUNCH();
    }
    return UWU_OK;
}
