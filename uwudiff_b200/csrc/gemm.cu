// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = alpha * sum_k A[m,k] * B[n,k]  (+ bias[n]) (+ bias_rows[m / rows_per_bias, n]) (+ residual[m,n])
//
// * operands bf16, accumulation fp32 in TMEM (two accumulator buffers of <=256 columns),
// * tiles 128 x block_n x 64, operands staged by TMA into 128-byte-swizzled shared memory,
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2..5 = epilogue
//   (tcgen05.ld -> registers -> fused epilogue -> global),
// * A may be: row-major [M,K] (K contiguous), "column-major" [K,M] (M contiguous, used for the
//   token-reduction weight-gradient GEMMs), or an NHWC activation read through a 4-D tensor map with a
//   table of (dn,dh,dw) taps (3x3 conv, stride-2 conv through phase planes, transposed-conv phases);
//   out-of-image taps are zero-filled by TMA.  A second A source covers channel-concatenated inputs.
// * B may be [N,K] (K contiguous: y = x W^T) or [K,N] (N contiguous).
//
// Replaces the cuBLASLt / cuDNN calls under diffusers' Linear / Conv2d.
#include <cstdlib>
#include "api_internal.h"
#include "common.cuh"

namespace uwu {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;
static constexpr int UMMA_K = 16;
static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
static constexpr int NUM_THREADS = 192;                      // 6 warps
static constexpr int MAX_STAGES = 8;
static constexpr int TMEM_COLS = 512;
static constexpr int ACC_STRIDE = 256;
static constexpr int EPI_STAGE_BYTES = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x [32 rows x 128 B]

struct GemmKernelArgs {
    CUtensorMap tmA;
    CUtensorMap tmA2;
    CUtensorMap tmB;
    CUtensorMap tmO;   // bf16 output, box {64 columns, 32 rows}, 128B swizzle (epilogue TMA stores)
    CUtensorMap tmO2;  // same for out2 (columns >= n_split)
    int epi_tma;       // 1: bf16 output through shared-memory staging + TMA stores (coalesced), else direct stores
    int res_tma;       // 1: the residual tile is prefetched by TMA (tmAux), one tile ahead, into per-chunk slots of each warp
    int M, N;
    int num_kb;  // number of 64-wide K blocks (conv: ntaps * kb_per_tap)
    int block_n;
    int num_m_tiles, num_n_tiles;
    int a_mode;  // 0 K-major 2D, 1 MN-major 2D, 2 conv 4D
    int b_mode;  // 0 K-major [N,K], 1 MN-major [K,N]
    int stages;
    int b_stage_bytes;
    int tx_bytes;
    uint32_t a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv;
    // conv
    int H, W, kb_per_tap, kb_src1;  // kb_src1: 64-blocks per tap that come from source 1
    int tap_dn[9], tap_dh[9], tap_dw[9];
    // epilogue
    void* out;
    void* out2;
    long long ldo, ldo2;
    int n_split;
    int out_fp32;
    const float* bias;
    const float* bias_rows;
    int rows_per_bias;
    const __nv_bfloat16* residual;
    long long ldr;
    float alpha;
    int accumulate;
    // segmented reduction (a_mode 1 / b_mode 1): k-block kb belongs to segment kb / kb_per_seg, whose operands start
    // seg * a_seg_off / seg * b_seg_off elements further along the contiguous (M / N) axis of A / B
    int kb_per_seg, a_seg_off, b_seg_off;
    // grouped N (a_mode 0): output columns [g * grp_n, (g+1) * grp_n) use A columns starting at g * a_grp_koff and the
    // SAME B ([K, grp_n]) for every group
    int grp_n, a_grp_koff;
    CUtensorMap tmBh;  // cluster mode: half-height box of B (each CTA of the pair loads one half and multicasts it)
    int cluster2;      // 1: launched as 2-CTA clusters sharing every B tile (same n-tile, adjacent m-tiles)
    int b_half_bytes;  // shared-memory offset of the second half of a B stage
    int b_3d;  // MN-major B through a 3-D tensor map {64 n, K, N/64}: ONE TMA instruction per stage instead of block_n/64
    // fused GEGLU epilogues (pair kernel): 1 forward (N = 2F: TMEM columns [0,128) = h, [128,256) = g of 128 hidden units),
    // 2 backward (N = F: the tile's dg is combined with h, g read from the saved pre-activation `aux`)
    int epi_mode, geglu_f, epi_warp_bytes;
    CUtensorMap tmAux;  // mode 2: pre-activation [M, 2F], box {64 columns, 32 rows}, 128B swizzle
    // wide tiles (pair kernel, K-major B): block_n = n_inst * bn_inst (<= 512) is computed by n_inst UMMAs per k-step that share
    // the A tile (fewer operand bytes per FLOP from L2, the limiter of this kernel); block_n > 256 leaves room for ONE accumulator
    int n_inst, bn_inst, b_inst_bytes, n_acc;
    int pair;      // 1: CTA-pair kernel (cta_group::2): one 256 x block_n tile per cluster of two CTAs; each CTA stages its
                   //    own 128 rows of A and HALF of the B tile, the leader CTA issues M = 256 UMMAs for both
    int stream_k;  // 1: the (tile, k-block) space is cut into equal contiguous ranges, one per CTA; partial tiles are
                   //    reduced with vector atomics into the fp32 output (which the host zeroed unless accumulating)
                   // 2: the reduction is cut into sk_slices slices; units (slice, tile) are dealt round-robin in slice-major
                   //    order, so the CTAs running at any moment share ONE k-slice of A and B through L2 (mode 1 spreads
                   //    them over the whole reduction and re-reads every operand slab from HBM once per tile)
    int sk_slices;
};

// One unit of work for the three warp roles: k-blocks [kb0, kb1) of output tile `tile`.
struct WorkIter {
    long long cur, end;
    int step, num_kb, stream_k;
    int tile, kb0, kb1;
    int pair = 0, rank = 0, nnt = 1;
    int slices = 1, ntiles = 1;
    __device__ __forceinline__ void init(const GemmKernelArgs& p, int num_tiles) {
        num_kb = p.num_kb;
        stream_k = p.stream_k;
        // workers: CTAs, or clusters of two when the unit of work is a 256-row tile pair (both CTAs walk the same units)
        int bid = blockIdx.x, nb = gridDim.x;
        if (p.cluster2 || p.pair) {
            pair = 1;
            rank = (int)(blockIdx.x & 1);
            nnt = p.num_n_tiles;
            bid >>= 1;
            nb >>= 1;
            num_tiles = ((p.num_m_tiles + 1) >> 1) * p.num_n_tiles;
        }
        if (stream_k == 2) {
            slices = p.sk_slices;
            ntiles = num_tiles;
            cur = bid;
            end = (long long)num_tiles * slices;
            step = nb;
        } else if (stream_k) {
            const long long units = (long long)num_tiles * num_kb;
            cur = units * bid / nb;
            end = units * (bid + 1) / nb;
            step = 0;
        } else {
            cur = bid;
            end = num_tiles;
            step = nb;
        }
    }
    __device__ __forceinline__ bool next() {
        if (cur >= end) return false;
        if (stream_k == 2) {
            const int sl = (int)(cur / ntiles);
            tile = (int)(cur - (long long)sl * ntiles);
            kb0 = (int)((long long)num_kb * sl / slices);
            kb1 = (int)((long long)num_kb * (sl + 1) / slices);
            cur += step;
        } else if (stream_k) {
            tile = (int)(cur / num_kb);
            kb0 = (int)(cur - (long long)tile * num_kb);
            const long long left = end - cur;
            kb1 = (left < (long long)(num_kb - kb0)) ? kb0 + (int)left : num_kb;
            cur += kb1 - kb0;
        } else {
            tile = (int)cur;
            kb0 = 0;
            kb1 = num_kb;
            cur += step;
        }
        if (pair) {  // unit -> this CTA's tile id (m-tile 2*m_pair + rank may lie past the last row tile: fully clipped)
            const int n_blk = tile % nnt, m_pair = tile / nnt;
            tile = (2 * m_pair + rank) * nnt + n_blk;
        }
        return true;
    }
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmKernelArgs p) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = A_STAGE_BYTES + p.b_stage_bytes;
    uint8_t* epi_stage = smem + (size_t)p.stages * stage_bytes;  // 1024-byte aligned (stage sizes are multiples of 1024)
    uint8_t* bar_base = epi_stage + 4 * p.epi_warp_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tmem_full = empty_bar + MAX_STAGES;  // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* epi_bar = tmem_empty + 2;            // [4] one per epilogue warp (TMA loads of epilogue operands)
    uint64_t* res_bar = epi_bar + 4;               // [4 warps][4 chunks] residual slots (res_tma)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 16);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmA2);
        tma_prefetch_desc(&p.tmB);
        if (p.epi_tma) {
            tma_prefetch_desc(&p.tmO);
            tma_prefetch_desc(&p.tmO2);
        }
        if (p.epi_mode == 2 || p.res_tma) tma_prefetch_desc(&p.tmAux);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], (!PAIR && p.cluster2) ? 2 : 1);  // multicast-cluster mode: both CTAs' MMA warps release a stage
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], PAIR ? 8 : 4);  // one arrive per epilogue warp (pair: of both CTAs, on the leader's barrier)
        }
        for (int s = 0; s < 4; ++s) mbar_init(&epi_bar[s], 1);
        for (int s = 0; s < 16; ++s) mbar_init(&res_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) {
            tmem_alloc2(tmem_slot, TMEM_COLS);
            tmem_relinquish2();
        } else {
            tmem_alloc(tmem_slot, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR || p.cluster2) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlaps the tail of the previous kernel

    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        // the whole warp walks the loop (warp-uniform operands); one elected lane issues the copies of a k-block
        {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = (uint32_t)p.tx_bytes;
            WorkIter wi;
            wi.init(p, num_tiles);
            while (wi.next()) {
                const int tile = wi.tile;
                const int n_blk = tile % p.num_n_tiles;
                const int m_blk = tile / p.num_n_tiles;
                const int m0 = m_blk * BLOCK_M;
                const int n0 = n_blk * p.block_n;
                int cw = 0, ch = 0, cn = 0;
                if (p.a_mode == 2) {
                    cw = m0 % p.W;
                    ch = (m0 / p.W) % p.H;
                    cn = m0 / (p.W * p.H);
                }
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    if (elect_one()) {
                        if (PAIR) {
                            // Both CTAs' boxes complete on the LEADER's full barrier (the leader expects the bytes of both).
                            // The peer may run ahead of the leader's expect_tx: its stage was released by the leader's MMA
                            // commit, i.e. the barrier is already in the phase these bytes belong to, and a phase cannot complete
                            // before the leader's own arrive.
                            const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], tx_bytes);
                            const int kseg = p.kb_per_seg > 0 ? kb / p.kb_per_seg : 0;
                            const int kbs = kb - kseg * p.kb_per_seg;
                            const int grp = p.grp_n > 0 ? n0 / p.grp_n : 0;
                            if (p.a_mode == 0) {
                                tma_load_2d_cg2(sa, &p.tmA, fb, kb * BLOCK_K + grp * p.a_grp_koff, m0);
                            } else if (p.a_mode == 1) {
                                const int am = m0 + kseg * p.a_seg_off;
                                tma_load_2d_cg2(sa, &p.tmA, fb, am, kbs * BLOCK_K);
                                tma_load_2d_cg2(sa + 8192, &p.tmA, fb, am + 64, kbs * BLOCK_K);
                            } else {
                                const int tap = kb / p.kb_per_tap;
                                const int cb = kb - tap * p.kb_per_tap;
                                if (cb < p.kb_src1)
                                    tma_load_4d_cg2(sa, &p.tmA, fb, cb * BLOCK_K, cw + p.tap_dw[tap], ch + p.tap_dh[tap], cn + p.tap_dn[tap]);
                                else
                                    tma_load_4d_cg2(sa, &p.tmA2, fb, (cb - p.kb_src1) * BLOCK_K, cw + p.tap_dw[tap], ch + p.tap_dh[tap],
                                                    cn + p.tap_dn[tap]);
                            }
                            if (p.epi_mode == 1)  // leader stages the 128 h rows of W, the peer the matching 128 g rows
                                tma_load_2d_cg2(sb, &p.tmBh, fb, kb * BLOCK_K, (int)cta_rank * p.geglu_f + n_blk * 128);
                            else if (p.b_mode == 0)
                                for (int i = 0; i < p.n_inst; ++i)  // instruction i covers columns [i * bn_inst, (i + 1) * bn_inst): half per CTA
                                    tma_load_2d_cg2(sb + i * p.b_inst_bytes, &p.tmBh, fb, kb * BLOCK_K,
                                                    n0 - grp * p.grp_n + i * p.bn_inst + (int)cta_rank * (p.bn_inst >> 1));
                            else
                                tma_load_3d_cg2(sb, &p.tmBh, fb, 0, kbs * BLOCK_K,
                                                ((n0 - grp * p.grp_n + kseg * p.b_seg_off) >> 6) + (int)cta_rank * (p.block_n >> 7));
                        } else {
                            mbar_expect_tx(&full_bar[stage], tx_bytes);
                            // ---- A ----
                            int kseg = 0, kbs = kb;  // segment and k-block within it
                            if (p.kb_per_seg > 0) {
                                kseg = kb / p.kb_per_seg;
                                kbs = kb - kseg * p.kb_per_seg;
                            }
                            const int grp = p.grp_n > 0 ? n0 / p.grp_n : 0;
                            if (p.a_mode == 0) {
                                tma_load_2d(sa, &p.tmA, &full_bar[stage], kb * BLOCK_K + grp * p.a_grp_koff, m0);
                            } else if (p.a_mode == 1) {
                                const int am = m0 + kseg * p.a_seg_off;
                                tma_load_2d(sa, &p.tmA, &full_bar[stage], am, kbs * BLOCK_K);
                                tma_load_2d(sa + 8192, &p.tmA, &full_bar[stage], am + 64, kbs * BLOCK_K);
                            } else {
                                const int tap = kb / p.kb_per_tap;
                                const int cb = kb - tap * p.kb_per_tap;
                                if (cb < p.kb_src1)
                                    tma_load_4d(sa, &p.tmA, &full_bar[stage], cb * BLOCK_K, cw + p.tap_dw[tap],
                                                ch + p.tap_dh[tap], cn + p.tap_dn[tap]);
                                else
                                    tma_load_4d(sa, &p.tmA2, &full_bar[stage], (cb - p.kb_src1) * BLOCK_K, cw + p.tap_dw[tap],
                                                ch + p.tap_dh[tap], cn + p.tap_dn[tap]);
                            }
                            // ---- B ----
                            if (p.cluster2) {
                                // this CTA fetches its half of the shared B tile and multicasts it into both CTAs' stage
                                const int rk = wi.rank;
                                if (p.b_mode == 0)
                                    tma_load_2d_mc(sb + rk * p.b_half_bytes, &p.tmBh, &full_bar[stage], kb * BLOCK_K,
                                                   n0 - grp * p.grp_n + rk * (p.block_n >> 1), (uint16_t)3);
                                else
                                    tma_load_3d_mc(sb + rk * p.b_half_bytes, &p.tmBh, &full_bar[stage], 0, kb * BLOCK_K,
                                                   (n0 >> 6) + rk * (p.block_n >> 7), (uint16_t)3);
                            } else if (p.b_mode == 0) {
                                tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BLOCK_K, n0 - grp * p.grp_n);
                            } else {
                                const int bn0 = n0 - grp * p.grp_n + kseg * p.b_seg_off;
                                if (p.b_3d) {
                                    tma_load_3d(sb, &p.tmB, &full_bar[stage], 0, kbs * BLOCK_K, bn0 >> 6);
                                } else {
                                    for (int j = 0; j * 64 < p.block_n; ++j)
                                        tma_load_2d(sb + j * 8192, &p.tmB, &full_bar[stage], bn0 + j * 64, kbs * BLOCK_K);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        // The whole warp walks the loop (warp-uniform control flow and operands: the descriptors stay in uniform registers);
        // lane 0 issues.  The per-k-block instruction count of this loop bounds the kernel when tiles are narrow: one thread has
        // ~block_n * 2 cycles per k-block to wait, issue 4 (8) UMMAs and commit, so everything that can be hoisted is.
        if (cta_rank == 0) {
            const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, (uint32_t)(PAIR ? p.bn_inst : p.block_n), p.a_mode == 1,
                                                   p.b_mode == 1);
            const uint32_t smem_base = smem_u32(smem);
            const uint64_t adesc0 = make_smem_desc(smem_base, p.a_lbo, p.a_sbo);
            const uint64_t bdesc0 = make_smem_desc(smem_base + A_STAGE_BYTES, p.b_lbo, p.b_sbo);
            const uint32_t a_kinc = p.a_kadv >> 4, b_kinc = p.b_kadv >> 4, stage_inc = (uint32_t)stage_bytes >> 4;
            const uint32_t b_iinc = (uint32_t)p.b_inst_bytes >> 4;
            const int n_stages = p.stages;
            int stage = 0;
            uint32_t stage_off = 0;  // (stage * stage_bytes) >> 4: added to the 14-bit start-address field of both descriptors
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            WorkIter wi;
            wi.init(p, num_tiles);
            while (wi.next()) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
                // instructions whose columns lie entirely past N are skipped (ragged last tile)
                int n_inst_tile = 1;
                if (PAIR && p.n_inst > 1) {
                    const int n0 = (wi.tile % p.num_n_tiles) * p.block_n;
                    n_inst_tile = min(p.n_inst, (p.N - n0 + p.bn_inst - 1) / p.bn_inst);
                }
                uint32_t accf = 0;  // first UMMA of the tile overwrites the accumulator
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = adesc0 + stage_off, bdesc = bdesc0 + stage_off;
                    if (elect_one()) {
                        if (PAIR) {
                            if (n_inst_tile == 1) {
#pragma unroll
                                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                    umma_bf16_cg2(d_tmem, adesc + k * a_kinc, bdesc + k * b_kinc, idesc, k == 0 ? accf : 1u);
                            } else {
#pragma unroll
                                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                    umma_bf16_cg2(d_tmem, adesc + k * a_kinc, bdesc + k * b_kinc, idesc, k == 0 ? accf : 1u);
                                    umma_bf16_cg2(d_tmem + (uint32_t)p.bn_inst, adesc + k * a_kinc, bdesc + k * b_kinc + b_iinc, idesc,
                                                  k == 0 ? accf : 1u);
                                }
                            }
                            umma_commit_cg2(&empty_bar[stage], (uint16_t)3);  // frees this stage in BOTH CTAs once the MMAs retire
                        } else {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma_bf16(d_tmem, adesc + k * a_kinc, bdesc + k * b_kinc, idesc, k == 0 ? accf : 1u);
                            if (p.cluster2)
                                umma_commit_multicast(&empty_bar[stage], (uint16_t)3);  // both producers may refill once we are done
                            else
                                umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs retire
                        }
                    }
                    accf = 1;
                    stage_off += stage_inc;
                    if (++stage == n_stages) {
                        stage = 0;
                        stage_off = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) {
                    if (PAIR)
                        umma_commit_cg2(&tmem_full[acc], (uint16_t)3);  // both CTAs' epilogues own half of the rows
                    else
                        umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++acc == p.n_acc) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================================== epilogue =====================================
        // warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        const int sub = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        WorkIter wi;
        wi.init(p, num_tiles);
        if (p.epi_tma) {
            // ---- bf16 output: TMEM -> registers -> fused epilogue -> 128B-swizzled smem -> TMA store ----
            // Each warp owns 32 rows and works in 64-column chunks through two private 4 KiB staging buffers; the store of
            // chunk c overlaps the math of chunk c+1.  Residual rows are read with fully coalesced 16-byte loads (4 rows
            // per warp instruction, one chunk ahead), transposed through the same staging buffer.
            uint8_t* stg = epi_stage + sub * p.epi_warp_bytes;
            const int lrow = lane >> 3, lchk = lane & 7;  // coalesced residual layout: 4 rows x 8 16-byte chunks per instruction
            int buf = 0;
            if (PAIR && p.epi_mode == 1) {
                // ---- GEGLU forward: pre-activation (h | g) -> `out`, h * gelu(g) -> `out2`; three staged TMA stores per
                // 64 hidden units (diffusers GEGLU.forward: hidden, gate = proj(x).chunk(2); hidden * gelu(gate)) ----
                uint8_t *sH = stg, *sG = stg + 4096, *sA = stg + 8192;
                while (wi.next()) {
                    const int tile = wi.tile;
                    const int n_blk = tile % p.num_n_tiles;
                    const int m_blk = tile / p.num_n_tiles;
                    const int row0 = m_blk * BLOCK_M + sub * 32;
                    const int hcol0 = n_blk * 128;
                    mbar_wait(&tmem_full[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
                    for (int c = 0; c < 2; ++c) {
#pragma unroll 1
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t rh[32], rg[32];
                            tmem_ld32(t_row + (uint32_t)(c * 64 + hh * 32), rh);
                            tmem_ld32(t_row + (uint32_t)(128 + c * 64 + hh * 32), rg);
                            tmem_ld_wait();
                            if (c == 1 && hh == 1) {  // accumulator fully read
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                            }
                            if (hh == 0) {
                                if (lane == 0) bulk_wait_group_read0();  // the previous chunk's three stores have read their buffers
                                __syncwarp();
                            }
                            const int cc = hcol0 + c * 64 + hh * 32;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float h[8], g[8], a[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    h[j] = __uint_as_float(rh[q * 8 + j]) * p.alpha;
                                    g[j] = __uint_as_float(rg[q * 8 + j]) * p.alpha;
                                }
                                if (p.bias != nullptr) {
                                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cc + q * 8));
                                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cc + q * 8 + 4));
                                    const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.bias + p.geglu_f + cc + q * 8));
                                    const float4 c1 = __ldg(reinterpret_cast<const float4*>(p.bias + p.geglu_f + cc + q * 8 + 4));
                                    h[0] += b0.x; h[1] += b0.y; h[2] += b0.z; h[3] += b0.w; h[4] += b1.x; h[5] += b1.y; h[6] += b1.z; h[7] += b1.w;
                                    g[0] += c0.x; g[1] += c0.y; g[2] += c0.z; g[3] += c0.w; g[4] += c1.x; g[5] += c1.y; g[6] += c1.z; g[7] += c1.w;
                                }
                                uint4 oh, og, oa;
                                oh.x = pack_bf16(h[0], h[1]); oh.y = pack_bf16(h[2], h[3]); oh.z = pack_bf16(h[4], h[5]); oh.w = pack_bf16(h[6], h[7]);
                                og.x = pack_bf16(g[0], g[1]); og.y = pack_bf16(g[2], g[3]); og.z = pack_bf16(g[4], g[5]); og.w = pack_bf16(g[6], g[7]);
                                // the activation is formed from the bf16-rounded pre-activation, exactly what the unfused
                                // kernel (and the backward pass, which re-reads it) sees
                                float2 t;
                                t = unpack_bf16(oh.x); h[0] = t.x; h[1] = t.y; t = unpack_bf16(oh.y); h[2] = t.x; h[3] = t.y;
                                t = unpack_bf16(oh.z); h[4] = t.x; h[5] = t.y; t = unpack_bf16(oh.w); h[6] = t.x; h[7] = t.y;
                                t = unpack_bf16(og.x); g[0] = t.x; g[1] = t.y; t = unpack_bf16(og.y); g[2] = t.x; g[3] = t.y;
                                t = unpack_bf16(og.z); g[4] = t.x; g[5] = t.y; t = unpack_bf16(og.w); g[6] = t.x; g[7] = t.y;
#pragma unroll
                                for (int j = 0; j < 8; ++j) a[j] = h[j] * gelu_fast_f(g[j]);
                                oa.x = pack_bf16(a[0], a[1]); oa.y = pack_bf16(a[2], a[3]); oa.z = pack_bf16(a[4], a[5]); oa.w = pack_bf16(a[6], a[7]);
                                const int off = lane * 128 + (((hh * 4 + q) ^ (lane & 7)) << 4);
                                *reinterpret_cast<uint4*>(sH + off) = oh;
                                *reinterpret_cast<uint4*>(sG + off) = og;
                                *reinterpret_cast<uint4*>(sA + off) = oa;
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&p.tmO, sH, hcol0 + c * 64, row0);
                            tma_store_2d(&p.tmO, sG, p.geglu_f + hcol0 + c * 64, row0);
                            tma_store_2d(&p.tmO2, sA, hcol0 + c * 64, row0);
                            bulk_commit_group();
                        }
                    }
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
                if (lane == 0) bulk_wait_group0();
            } else if (PAIR && p.epi_mode == 2) {
                // ---- GEGLU backward: the tile is d = dL/d(h * gelu(g)); with h, g from the saved pre-activation (TMA loads into
                // the warp's staging buffers): dh = d * gelu(g) -> out[:, n], dg = d * h * gelu'(g) -> out[:, F + n] ----
                uint8_t *iH = stg, *iG = stg + 4096, *oH = stg + 8192, *oG = stg + 12288;
                uint32_t ephase = 0;
                while (wi.next()) {
                    const int tile = wi.tile;
                    const int n_blk = tile % p.num_n_tiles;
                    const int m_blk = tile / p.num_n_tiles;
                    const int row0 = m_blk * BLOCK_M + sub * 32;
                    const int n0 = n_blk * p.block_n;
                    const int nchunks = p.block_n >> 6;
                    if (elect_one()) {
                        mbar_expect_tx(&epi_bar[sub], 8192);
                        tma_load_2d(iH, &p.tmAux, &epi_bar[sub], n0, row0);
                        tma_load_2d(iG, &p.tmAux, &epi_bar[sub], p.geglu_f + n0, row0);
                    }
                    mbar_wait(&tmem_full[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
                    for (int ch = 0; ch < nchunks; ++ch) {
                        const int col0 = n0 + ch * 64;
                        uint32_t r[2][32];
                        tmem_ld32(t_row + (uint32_t)(ch * 64), r[0]);
                        tmem_ld32(t_row + (uint32_t)(ch * 64 + 32), r[1]);
                        tmem_ld_wait();
                        if (ch == nchunks - 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                        }
                        mbar_wait(&epi_bar[sub], ephase);
                        ephase ^= 1;
                        uint4 vh[8], vg[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int off = lane * 128 + ((j ^ (lane & 7)) << 4);
                            vh[j] = *reinterpret_cast<const uint4*>(iH + off);
                            vg[j] = *reinterpret_cast<const uint4*>(iG + off);
                        }
                        __syncwarp();  // every lane has its h / g rows in registers: the input buffers may be refilled
                        fence_proxy_async();
                        if (ch + 1 < nchunks && elect_one()) {
                            mbar_expect_tx(&epi_bar[sub], 8192);
                            tma_load_2d(iH, &p.tmAux, &epi_bar[sub], col0 + 64, row0);
                            tma_load_2d(iG, &p.tmAux, &epi_bar[sub], p.geglu_f + col0 + 64, row0);
                        }
                        if (lane == 0) bulk_wait_group_read0();  // the previous chunk's stores have read the output buffers
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float h[8], g[8], dh[8], dg[8];
                            float2 t;
                            t = unpack_bf16(vh[j].x); h[0] = t.x; h[1] = t.y; t = unpack_bf16(vh[j].y); h[2] = t.x; h[3] = t.y;
                            t = unpack_bf16(vh[j].z); h[4] = t.x; h[5] = t.y; t = unpack_bf16(vh[j].w); h[6] = t.x; h[7] = t.y;
                            t = unpack_bf16(vg[j].x); g[0] = t.x; g[1] = t.y; t = unpack_bf16(vg[j].y); g[2] = t.x; g[3] = t.y;
                            t = unpack_bf16(vg[j].z); g[4] = t.x; g[5] = t.y; t = unpack_bf16(vg[j].w); g[6] = t.x; g[7] = t.y;
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                // the unfused path rounds d to bf16 between the GEMM and the GEGLU kernel: same here
                                const float d = __bfloat162float(__float2bfloat16(__uint_as_float(r[j >> 2][(j & 3) * 8 + e]) * p.alpha));
                                float cdf, pdf;
                                gelu_cdf_pdf(g[e], cdf, pdf);
                                dh[e] = d * g[e] * cdf;
                                dg[e] = d * h[e] * fmaf(g[e], pdf, cdf);
                            }
                            uint4 uh, ug;
                            uh.x = pack_bf16(dh[0], dh[1]); uh.y = pack_bf16(dh[2], dh[3]); uh.z = pack_bf16(dh[4], dh[5]); uh.w = pack_bf16(dh[6], dh[7]);
                            ug.x = pack_bf16(dg[0], dg[1]); ug.y = pack_bf16(dg[2], dg[3]); ug.z = pack_bf16(dg[4], dg[5]); ug.w = pack_bf16(dg[6], dg[7]);
                            const int off = lane * 128 + ((j ^ (lane & 7)) << 4);
                            *reinterpret_cast<uint4*>(oH + off) = uh;
                            *reinterpret_cast<uint4*>(oG + off) = ug;
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&p.tmO, oH, col0, row0);
                            tma_store_2d(&p.tmO, oG, p.geglu_f + col0, row0);
                            bulk_commit_group();
                        }
                    }
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
                if (lane == 0) bulk_wait_group0();
            } else {
            uint8_t* rstg = stg + 8192;          // [4 chunks][32 rows x 128 B] residual slots of this warp (res_tma)
            uint64_t* rbar = res_bar + sub * 4;
            uint32_t rph = 0;                    // phase bit of each chunk's barrier
            auto issue_res = [&](int tile_n, int ch) {  // one lane
                const int n0n = (tile_n % p.num_n_tiles) * p.block_n;
                if (n0n + ch * 64 < min(p.N, n0n + p.block_n)) {
                    mbar_expect_tx(&rbar[ch], 4096);
                    tma_load_2d(rstg + ch * 4096, &p.tmAux, &rbar[ch], n0n + ch * 64, (tile_n / p.num_n_tiles) * BLOCK_M + sub * 32);
                }
            };
            if (p.res_tma && lane == 0) {
                WorkIter w0 = wi;
                if (w0.next())
                    for (int ch = 0; ch < 4; ++ch) issue_res(w0.tile, ch);
            }
            while (wi.next()) {
                const int tile = wi.tile;
                const int n_blk = tile % p.num_n_tiles;
                const int m_blk = tile / p.num_n_tiles;
                const int row0 = m_blk * BLOCK_M + sub * 32;
                const int row = row0 + lane;
                const int n0 = n_blk * p.block_n;
                const int n_end = min(p.N, n0 + p.block_n);
                const int nchunks = (n_end - n0 + 63) >> 6;
                const bool row_ok = row < p.M;
                const float* brow = (p.bias_rows != nullptr && row_ok) ? p.bias_rows + (size_t)(row / p.rows_per_bias) * p.N : nullptr;
                // residual by TMA: every 64-column chunk of the NEXT tile is requested as soon as this tile has consumed the
                // chunk's slot, so the load has a whole tile epilogue to land (the register path below prefetches one chunk)
                WorkIter wl = wi;
                const bool has_next = p.res_tma && wl.next();
                const int tile_next = wl.tile;
                uint4 rnext[8];
                auto load_res = [&](int col0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = row0 + lrow + 4 * i, cc = col0 + lchk * 8;
                        rnext[i] = (r < p.M && cc < n_end) ? *reinterpret_cast<const uint4*>(p.residual + (size_t)r * p.ldr + cc)
                                                           : make_uint4(0, 0, 0, 0);
                    }
                };
                if (p.residual != nullptr && !p.res_tma) load_res(n0);
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int col0 = n0 + ch * 64;
                    uint32_t r[2][32];
                    tmem_ld32(t_row + (uint32_t)(ch * 64), r[0]);
                    tmem_ld32(t_row + (uint32_t)(ch * 64 + 32), r[1]);
                    tmem_ld_wait();
                    if (ch == nchunks - 1) {  // accumulator fully read: hand it back to the MMA warp before the stores
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (PAIR)
                                mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                            else
                                mbar_arrive(&tmem_empty[acc]);
                        }
                    }
                    uint8_t* sbuf = stg + buf * 4096;
                    // the TMA store that last used this buffer (two chunks ago) must have finished reading it
                    if (lane == 0) bulk_wait_group_read1();
                    __syncwarp();
                    uint4 rc[8];
                    if (p.res_tma) {
                        mbar_wait(&rbar[ch], (rph >> ch) & 1u);
                        rph ^= 1u << ch;
                        const uint8_t* rs = rstg + ch * 4096;
#pragma unroll
                        for (int j = 0; j < 8; ++j) rc[j] = *reinterpret_cast<const uint4*>(rs + lane * 128 + ((j ^ (lane & 7)) << 4));
                        __syncwarp();
                        if (has_next && lane == 0) {
                            fence_proxy_async();
                            issue_res(tile_next, ch);
                        }
                    } else if (p.residual != nullptr) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = lrow + 4 * i;
                            *reinterpret_cast<uint4*>(sbuf + rr * 128 + ((lchk ^ (rr & 7)) << 4)) = rnext[i];
                        }
                        if (ch + 1 < nchunks) load_res(col0 + 64);  // next chunk's residual in flight during this chunk's math
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j) rc[j] = *reinterpret_cast<const uint4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4));
                    }
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int c32 = col0 + hh * 32;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[hh][j]) * p.alpha;
                        if (c32 < n_end) {
                            const bool full = c32 + 32 <= n_end;
                            if (p.bias != nullptr) {
                                const float* bp = p.bias + c32;
                                if (full && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        const float4 t = __ldg(reinterpret_cast<const float4*>(bp) + q);
                                        v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (c32 + j < n_end) v[j] += __ldg(bp + j);
                                }
                            }
                            if (brow != nullptr) {
                                const float* bp = brow + c32;
                                if (full && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        const float4 t = __ldg(reinterpret_cast<const float4*>(bp) + q);
                                        v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (c32 + j < n_end) v[j] += __ldg(bp + j);
                                }
                            }
                        }
                        if (p.residual != nullptr) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 u = rc[hh * 4 + q];
                                const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
                                v[q * 8 + 0] += f0.x; v[q * 8 + 1] += f0.y; v[q * 8 + 2] += f1.x; v[q * 8 + 3] += f1.y;
                                v[q * 8 + 4] += f2.x; v[q * 8 + 5] += f2.y; v[q * 8 + 6] += f3.x; v[q * 8 + 7] += f3.y;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 o;
                            o.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
                            o.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
                            o.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
                            o.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
                            const int j = hh * 4 + q;
                            *reinterpret_cast<uint4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                        }
                    }
                    // a 64-wide box may only be stored when it stays inside this tile's columns (or runs off the end of
                    // the tensor, where TMA clips); the narrow tail chunk of tiles with block_n % 64 != 0 is stored directly
                    const bool box_ok = (col0 + 64 <= n0 + p.block_n) || (n0 + p.block_n >= p.N);
                    if (box_ok) {
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            if (col0 >= p.n_split)
                                tma_store_2d(&p.tmO2, sbuf, col0 - p.n_split, row0);
                            else
                                tma_store_2d(&p.tmO, sbuf, col0, row0);
                            bulk_commit_group();
                        }
                        buf ^= 1;
                    } else if (row_ok) {
                        __nv_bfloat16* op = (col0 >= p.n_split)
                                                ? reinterpret_cast<__nv_bfloat16*>(p.out2) + (size_t)row * p.ldo2 + (col0 - p.n_split)
                                                : reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col0;
                        const int w8 = (n_end - col0) >> 3;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < w8)
                                *reinterpret_cast<uint4*>(op + j * 8) =
                                    *reinterpret_cast<const uint4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4));
                    }
                }
                if (has_next && lane == 0)  // chunks the next tile has and this (ragged) one did not
                    for (int ch = nchunks; ch < 4; ++ch) issue_res(tile_next, ch);
                if (++acc == p.n_acc) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            if (lane == 0) bulk_wait_group0();  // all stores have landed before the CTA releases its shared memory
            }
        } else
        while (wi.next()) {
            const int tile = wi.tile;
            const int n_blk = tile % p.num_n_tiles;
            const int m_blk = tile / p.num_n_tiles;
            const int row = m_blk * BLOCK_M + sub * 32 + lane;
            const int n0 = n_blk * p.block_n;
            const int n_end = min(p.N, n0 + p.block_n);
            const bool row_ok = row < p.M;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
            const float* brow = (p.bias_rows != nullptr && row_ok)
                                    ? p.bias_rows + (size_t)(row / p.rows_per_bias) * p.N
                                    : nullptr;
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t r[32];
                tmem_ld32(t_row + (uint32_t)c, r);
                tmem_ld_wait();
                const int col0 = n0 + c;
                if (!row_ok || col0 >= n_end) continue;
                // destination (out or out2 after the split column)
                uint8_t* obase;
                long long ld;
                int ocol;
                if (col0 >= p.n_split) {
                    obase = reinterpret_cast<uint8_t*>(p.out2);
                    ld = p.ldo2;
                    ocol = col0 - p.n_split;
                } else {
                    obase = reinterpret_cast<uint8_t*>(p.out);
                    ld = p.ldo;
                    ocol = col0;
                }
                const int ncols = min(32, n_end - col0);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
                const bool vec_ok = (ncols == 32);
                if (p.bias != nullptr) {
                    const float* bp = p.bias + col0;
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(bp) + q);
                            v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) v[j] += __ldg(bp + j);
                    }
                }
                if (brow != nullptr) {
                    const float* bp = brow + col0;
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(bp) + q);
                            v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) v[j] += __ldg(bp + j);
                    }
                }
                if (p.residual != nullptr) {
                    const __nv_bfloat16* rp = p.residual + (size_t)row * p.ldr + col0;
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 u = *reinterpret_cast<const uint4*>(rp + q * 8);
                            float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z),
                                   f3 = unpack_bf16(u.w);
                            v[q * 8 + 0] += f0.x; v[q * 8 + 1] += f0.y; v[q * 8 + 2] += f1.x; v[q * 8 + 3] += f1.y;
                            v[q * 8 + 4] += f2.x; v[q * 8 + 5] += f2.y; v[q * 8 + 6] += f3.x; v[q * 8 + 7] += f3.y;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) v[j] += __bfloat162float(rp[j]);
                    }
                }
                if (p.stream_k) {
                    float* op = reinterpret_cast<float*>(obase) + (size_t)row * ld + ocol;
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) red_add_v4(op + q * 4, v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) atomicAdd(op + j, v[j]);
                    }
                } else if (p.out_fp32) {
                    float* op = reinterpret_cast<float*>(obase) + (size_t)row * ld + ocol;
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 o = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                            if (p.accumulate) {
                                float4 old = *reinterpret_cast<float4*>(op + q * 4);
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            *reinterpret_cast<float4*>(op + q * 4) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) op[j] = p.accumulate ? op[j] + v[j] : v[j];
                    }
                } else {
                    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(obase) + (size_t)row * ld + ocol;
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 o;
                            o.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
                            o.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
                            o.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
                            o.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
                            *reinterpret_cast<uint4*>(op + q * 8) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) op[j] = __float2bfloat16(v[j]);
                    }
                }
            }
            // release the accumulator buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR)
                    mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                else
                    mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == p.n_acc) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR || p.cluster2) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it / arrive on its barriers
    if (warp == 2) {
        tc_fence_after();
        if (PAIR)
            tmem_dealloc2(tmem_base, TMEM_COLS);
        else
            tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static bool num_tiles_ge(const GemmKernelArgs& p, int n) { return p.num_m_tiles * p.num_n_tiles >= n; }

static int pick_block_n(long long N) {
    if (N <= 16) return 16;
    if (N <= 32) return 32;
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    if (N % 256 == 0) return 256;
    // measured (profiles/r01_gemm_block_n_sweep.log): from N = 512 up, 256-wide tiles with a partial last tile beat the
    // exact divisors 160 / 192 (N = 640: 915 vs 586 TFLOP/s, N = 1920: 1110 vs 832), which also need per-chunk loads for
    // MN-major B and a direct-store tail in the epilogue
    if (N >= 512) return 256;
    if (N % 160 == 0) return 160;
    if (N % 128 == 0) return 128;
    if (N % 192 == 0) return 192;
    return 256;
}

}  // namespace uwu

using namespace uwu;

extern "C" int uwu_gemm(const uwu_gemm_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(d != nullptr, "uwu_gemm: null descriptor");
    UWU_CHECK_ARG(d->a && d->b && d->out, "uwu_gemm: null operand pointer");
    UWU_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "uwu_gemm: non-positive shape M=%lld N=%lld K=%lld",
                  (long long)d->M, (long long)d->N, (long long)d->K);
    UWU_CHECK_ARG(d->M < (1ll << 31) && d->N < (1ll << 31) && d->K < (1ll << 31), "uwu_gemm: shape exceeds int32");
    UWU_CHECK_ARG(d->a_layout >= 0 && d->a_layout <= 2, "uwu_gemm: bad a_layout %d", d->a_layout);
    UWU_CHECK_ARG(d->b_layout >= 0 && d->b_layout <= 1, "uwu_gemm: bad b_layout %d", d->b_layout);
    UWU_CHECK_ARG(d->out_dtype == UWU_BF16 || d->out_dtype == UWU_F32, "uwu_gemm: bad out dtype %d", d->out_dtype);
    UWU_CHECK_ARG(!(d->accumulate && d->out_dtype != UWU_F32), "uwu_gemm: accumulate requires fp32 output");

    static thread_local GemmKernelArgs args;  // large (3 tensor maps); avoid re-zeroing cost on the stack
    GemmKernelArgs& p = args;
    p.M = (int)d->M;
    p.N = (int)d->N;
    int bn = d->block_n > 0 ? d->block_n : pick_block_n(d->N);
    p.epi_mode = d->epi_mode;
    p.geglu_f = 0;
    p.epi_warp_bytes = 8192;
    if (d->epi_mode != 0) {
        UWU_CHECK_ARG(d->epi_mode == 1 || d->epi_mode == 2, "uwu_gemm: bad epi_mode %d", d->epi_mode);
        UWU_CHECK_ARG(d->out_dtype == UWU_BF16 && !d->residual && !d->bias_rows && !d->accumulate && d->k_segs <= 1 && d->grp_n == 0,
                      "uwu_gemm: the fused GEGLU epilogues take a bf16 output and no residual / bias_rows / accumulate");
        bn = 256;
        if (d->epi_mode == 1) {
            UWU_CHECK_ARG(d->N % 256 == 0 && d->b_layout == UWU_B_NK && d->out2 != nullptr,
                          "uwu_gemm(GEGLU fwd): N = 2F with F a multiple of 128, K-major weights and out2 are required");
            p.geglu_f = (int)(d->N / 2);
            p.epi_warp_bytes = 12288;
        } else {
            UWU_CHECK_ARG(d->N % 256 == 0 && d->aux != nullptr && d->ld_aux % 8 == 0 && d->out2 == nullptr && !d->bias,
                          "uwu_gemm(GEGLU bwd): N = F must be a multiple of 256; aux (the saved pre-activation) is required");
            p.geglu_f = (int)d->N;
            p.epi_warp_bytes = 16384;
        }
    }
    // wide tiles: 320 columns as two 160-wide UMMAs per k-step sharing the A tile (CTA-pair kernel, K-major weights, long
    // reductions: the 3x3 convolutions, whose N = 320 / 640 / 1280 otherwise run 160-wide or ragged 256-wide tiles)
    int n_inst = 1;
    {
        static int want_wide = -1, pair_env = 1;
        if (want_wide < 0) {
            const char* e = getenv("UWU_GEMM_WIDE");
            want_wide = e ? atoi(e) : 1;
            const char* e2 = getenv("UWU_GEMM_PAIR");
            pair_env = e2 ? atoi(e2) : 2;
        }
        const long long kblocks = (d->K + 63) / 64;
        if (want_wide && pair_env && sm_count() % 2 == 0 && d->block_n == 0 && d->epi_mode == 0 && d->b_layout == UWU_B_NK &&
            d->k_segs <= 1 && d->grp_n == 0 && d->M >= 256 && d->N % 320 == 0 && d->out_dtype == UWU_BF16 && d->out2 == nullptr &&
            ((d->N <= 640 && kblocks >= 40) || kblocks >= 90)) {
            bn = 320;
            n_inst = 2;
        }
    }
    UWU_CHECK_ARG(bn % 16 == 0 && bn >= 16 && (bn <= 256 || n_inst > 1), "uwu_gemm: block_n %d must be a multiple of 16 in [16,256]", bn);
    p.block_n = bn;
    p.num_m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    p.num_n_tiles = d->epi_mode == 1 ? p.geglu_f / 128 : (p.N + bn - 1) / bn;
    p.a_mode = d->a_layout;
    p.b_mode = d->b_layout;
    p.kb_per_seg = 0; p.a_seg_off = 0; p.b_seg_off = 0; p.b_3d = 0;
    p.grp_n = 0; p.a_grp_koff = 0;
    if (d->grp_n > 0) {
        UWU_CHECK_ARG(d->a_layout == UWU_A_ROW, "uwu_gemm: grp_n needs a row-major A");  // B: [K, grp_n] (KN) or [grp_n, K] (NK)
        UWU_CHECK_ARG(d->grp_n % bn == 0, "uwu_gemm: block_n %d must divide grp_n %d", bn, d->grp_n);
        p.grp_n = d->grp_n;
        p.a_grp_koff = d->a_grp_koff;
    }

    // ---------------- A tensor map(s) ----------------
    if (d->a_layout == UWU_A_ROW) {
        UWU_CHECK_ARG(d->lda % 8 == 0 && d->lda >= d->K, "uwu_gemm: lda %lld must be >= K and a multiple of 8", (long long)d->lda);
        const int n_grp = d->grp_n > 0 ? (int)((d->N + d->grp_n - 1) / d->grp_n) : 1;
        uint64_t dims[2] = {(uint64_t)d->K + (uint64_t)(n_grp - 1) * (uint64_t)d->a_grp_koff, (uint64_t)d->M};
        UWU_CHECK_ARG((int64_t)dims[0] <= d->lda, "uwu_gemm: grouped A columns exceed lda");
        uint64_t str[1] = {(uint64_t)d->lda * 2};
        uint32_t box[2] = {BLOCK_K, BLOCK_M};
        if (encode_tmap_bf16(&p.tmA, d->a, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
        p.tmA2 = p.tmA;
        p.num_kb = (int)((d->K + BLOCK_K - 1) / BLOCK_K);
        p.a_lbo = 0; p.a_sbo = 1024; p.a_kadv = 32;
    } else if (d->a_layout == UWU_A_COL) {
        UWU_CHECK_ARG(d->lda % 8 == 0 && d->lda >= d->M, "uwu_gemm: lda %lld must be >= M and a multiple of 8", (long long)d->lda);
        const int segs = d->k_segs > 1 ? d->k_segs : 1;
        UWU_CHECK_ARG(segs == 1 || d->b_layout == UWU_B_KN, "uwu_gemm: k_segs needs the A_COL x B_KN (token-reduction) form");
        uint64_t dims[2] = {(uint64_t)d->M + (uint64_t)(segs - 1) * (uint64_t)d->a_seg_off, (uint64_t)d->K};
        UWU_CHECK_ARG((int64_t)dims[0] <= d->lda, "uwu_gemm: segmented A columns exceed lda");
        uint64_t str[1] = {(uint64_t)d->lda * 2};
        uint32_t box[2] = {64, BLOCK_K};
        if (encode_tmap_bf16(&p.tmA, d->a, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
        p.tmA2 = p.tmA;
        p.num_kb = (int)((d->K + BLOCK_K - 1) / BLOCK_K);
        p.a_lbo = 8192; p.a_sbo = 1024; p.a_kadv = 2048;
        if (segs > 1) {
            p.kb_per_seg = p.num_kb;
            p.num_kb *= segs;
            p.a_seg_off = d->a_seg_off;
            p.b_seg_off = d->b_seg_off;
        }
    } else {
        // NHWC activation(s): [n_img_buf, H, W, C]
        const int H = d->H, W = d->W, C1 = d->Cin1, C2 = d->Cin2;
        UWU_CHECK_ARG(H > 0 && W > 0 && C1 > 0 && C2 >= 0, "uwu_gemm(conv): bad geometry H=%d W=%d C1=%d C2=%d", H, W, C1, C2);
        UWU_CHECK_ARG(C1 % 64 == 0 && C2 % 64 == 0, "uwu_gemm(conv): channel counts must be multiples of 64 (got %d,%d)", C1, C2);
        UWU_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= 9, "uwu_gemm(conv): ntaps %d out of range", d->ntaps);
        UWU_CHECK_ARG(d->K == (long long)d->ntaps * (C1 + C2), "uwu_gemm(conv): K %lld != ntaps*(C1+C2)", (long long)d->K);
        UWU_CHECK_ARG(C2 == 0 || d->a2 != nullptr, "uwu_gemm(conv): Cin2 > 0 but a2 is null");
        int bw, bh, bnimg;
        if (W >= 128) {
            UWU_CHECK_ARG(W % 128 == 0, "uwu_gemm(conv): W=%d must be a multiple of 128 when >= 128", W);
            bw = 128; bh = 1; bnimg = 1;
        } else {
            UWU_CHECK_ARG(128 % W == 0, "uwu_gemm(conv): W=%d must divide 128", W);
            bw = W;
            int rows = 128 / W;
            if (H >= rows) {
                UWU_CHECK_ARG(H % rows == 0, "uwu_gemm(conv): H=%d incompatible with 128-pixel tiles", H);
                bh = rows; bnimg = 1;
            } else {
                UWU_CHECK_ARG(rows % H == 0, "uwu_gemm(conv): H=%d incompatible with 128-pixel tiles", H);
                bh = H; bnimg = rows / H;
            }
        }
        const int nbuf = d->n_img_buf;
        UWU_CHECK_ARG(nbuf > 0, "uwu_gemm(conv): n_img_buf must be > 0");
        {
            uint64_t dims[4] = {(uint64_t)C1, (uint64_t)W, (uint64_t)H, (uint64_t)nbuf};
            uint64_t str[3] = {(uint64_t)C1 * 2, (uint64_t)C1 * W * 2, (uint64_t)C1 * W * H * 2};
            uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bnimg};
            if (encode_tmap_bf16(&p.tmA, d->a, 4, dims, str, box, 1)) return UWU_ERR_INVALID;
        }
        if (C2 > 0) {
            uint64_t dims[4] = {(uint64_t)C2, (uint64_t)W, (uint64_t)H, (uint64_t)nbuf};
            uint64_t str[3] = {(uint64_t)C2 * 2, (uint64_t)C2 * W * 2, (uint64_t)C2 * W * H * 2};
            uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bnimg};
            if (encode_tmap_bf16(&p.tmA2, d->a2, 4, dims, str, box, 1)) return UWU_ERR_INVALID;
        } else {
            p.tmA2 = p.tmA;
        }
        p.H = H; p.W = W;
        p.kb_per_tap = (C1 + C2) / 64;
        p.kb_src1 = C1 / 64;
        p.num_kb = d->ntaps * p.kb_per_tap;
        for (int t = 0; t < 9; ++t) {
            p.tap_dn[t] = t < d->ntaps ? d->tap_dn[t] : 0;
            p.tap_dh[t] = t < d->ntaps ? d->tap_dh[t] : 0;
            p.tap_dw[t] = t < d->ntaps ? d->tap_dw[t] : 0;
        }
        p.a_mode = 2;
        p.a_lbo = 0; p.a_sbo = 1024; p.a_kadv = 32;
    }

    // ---------------- B tensor map ----------------
    if (d->b_layout == UWU_B_NK) {
        UWU_CHECK_ARG(d->ldb % 8 == 0 && d->ldb >= d->K, "uwu_gemm: ldb %lld must be >= K and a multiple of 8", (long long)d->ldb);
        uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)(d->grp_n > 0 ? d->grp_n : d->N)};
        uint64_t str[1] = {(uint64_t)d->ldb * 2};
        uint32_t box[2] = {BLOCK_K, (uint32_t)(bn > 256 ? 256 : bn)};  // (wide tiles never use this map: per-instruction half boxes)
        if (encode_tmap_bf16(&p.tmB, d->b, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
        p.b_stage_bytes = bn * BLOCK_K * 2;
        p.b_lbo = 0; p.b_sbo = 1024; p.b_kadv = 32;
    } else {
        UWU_CHECK_ARG(d->ldb % 8 == 0 && (d->grp_n > 0 || d->ldb >= d->N), "uwu_gemm: ldb %lld must be >= N and a multiple of 8",
                      (long long)d->ldb);
        const int segs = (d->a_layout == UWU_A_COL && d->k_segs > 1) ? d->k_segs : 1;
        uint64_t bw = d->grp_n > 0 ? (uint64_t)d->grp_n : (uint64_t)d->N + (uint64_t)(segs - 1) * (uint64_t)d->b_seg_off;
        UWU_CHECK_ARG((int64_t)bw <= d->ldb, "uwu_gemm: segmented / grouped B columns exceed ldb");
        p.b_3d = 0;
        if (bw % 64 == 0 && bn % 64 == 0 && (d->grp_n == 0 || d->grp_n % 64 == 0) && (segs == 1 || d->b_seg_off % 64 == 0)) {
            // (grouped / segmented operands too: the group / segment only shifts the 64-column coordinate)
            uint64_t dims[3] = {64, (uint64_t)d->K, (uint64_t)(bw / 64)};
            uint64_t str[2] = {(uint64_t)d->ldb * 2, 128};
            uint32_t box[3] = {64, BLOCK_K, (uint32_t)(bn / 64)};
            if (encode_tmap_bf16(&p.tmB, d->b, 3, dims, str, box, 1)) return UWU_ERR_INVALID;
            p.b_3d = 1;
        } else {
            uint64_t dims[2] = {bw, (uint64_t)d->K};
            uint64_t str[1] = {(uint64_t)d->ldb * 2};
            uint32_t box[2] = {64, BLOCK_K};
            if (encode_tmap_bf16(&p.tmB, d->b, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
        }
        p.b_stage_bytes = ((bn + 63) / 64) * 8192;
        p.b_lbo = 8192; p.b_sbo = 1024; p.b_kadv = 2048;
    }
    p.tx_bytes = A_STAGE_BYTES + p.b_stage_bytes;
    // ldb check for plain B_KN uses N, for grouped B the group width (done above)
    // B stage must keep the next A stage 1024-byte aligned
    p.b_stage_bytes = (p.b_stage_bytes + 1023) & ~1023;

    if (d->dbg_a_lbo) p.a_lbo = (uint32_t)d->dbg_a_lbo;
    if (d->dbg_a_sbo) p.a_sbo = (uint32_t)d->dbg_a_sbo;
    if (d->dbg_a_kadv) p.a_kadv = (uint32_t)d->dbg_a_kadv;
    if (d->dbg_b_lbo) p.b_lbo = (uint32_t)d->dbg_b_lbo;
    if (d->dbg_b_sbo) p.b_sbo = (uint32_t)d->dbg_b_sbo;
    if (d->dbg_b_kadv) p.b_kadv = (uint32_t)d->dbg_b_kadv;

    // ---------------- epilogue ----------------
    p.out = d->out;
    p.ldo = d->ldo > 0 ? d->ldo : d->N;
    p.out2 = d->out2;
    p.ldo2 = d->ldo2;
    p.n_split = (d->out2 && d->epi_mode == 0) ? d->n_split : 0x7fffffff;
    UWU_CHECK_ARG(d->out2 == nullptr || d->epi_mode != 0 || (d->n_split > 0 && d->n_split % 32 == 0 && d->n_split % bn == 0),
                  "uwu_gemm: n_split %d must be a positive multiple of block_n %d", d->n_split, bn);
    p.out_fp32 = d->out_dtype == UWU_F32;
    p.bias = d->bias;
    p.bias_rows = d->bias_rows;
    p.rows_per_bias = d->rows_per_bias > 0 ? d->rows_per_bias : 1;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
    p.ldr = d->ldr > 0 ? d->ldr : d->N;
    p.alpha = d->alpha;
    p.accumulate = d->accumulate;
    // bf16 outputs whose rows are 16-byte aligned go through the staged TMA-store epilogue
    p.epi_tma = 0;
    {
        const bool res_ok = d->residual == nullptr || (p.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0);
        const bool o2_ok = d->out2 == nullptr || (p.ldo2 % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out2) & 15) == 0 && d->n_split % 64 == 0);
        if (d->epi_mode == 0 && !p.out_fp32 && p.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 && p.N % 8 == 0 && res_ok &&
            o2_ok) {
            const long long n1 = d->out2 ? (long long)d->n_split : (long long)p.N;
            uint64_t dims[2] = {(uint64_t)n1, (uint64_t)p.M};
            uint64_t str[1] = {(uint64_t)p.ldo * 2};
            uint32_t box[2] = {64, 32};
            if (encode_tmap_bf16(&p.tmO, d->out, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
            p.tmO2 = p.tmO;
            if (d->out2) {
                uint64_t dims2[2] = {(uint64_t)(p.N - d->n_split), (uint64_t)p.M};
                uint64_t str2[1] = {(uint64_t)p.ldo2 * 2};
                if (encode_tmap_bf16(&p.tmO2, d->out2, 2, dims2, str2, box, 1)) return UWU_ERR_INVALID;
            }
            p.epi_tma = 1;
        }
    }
    // short reductions are bound by the epilogue, and the epilogue by the latency of the residual rows: fetch them by TMA a
    // tile ahead.  It costs 64 KB of the stage ring (6 -> 4 stages): measured (profiles/r02_gemm_res_tma.log) K = 1280: 60.6 -> 52.8
    // us at N = 1280, K = 640: 100 -> 76 us, but K = 2560 loses 9 % and K = 3840 3 %, hence the bound on K
    p.res_tma = 0;
    {
        static int want = -1;
        if (want < 0) {
            const char* e = getenv("UWU_GEMM_RES_TMA");
            want = e ? atoi(e) : 1;
        }
        if (want && p.epi_tma && d->epi_mode == 0 && d->residual != nullptr && bn <= 256 && n_inst == 1 && d->K <= (want > 1 ? (1 << 30) : 1536)) {
            uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M};
            uint64_t str[1] = {(uint64_t)p.ldr * 2};
            uint32_t box[2] = {64, 32};
            if (encode_tmap_bf16(&p.tmAux, d->residual, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
            p.res_tma = 1;
            p.epi_warp_bytes = 8192 + 4 * 4096;
        }
    }
    if (d->epi_mode != 0) {
        // fused GEGLU: `out` is [M, 2F] in both modes (pre-activation / its gradient), forward adds `out2` [M, F] = h * gelu(g),
        // backward reads the saved pre-activation through tmAux
        const long long F2 = 2ll * p.geglu_f;
        UWU_CHECK_ARG(p.ldo % 8 == 0 && p.ldo >= F2 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
                      "uwu_gemm(GEGLU): out must be a 16-byte aligned [M, 2F] bf16 matrix");
        uint32_t box[2] = {64, 32};
        {
            uint64_t dims[2] = {(uint64_t)F2, (uint64_t)p.M};
            uint64_t str[1] = {(uint64_t)p.ldo * 2};
            if (encode_tmap_bf16(&p.tmO, d->out, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
            p.tmO2 = p.tmO;
        }
        if (d->epi_mode == 1) {
            UWU_CHECK_ARG(p.ldo2 % 8 == 0 && p.ldo2 >= p.geglu_f && (reinterpret_cast<uintptr_t>(d->out2) & 15) == 0,
                          "uwu_gemm(GEGLU fwd): out2 must be a 16-byte aligned [M, F] bf16 matrix");
            uint64_t dims[2] = {(uint64_t)p.geglu_f, (uint64_t)p.M};
            uint64_t str[1] = {(uint64_t)p.ldo2 * 2};
            if (encode_tmap_bf16(&p.tmO2, d->out2, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
        } else {
            UWU_CHECK_ARG(d->ld_aux >= F2 && (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0, "uwu_gemm(GEGLU bwd): bad aux");
            uint64_t dims[2] = {(uint64_t)F2, (uint64_t)p.M};
            uint64_t str[1] = {(uint64_t)d->ld_aux * 2};
            if (encode_tmap_bf16(&p.tmAux, d->aux, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
        }
        p.epi_tma = 1;
    }
    // CTA-pair eligibility (decided before the stream-K schedule: the schedule's workers are then clusters, its tiles pairs)
    static int want_pair = -1, want_mc = -1;
    if (want_pair < 0) {
        const char* e = getenv("UWU_GEMM_PAIR");
        want_pair = e ? atoi(e) : 2;  // 2: everywhere it fits; 1: not for the split-K (weight-gradient) schedules; 0: never
        const char* e2 = getenv("UWU_GEMM_CLUSTER");
        want_mc = e2 ? atoi(e2) : 0;  // multicast-only clusters: measured neutral in round 1, kept for comparison
    }
    const bool half_ok = d->b_layout == UWU_B_NK ? (bn % 16 == 0) : (p.b_3d && bn % 128 == 0);
    const bool pair_shape_ok = p.num_m_tiles >= 2 && half_ok && sm_count() % 2 == 0;  // (k_segs / grp_n: KN operands, need b_3d)
    const bool sk_candidate = d->stream_k != 0 && p.out_fp32 && !d->bias && !d->bias_rows && !d->residual && p.num_kb >= 32;
    // split-K (weight-gradient) schedules: pairs only when the row tiles pair up evenly (measured: G[640,*] and G[1920,*], 5 / 15
    // row tiles, lose 8 % to the half-empty last pair; G[1280,*] .. G[5120,*] gain 4 - 14 %)
    const bool use_pair = want_pair && pair_shape_ok && (!sk_candidate || (want_pair >= 2 && p.num_m_tiles % 2 == 0));

    UWU_CHECK_ARG(d->epi_mode == 0 || use_pair,
                  "uwu_gemm(GEGLU): the fused epilogues need the CTA-pair kernel (M >= 256, UWU_GEMM_PAIR != 0)");
    // stream-K: few output tiles but a long reduction (token-reduction weight gradients) would leave most SMs idle
    const int n_tiles_all = use_pair ? ((p.num_m_tiles + 1) / 2) * p.num_n_tiles : p.num_m_tiles * p.num_n_tiles;
    int sk = d->stream_k;
    p.sk_slices = 1;
    if (sk < 0 || sk == 2) {
        const int sms = use_pair ? sm_count() / 2 : sm_count();
        const bool ok = p.out_fp32 && !d->bias && !d->bias_rows && !d->residual && p.num_kb >= 32;
        const int waves = (n_tiles_all + sms - 1) / sms;
        const double operand_bytes = (double)p.num_kb * 64.0 * ((double)p.M + (double)p.N) * 2.0;
        if (!ok) {
            sk = 0;
        } else if (sk < 0 && operand_bytes <= 96e6) {
            // both operands stay L2-resident whatever the schedule: perfectly balanced contiguous ranges
            sk = ((double)n_tiles_all < 0.85 * (double)waves * sms) ? 1 : 0;
        } else {
            // S slices so that S * tiles fills whole waves; a unit costs max(its k-blocks, ~24 k-blocks of atomic epilogue)
            const double t1 = (double)waves * p.num_kb;
            double best = t1;
            int best_s = 1;
            int max_s = p.num_kb / 8;
            if (max_s > 64) max_s = 64;
            for (int S = 2; S <= max_s; ++S) {
                const long long units = (long long)n_tiles_all * S;
                const double per = (double)p.num_kb / S;
                const double t = (double)((units + sms - 1) / sms) * (per > 24.0 ? per : 24.0);
                if (t < best * 0.98) {
                    best = t;
                    best_s = S;
                }
            }
            if (best_s > 1 && best <= 0.8 * t1) {
                sk = 2;
                p.sk_slices = best_s;
            } else {
                sk = 0;
            }
        }
    }
    if (sk) {
        UWU_CHECK_ARG(p.out_fp32 && !d->bias && !d->bias_rows && !d->residual,
                      "uwu_gemm: stream_k needs an fp32 output and a plain (alpha-only) epilogue");
        if (!d->accumulate) {
            // atomics add onto the destination: give "out = A.B" semantics by clearing it first
            UWU_CHECK_CUDA(cudaMemset2DAsync(p.out, (size_t)p.ldo * 4, 0, (size_t)min((long long)p.N, (long long)p.n_split) * 4,
                                             (size_t)p.M, stream));
            if (d->out2)
                UWU_CHECK_CUDA(cudaMemset2DAsync(p.out2, (size_t)p.ldo2 * 4, 0, (size_t)(p.N - p.n_split) * 4, (size_t)p.M, stream));
        }
    }
    p.stream_k = sk;

    // ---------------- CTA pair (cta_group::2) ----------------
    // One 256 x block_n tile per cluster of two CTAs: every CTA stages its own 128 rows of A and HALF of the B tile, the
    // leader issues M = 256 UMMAs.  Per FLOP that is a third less shared-memory fill traffic and half the B operand reads of
    // the 1-CTA kernel (128 x 256 tiles).  UWU_GEMM_PAIR=0 falls back to the 1-CTA kernel everywhere.
    p.pair = 0;
    p.cluster2 = 0;
    p.b_half_bytes = 0;
    p.tmBh = p.tmB;
    p.n_inst = 1;
    p.bn_inst = bn;
    p.b_inst_bytes = 0;
    p.n_acc = 2;
    UWU_CHECK_ARG(n_inst == 1 || use_pair, "uwu_gemm: internal error (wide tile without the CTA-pair kernel)");
    if ((use_pair || (want_mc && pair_shape_ok && !p.stream_k)) ) {
        if (d->b_layout == UWU_B_NK) {
            uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)(d->grp_n > 0 ? d->grp_n : d->N)};
            uint64_t str[1] = {(uint64_t)d->ldb * 2};
            p.n_inst = n_inst;
            p.bn_inst = bn / n_inst;
            p.b_inst_bytes = (p.bn_inst / 2) * BLOCK_K * 2;
            p.n_acc = bn > 256 ? 1 : 2;
            uint32_t box[2] = {BLOCK_K, (uint32_t)(p.bn_inst / 2)};
            if (encode_tmap_bf16(&p.tmBh, d->b, 2, dims, str, box, 1)) return UWU_ERR_INVALID;
            p.b_half_bytes = n_inst * p.b_inst_bytes;
        } else {
            const int segs2 = (d->a_layout == UWU_A_COL && d->k_segs > 1) ? d->k_segs : 1;
            const uint64_t bw2 = d->grp_n > 0 ? (uint64_t)d->grp_n : (uint64_t)d->N + (uint64_t)(segs2 - 1) * (uint64_t)d->b_seg_off;
            uint64_t dims[3] = {64, (uint64_t)d->K, (uint64_t)(bw2 / 64)};
            uint64_t str[2] = {(uint64_t)d->ldb * 2, 128};
            uint32_t box[3] = {64, BLOCK_K, (uint32_t)(bn / 128)};
            if (encode_tmap_bf16(&p.tmBh, d->b, 3, dims, str, box, 1)) return UWU_ERR_INVALID;
            p.b_half_bytes = (bn / 128) * 8192;
        }
        if (use_pair) {
            p.pair = 1;
            p.b_stage_bytes = p.b_half_bytes;                       // each CTA holds only its half
            p.tx_bytes = 2 * (A_STAGE_BYTES + p.b_half_bytes);      // both CTAs' boxes complete on the leader's barrier
        } else if (num_tiles_ge(p, 2 * sm_count())) {
            p.cluster2 = 1;
        }
    }

    // ---------------- launch ----------------
    const int stage_bytes = A_STAGE_BYTES + p.b_stage_bytes;
    const int epi_bytes = 4 * p.epi_warp_bytes;
    const int smem_budget = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - epi_bytes;
    int stages = smem_budget / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    UWU_CHECK_ARG(stages >= 2, "uwu_gemm: tile too large for shared memory");
    p.stages = stages;
    const size_t smem_bytes = (size_t)stages * stage_bytes + epi_bytes + 1024 + 512;

    static bool attr_set = false;
    if (!attr_set) {
        UWU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        UWU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    // workers = CTAs, or clusters of two CTAs walking (m-pair, n-tile) units
    const bool clustered = p.pair || p.cluster2;
    const long long work_tiles = clustered ? (long long)((p.num_m_tiles + 1) / 2) * p.num_n_tiles
                                           : (long long)p.num_m_tiles * p.num_n_tiles;
    long long workers = clustered ? sm_count() / 2 : sm_count();
    if (!p.stream_k && workers > work_tiles) workers = work_tiles;
    if (p.stream_k == 1 && workers > work_tiles * p.num_kb) workers = work_tiles * p.num_kb;
    if (p.stream_k == 2 && workers > work_tiles * p.sk_slices) workers = work_tiles * p.sk_slices;
    const int grid = (int)(clustered ? 2 * workers : workers);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.pair || p.cluster2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    static int pair_pdl = -1;
    if (pair_pdl < 0) {
        const char* e = getenv("UWU_GEMM_PAIR_PDL");
        pair_pdl = e ? atoi(e) : 1;
    }
    if (!(p.pair || p.cluster2) || pair_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (p.pair)
        UWU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<true>, p));
    else
        UWU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<false>, p));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return UWU_OK;
}
