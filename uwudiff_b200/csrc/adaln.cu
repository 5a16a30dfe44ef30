// adaLN-Zero glue of the DiT blocks (BASELINE.json configs[3]; north_star (c) "adaLN modulate"): bandwidth-bound kernels
// over token-major bf16 activations x[M, C], M = B * T, with per-sample fp32 modulation rows mod[B, ld_mod]
// (the output of `SiLU -> Linear(D, 6D)`; shift / scale / gate are column windows of it, addressed by offset).
//
// The reference tree has no DiT model; the block algebra follows the `ada_norm_zero` branch of its patched diffusers block
// (/root/reference/src/duwu/modules/rope_unet.py:306-309 modulate, :344-345 gate_msa, :395-398 second modulate, :406-407
// gate_mlp) and the public DiT definition (restated in oracle/dit_oracle.py).
//
//   adaln_fwd:          y = LN(x) * (1 + scale[b]) + shift[b]            LN without affine, stats = {mean, rstd} per row
//   adaln_bwd:          dx = LN'(dy * (1 + scale[b])) (+ dres);  dshift[b] = sum_t dy;  dscale[b] = sum_t dy * xhat
//   gate_residual_fwd:  out = x + gate[b] * y
//   gate_residual_bwd:  dy = gate[b] * dout;  dgate[b] = sum_t dout * y   (dx = dout, the residual stream itself)
//   patchify / unpatchify: NCHW fp32 images <-> [B*T, ld] token rows (patch-embed and final-layer layouts)
//
// One warp per row, the row held in registers (one HBM read per operand, one write); per-sample reductions are
// accumulated in registers over the block's rows, reduced through shared memory and stored once (deterministic).
#include "api_internal.h"
#include "common.cuh"

namespace uwu {
namespace {

constexpr int AD_WARPS = 8;

UWU_DEVINL void unpack8(const uint4& u, float (&v)[8]) {
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
UWU_DEVINL uint4 pack8(const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    return u;
}
UWU_DEVINL void ldf8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <int kMaxV>
__global__ void __launch_bounds__(AD_WARPS * 32) adaln_fwd_kernel(const __nv_bfloat16* __restrict__ x, int M, int C, float eps,
                                                                  const float* __restrict__ mod, long long ld_mod, int shift_off,
                                                                  int scale_off, int rows_per_mod,
                                                                  __nv_bfloat16* __restrict__ y, float* __restrict__ stats) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cv = C / 8;
    const float inv_c = 1.0f / (float)C;
    for (int row = blockIdx.x * AD_WARPS + warp; row < M; row += gridDim.x * AD_WARPS) {
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
        if (lane == 0) {
            stats[(size_t)row * 2] = mean;
            stats[(size_t)row * 2 + 1] = rstd;
        }
        const float* mrow = mod + (size_t)(row / rows_per_mod) * ld_mod;
        __nv_bfloat16* yr = y + (size_t)row * C;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            if (v < cv) {
                float f[8], sc[8], sh[8];
                unpack8(raw[k], f);
                ldf8(mrow + scale_off + v * 8, sc);
                ldf8(mrow + shift_off + v * 8, sh);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf((f[j] - mean) * rstd, 1.0f + sc[j], sh[j]);
                *reinterpret_cast<uint4*>(yr + v * 8) = pack8(f);
            }
        }
    }
}

// One block per sample: its AD_WARPS warps stream the sample's rows; dshift / dscale partials live in registers.
template <int kMaxV>
__global__ void __launch_bounds__(AD_WARPS * 32) adaln_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                  const __nv_bfloat16* __restrict__ dy, int C,
                                                                  const float* __restrict__ mod, long long ld_mod, int scale_off,
                                                                  const float* __restrict__ stats, int rows_per_mod,
                                                                  const __nv_bfloat16* __restrict__ dres,
                                                                  __nv_bfloat16* __restrict__ dx,
                                                                  __nv_bfloat16* __restrict__ dmod, long long ld_dmod,
                                                                  int dshift_off, int dscale_off) {
    extern __shared__ float red[];  // [AD_WARPS][2][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cv = C / 8;
    const float inv_c = 1.0f / (float)C;
    const int b = blockIdx.x;
    const float* mrow = mod + (size_t)b * ld_mod + scale_off;
    float g1[kMaxV][8];   // 1 + scale
    float a_sh[kMaxV][8], a_sc[kMaxV][8];
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
        const int v = lane + k * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            g1[k][j] = 1.0f;
            a_sh[k][j] = 0.f;
            a_sc[k][j] = 0.f;
        }
        if (v < cv) {
            float sc[8];
            ldf8(mrow + v * 8, sc);
#pragma unroll
            for (int j = 0; j < 8; ++j) g1[k][j] = 1.0f + sc[j];
        }
    }
    for (int r = warp; r < rows_per_mod; r += AD_WARPS) {
        const size_t row = (size_t)b * rows_per_mod + r;
        const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
        uint4 rx[kMaxV], rd[kMaxV];
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            rx[k] = v < cv ? *reinterpret_cast<const uint4*>(x + row * C + v * 8) : make_uint4(0, 0, 0, 0);
            rd[k] = v < cv ? *reinterpret_cast<const uint4*>(dy + row * C + v * 8) : make_uint4(0, 0, 0, 0);
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            if (lane + k * 32 < cv) {
                float fx[8], fd[8];
                unpack8(rx[k], fx);
                unpack8(rd[k], fd);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (fx[j] - mean) * rstd;
                    const float g = fd[j] * g1[k][j];
                    s1 += g;
                    s2 = fmaf(g, xh, s2);
                    a_sh[k][j] += fd[j];
                    a_sc[k][j] = fmaf(fd[j], xh, a_sc[k][j]);
                }
            }
        }
        const float m1 = warp_sum(s1) * inv_c, m2 = warp_sum(s2) * inv_c;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            if (v < cv) {
                float fx[8], fd[8], o[8];
                unpack8(rx[k], fx);
                unpack8(rd[k], fd);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (fx[j] - mean) * rstd;
                    o[j] = rstd * (fd[j] * g1[k][j] - m1 - xh * m2);
                }
                if (dres) {
                    float fr[8];
                    unpack8(*reinterpret_cast<const uint4*>(dres + row * C + v * 8), fr);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] += fr[j];
                }
                *reinterpret_cast<uint4*>(dx + row * C + v * 8) = pack8(o);
            }
        }
    }
    // cross-warp reduction of the per-sample sums
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
        const int v = lane + k * 32;
        if (v < cv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                red[(warp * 2 + 0) * C + v * 8 + j] = a_sh[k][j];
                red[(warp * 2 + 1) * C + v * 8 + j] = a_sc[k][j];
            }
        }
    }
    __syncthreads();
    __nv_bfloat16* drow = dmod + (size_t)b * ld_dmod;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sh = 0.f, sc = 0.f;
#pragma unroll
        for (int w = 0; w < AD_WARPS; ++w) {
            sh += red[(w * 2 + 0) * C + c];
            sc += red[(w * 2 + 1) * C + c];
        }
        drow[dshift_off + c] = __float2bfloat16(sh);
        drow[dscale_off + c] = __float2bfloat16(sc);
    }
}

__global__ void __launch_bounds__(256) gate_residual_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ y, long long M, int C,
                                                                const float* __restrict__ mod, long long ld_mod, int gate_off,
                                                                int rows_per_mod, __nv_bfloat16* __restrict__ out) {
    const int cv = C / 8;
    const long long total = M * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cv;
        const int v = (int)(i - row * cv);
        float fx[8], fy[8], g[8];
        unpack8(*reinterpret_cast<const uint4*>(x + row * C + v * 8), fx);
        unpack8(*reinterpret_cast<const uint4*>(y + row * C + v * 8), fy);
        ldf8(mod + (row / rows_per_mod) * ld_mod + gate_off + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) fx[j] = fmaf(g[j], fy[j], fx[j]);
        *reinterpret_cast<uint4*>(out + row * C + v * 8) = pack8(fx);
    }
}

template <int kMaxV>
__global__ void __launch_bounds__(AD_WARPS * 32) gate_residual_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                                          const __nv_bfloat16* __restrict__ y, int C,
                                                                          const float* __restrict__ mod, long long ld_mod,
                                                                          int gate_off, int rows_per_mod,
                                                                          __nv_bfloat16* __restrict__ dy,
                                                                          __nv_bfloat16* __restrict__ dmod, long long ld_dmod,
                                                                          int dgate_off) {
    extern __shared__ float red[];  // [AD_WARPS][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cv = C / 8;
    const int b = blockIdx.x;
    const float* mrow = mod + (size_t)b * ld_mod + gate_off;
    float g[kMaxV][8], acc[kMaxV][8];
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
        const int v = lane + k * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            g[k][j] = 0.f;
            acc[k][j] = 0.f;
        }
        if (v < cv) ldf8(mrow + v * 8, g[k]);
    }
    for (int r = warp; r < rows_per_mod; r += AD_WARPS) {
        const size_t row = (size_t)b * rows_per_mod + r;
#pragma unroll
        for (int k = 0; k < kMaxV; ++k) {
            const int v = lane + k * 32;
            if (v < cv) {
                float fd[8], fy[8], o[8];
                unpack8(*reinterpret_cast<const uint4*>(dout + row * C + v * 8), fd);
                unpack8(*reinterpret_cast<const uint4*>(y + row * C + v * 8), fy);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = fd[j] * g[k][j];
                    acc[k][j] = fmaf(fd[j], fy[j], acc[k][j]);
                }
                *reinterpret_cast<uint4*>(dy + row * C + v * 8) = pack8(o);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
        const int v = lane + k * 32;
        if (v < cv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) red[warp * C + v * 8 + j] = acc[k][j];
        }
    }
    __syncthreads();
    __nv_bfloat16* drow = dmod + (size_t)b * ld_dmod + dgate_off;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < AD_WARPS; ++w) s += red[w * C + c];
        drow[c] = __float2bfloat16(s);
    }
}

// order 0: column = c * p*p + ph * p + pw  (Conv2d(C, D, p, stride p) weight flattening: patch embedding)
// order 1: column = (ph * p + pw) * Ctok + c  (DiT unpatchify: tokens carry [p, p, Cout])
UWU_DEVINL int patch_col(int order, int c, int ph, int pw, int p, int Ctok) {
    return order == 0 ? c * p * p + ph * p + pw : (ph * p + pw) * Ctok + c;
}

// img[B, Cimg, H, W] fp32 -> tok[B*T, ld] bf16; columns of channels >= Cimg and columns >= Ctok*p*p are zero
__global__ void patchify_kernel(const float* __restrict__ img, int B, int Cimg, int H, int W, int p, int order, int Ctok,
                                __nv_bfloat16* __restrict__ tok, long long ld) {
    const int hp = H / p, wp = W / p;
    const long long total = (long long)B * hp * wp * ld;
    const int ncol = Ctok * p * p;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % ld);
        const long long t = i / ld;
        float v = 0.f;
        if (col < ncol) {
            int c, ph, pw;
            if (order == 0) {
                c = col / (p * p);
                ph = (col / p) % p;
                pw = col % p;
            } else {
                c = col % Ctok;
                pw = (col / Ctok) % p;
                ph = col / (Ctok * p);
            }
            if (c < Cimg) {
                const int tw = (int)(t % wp), th = (int)((t / wp) % hp);
                const long long b = t / ((long long)wp * hp);
                v = img[((b * Cimg + c) * H + th * p + ph) * W + tw * p + pw];
            }
        }
        tok[i] = __float2bfloat16(v);
    }
}

// tok[B*T, ld] (bf16 or fp32) -> img[B, Cimg, H, W] fp32 taking the first Cimg of Ctok channels
template <typename T>
__global__ void unpatchify_kernel(const T* __restrict__ tok, long long ld, int B, int Cimg, int H, int W, int p, int order,
                                  int Ctok, float* __restrict__ img) {
    const int hp = H / p, wp = W / p;
    const long long total = (long long)B * Cimg * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H);
        const int c = (int)((i / ((long long)W * H)) % Cimg);
        const long long b = i / ((long long)W * H * Cimg);
        const long long t = (b * hp + h / p) * wp + w / p;
        img[i] = (float)tok[t * ld + patch_col(order, c, h % p, w % p, p, Ctok)];
    }
}

// out[b, :] = table[idx[b], :]  (fp32 table -> bf16 rows)
__global__ void embed_gather_kernel(const float* __restrict__ table, const long long* __restrict__ idx, int B, int D, int V,
                                    __nv_bfloat16* __restrict__ out) {
    const long long total = (long long)B * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / D), c = (int)(i - (long long)b * D);
        long long r = idx[b];
        r = r < 0 ? 0 : (r >= V ? V - 1 : r);
        out[i] = __float2bfloat16(table[r * D + c]);
    }
}
// dtable[idx[b], :] += dout[b, :]
__global__ void embed_scatter_kernel(const __nv_bfloat16* __restrict__ dout, const long long* __restrict__ idx, int B, int D,
                                     int V, float* __restrict__ dtable) {
    const long long total = (long long)B * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / D), c = (int)(i - (long long)b * D);
        long long r = idx[b];
        r = r < 0 ? 0 : (r >= V ? V - 1 : r);
        atomicAdd(dtable + r * D + c, __bfloat162float(dout[i]));
    }
}

int grid_for(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace
}  // namespace uwu

using namespace uwu;
typedef __nv_bfloat16 bf16;

#define UWU_AD_DISPATCH(cv, CALL)        \
    do {                                  \
        if ((cv) <= 32) { CALL(1); }      \
        else if ((cv) <= 64) { CALL(2); } \
        else if ((cv) <= 96) { CALL(3); } \
        else if ((cv) <= 160) { CALL(5); }\
        else { CALL(8); }                 \
    } while (0)

static int check_mod(const char* who, int64_t M, int32_t C, const void* mod, int64_t ld_mod, int32_t rows_per_mod) {
    UWU_CHECK_ARG(M > 0 && C > 0 && C % 8 == 0 && C <= 2048, "%s: bad shape M=%lld C=%d (C multiple of 8, <= 2048)", who,
                  (long long)M, C);
    UWU_CHECK_ARG(rows_per_mod > 0 && M % rows_per_mod == 0, "%s: M=%lld is not a multiple of rows_per_mod=%d", who, (long long)M,
                  rows_per_mod);
    UWU_CHECK_ARG(mod && ld_mod % 4 == 0 && (reinterpret_cast<uintptr_t>(mod) & 15) == 0, "%s: modulation rows must be 16-byte aligned",
                  who);
    return UWU_OK;
}

extern "C" int uwu_adaln_fwd(const void* x, int64_t M, int32_t C, float eps, const float* mod, int64_t ld_mod, int32_t shift_off,
                             int32_t scale_off, int32_t rows_per_mod, void* y, float* stats, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = check_mod("uwu_adaln_fwd", M, C, mod, ld_mod, rows_per_mod)) return rc;
    UWU_CHECK_ARG(x && y && stats && shift_off % 4 == 0 && scale_off % 4 == 0, "uwu_adaln_fwd: bad pointer / offset");
    int blocks = (int)((M + AD_WARPS - 1) / AD_WARPS);
    const int per_sm = C / 8 <= 32 ? 8 : C / 8 <= 96 ? 4 : C / 8 <= 160 ? 3 : 2;  // resident blocks per SM at the variant's registers
    if (blocks > sm_count() * per_sm) blocks = sm_count() * per_sm;
#define CALL(V) adaln_fwd_kernel<V><<<blocks, AD_WARPS * 32, 0, stream>>>(reinterpret_cast<const bf16*>(x), (int)M, C, eps, mod, ld_mod, shift_off, scale_off, rows_per_mod, reinterpret_cast<bf16*>(y), stats)
    UWU_AD_DISPATCH(C / 8, CALL);
#undef CALL
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_adaln_bwd(const void* x, const void* dy, int64_t M, int32_t C, const float* mod, int64_t ld_mod,
                             int32_t scale_off, const float* stats, int32_t rows_per_mod, const void* dres, void* dx,
                             void* dmod_bf16, int64_t ld_dmod, int32_t dshift_off, int32_t dscale_off, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = check_mod("uwu_adaln_bwd", M, C, mod, ld_mod, rows_per_mod)) return rc;
    UWU_CHECK_ARG(x && dy && dx && stats && dmod_bf16 && scale_off % 4 == 0, "uwu_adaln_bwd: bad pointer / offset");
    const int B = (int)(M / rows_per_mod);
    const size_t smem = (size_t)AD_WARPS * 2 * C * sizeof(float);
#define CALL(V)                                                                                                             \
    do {                                                                                                                    \
        static bool attr = false;                                                                                           \
        if (!attr) {                                                                                                        \
            UWU_CHECK_CUDA(cudaFuncSetAttribute(adaln_bwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, AD_WARPS * 2 * 2048 * 4)); \
            attr = true;                                                                                                    \
        }                                                                                                                   \
        adaln_bwd_kernel<V><<<B, AD_WARPS * 32, smem, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), C, mod, \
            ld_mod, scale_off, stats, rows_per_mod, reinterpret_cast<const bf16*>(dres), reinterpret_cast<bf16*>(dx),      \
            reinterpret_cast<bf16*>(dmod_bf16), ld_dmod, dshift_off, dscale_off);                                           \
    } while (0)
    UWU_AD_DISPATCH(C / 8, CALL);
#undef CALL
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_gate_residual_fwd(const void* x, const void* y, int64_t M, int32_t C, const float* mod, int64_t ld_mod,
                                     int32_t gate_off, int32_t rows_per_mod, void* out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = check_mod("uwu_gate_residual_fwd", M, C, mod, ld_mod, rows_per_mod)) return rc;
    UWU_CHECK_ARG(x && y && out && gate_off % 4 == 0, "uwu_gate_residual_fwd: bad pointer / offset");
    gate_residual_fwd_kernel<<<grid_for(M * (C / 8), 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(y),
                                                                            M, C, mod, ld_mod, gate_off, rows_per_mod,
                                                                            reinterpret_cast<bf16*>(out));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_gate_residual_bwd(const void* dout, const void* y, int64_t M, int32_t C, const float* mod, int64_t ld_mod,
                                     int32_t gate_off, int32_t rows_per_mod, void* dy, void* dmod_bf16, int64_t ld_dmod,
                                     int32_t dgate_off, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = check_mod("uwu_gate_residual_bwd", M, C, mod, ld_mod, rows_per_mod)) return rc;
    UWU_CHECK_ARG(dout && y && dy && dmod_bf16 && gate_off % 4 == 0, "uwu_gate_residual_bwd: bad pointer / offset");
    const int B = (int)(M / rows_per_mod);
    const size_t smem = (size_t)AD_WARPS * C * sizeof(float);
#define CALL(V)                                                                                                              \
    do {                                                                                                                     \
        static bool attr = false;                                                                                            \
        if (!attr) {                                                                                                         \
            UWU_CHECK_CUDA(cudaFuncSetAttribute(gate_residual_bwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, AD_WARPS * 2048 * 4)); \
            attr = true;                                                                                                     \
        }                                                                                                                    \
        gate_residual_bwd_kernel<V><<<B, AD_WARPS * 32, smem, stream>>>(reinterpret_cast<const bf16*>(dout), reinterpret_cast<const bf16*>(y), C, \
            mod, ld_mod, gate_off, rows_per_mod, reinterpret_cast<bf16*>(dy), reinterpret_cast<bf16*>(dmod_bf16), ld_dmod, dgate_off); \
    } while (0)
    UWU_AD_DISPATCH(C / 8, CALL);
#undef CALL
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_patchify(const float* img, int32_t B, int32_t Cimg, int32_t H, int32_t W, int32_t p, int32_t order,
                            int32_t Ctok, void* tok_bf16, int64_t ld, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(B > 0 && Cimg > 0 && p > 0 && H % p == 0 && W % p == 0 && Ctok >= Cimg && ld >= (int64_t)Ctok * p * p &&
                      (order == 0 || order == 1),
                  "uwu_patchify: bad shape");
    UWU_CHECK_ARG(img && tok_bf16, "uwu_patchify: null pointer");
    const long long total = (long long)B * (H / p) * (W / p) * ld;
    patchify_kernel<<<grid_for(total, 256), 256, 0, stream>>>(img, B, Cimg, H, W, p, order, Ctok, reinterpret_cast<bf16*>(tok_bf16), ld);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_unpatchify(const void* tok, int32_t tok_dtype, int64_t ld, int32_t B, int32_t Cimg, int32_t H, int32_t W,
                              int32_t p, int32_t order, int32_t Ctok, float* img, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(B > 0 && Cimg > 0 && p > 0 && H % p == 0 && W % p == 0 && Ctok >= Cimg && ld >= (int64_t)Ctok * p * p &&
                      (order == 0 || order == 1),
                  "uwu_unpatchify: bad shape");
    UWU_CHECK_ARG(tok && img, "uwu_unpatchify: null pointer");
    const long long total = (long long)B * Cimg * H * W;
    if (tok_dtype == UWU_F32)
        unpatchify_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(tok), ld, B, Cimg, H, W, p, order, Ctok, img);
    else
        unpatchify_kernel<bf16><<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(tok), ld, B, Cimg, H, W, p, order, Ctok, img);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_embed_gather(const float* table, const int64_t* idx, int32_t B, int32_t D, int32_t V, void* out_bf16,
                                void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(table && idx && out_bf16 && B > 0 && D > 0 && V > 0, "uwu_embed_gather: bad arguments");
    embed_gather_kernel<<<grid_for((long long)B * D, 256), 256, 0, stream>>>(table, reinterpret_cast<const long long*>(idx), B, D, V,
                                                                            reinterpret_cast<bf16*>(out_bf16));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_embed_scatter_add(const void* dout_bf16, const int64_t* idx, int32_t B, int32_t D, int32_t V, float* dtable,
                                     void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(dout_bf16 && idx && dtable && B > 0 && D > 0 && V > 0, "uwu_embed_scatter_add: bad arguments");
    embed_scatter_kernel<<<grid_for((long long)B * D, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dout_bf16),
                                                                             reinterpret_cast<const long long*>(idx), B, D, V, dtable);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
