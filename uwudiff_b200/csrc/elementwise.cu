// Elementwise / layout glue over channels-last bf16 activations: GEGLU fwd/bwd, SiLU, residual add,
// NCHW<->NHWC boundary conversion (the reference API is NCHW, src/duwu/data/base.py:13), nearest-2x upsample
// fwd/bwd, stride-2 phase split (space-to-depth) and its inverse.  All vectorised 16-byte accesses.
//
// Replaces ATen elementwise kernels under diffusers GEGLU / Upsample2D / Downsample2D / residual adds
// [third-party, restated in oracle/unet_oracle.py; GEGLU chunk order hidden, gate per SURVEY.md Appendix A.3].
#include "api_internal.h"
#include "common.cuh"

namespace uwu {

UWU_DEVINL void ld8e(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
UWU_DEVINL void up8e(const uint4& u, float (&v)[8]) {
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
UWU_DEVINL void st8e(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// out[m, f] = in[m, f] * gelu(in[m, F + f]).  Each block owns a contiguous range of rows and its threads walk the
// (row, 8-wide column vector) pairs of that range with incremental 32-bit index updates (no 64-bit division).
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const __nv_bfloat16* __restrict__ in, long long M, int F,
                                                        __nv_bfloat16* __restrict__ out) {
    pdl_trigger();
    const int fv = F >> 3;
    const long long rows_per = (M + gridDim.x - 1) / gridDim.x;
    const long long m0 = blockIdx.x * rows_per;
    const int nrows = (int)min(rows_per, M - m0);
    if (nrows <= 0) return;
    int r = threadIdx.x / fv, v = threadIdx.x - r * fv;
    const int dr = blockDim.x / fv, dv = blockDim.x - dr * fv;
    while (r < nrows) {
        const __nv_bfloat16* ip = in + (m0 + r) * 2 * F;
        float h[8], g[8];
        ld8e(ip + v * 8, h);
        ld8e(ip + F + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] *= gelu_fast_f(g[j]);
        st8e(out + (m0 + r) * F + v * 8, h);
        r += dr;
        v += dv;
        if (v >= fv) {
            v -= fv;
            ++r;
        }
    }
}
// din[m, f] = dout * gelu(g);  din[m, F+f] = dout * h * gelu'(g),  gelu'(g) = Phi(g) + g phi(g)
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ dout,
                                                        long long M, int F, __nv_bfloat16* __restrict__ din) {
    pdl_trigger();
    const int fv = F >> 3;
    const long long rows_per = (M + gridDim.x - 1) / gridDim.x;
    const long long m0 = blockIdx.x * rows_per;
    const int nrows = (int)min(rows_per, M - m0);
    if (nrows <= 0) return;
    int r = threadIdx.x / fv, v = threadIdx.x - r * fv;
    const int dr = blockDim.x / fv, dv = blockDim.x - dr * fv;
    while (r < nrows) {
        const __nv_bfloat16* ip = in + (m0 + r) * 2 * F;
        __nv_bfloat16* op = din + (m0 + r) * 2 * F;
        float h[8], g[8], d[8], dh[8], dg[8];
        ld8e(ip + v * 8, h);
        ld8e(ip + F + v * 8, g);
        ld8e(dout + (m0 + r) * F + v * 8, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float cdf, pdf;
            gelu_cdf_pdf(g[j], cdf, pdf);
            dh[j] = d[j] * g[j] * cdf;
            dg[j] = d[j] * h[j] * fmaf(g[j], pdf, cdf);
        }
        st8e(op + v * 8, dh);
        st8e(op + F + v * 8, dg);
        r += dr;
        v += dv;
        if (v >= fv) {
            v -= fv;
            ++r;
        }
    }
}

// tanh on the XU pipe (one MUFU.TANH, ~2^-11 relative error: below the bf16 rounding of the result)
UWU_DEVINL float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
UWU_DEVINL float gelu_tanh_f(float x) {
    const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
    return 0.5f * x * (1.0f + tanh_fast(u));
}
UWU_DEVINL float gelu_tanh_grad_f(float x) {
    const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
    const float t = tanh_fast(u);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * 0.7978845608028654f * fmaf(3.0f * 0.044715f * x, x, 1.0f);
}

// mode 0: y = silu(x); mode 1: y = x * silu'(a) (x = dy, a = pre-activation); mode 2: y = x + a; mode 3: y = x;
// mode 4: y = gelu_tanh(x); mode 5: y = x * gelu_tanh'(a)   (DiT MLP activation)
// mode 6: y = x * sigmoid(1.702 x) ("quick_gelu", CLIP-L text tower); mode 7: y = gelu_erf(x) (CLIP-bigG text tower)
__global__ void ew_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ a, long long nvec,
                          int mode, __nv_bfloat16* __restrict__ y) {
    pdl_trigger();
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool two = mode == 1 || mode == 2 || mode == 5;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * U) {
        uint4 rx[U], ra[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < nvec) {
                rx[u] = *reinterpret_cast<const uint4*>(x + i * 8);
                if (two) ra[u] = *reinterpret_cast<const uint4*>(a + i * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i >= nvec) break;
            float f[8], b[8];
            up8e(rx[u], f);
            if (two) up8e(ra[u], b);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (mode == 0) f[j] = silu_f(f[j]);
                else if (mode == 1) f[j] *= silu_grad_f(b[j]);
                else if (mode == 2) f[j] += b[j];
                else if (mode == 4) f[j] = gelu_tanh_f(f[j]);
                else if (mode == 5) f[j] *= gelu_tanh_grad_f(b[j]);
                else if (mode == 6) f[j] = __fdividef(f[j], 1.0f + __expf(-1.702f * f[j]));
                else if (mode == 7) f[j] = gelu_fast_f(f[j]);
            }
            st8e(y + i * 8, f);
        }
    }
}

// NCHW (fp32 or bf16) -> NHWC bf16 with zero channel padding to Cpad
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ src, int N, int C, int HW, int Cpad, __nv_bfloat16* __restrict__ dst) {
    const long long total = (long long)N * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / HW;
        const int p = (int)(i - n * HW);
        __nv_bfloat16* d = dst + i * Cpad;
        for (int c = 0; c < Cpad; ++c) {
            float v = 0.f;
            if (c < C) v = (float)src[(n * C + c) * HW + p];
            d[c] = __float2bfloat16(v);
        }
    }
}
// rows [N*HW, ld] (bf16 or fp32), first C columns -> NCHW fp32
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int N, int C, int HW, long long ld, float* __restrict__ dst) {
    const long long total = (long long)N * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / HW;
        const int p = (int)(i - n * HW);
        for (int c = 0; c < C; ++c) dst[(n * C + c) * HW + p] = (float)src[i * ld + c];
    }
}

// nearest 2x upsample: y[n, 2h+a, 2w+b, c] = x[n, h, w, c]
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                                  __nv_bfloat16* __restrict__ y) {
    const int cv = C / 8;
    const long long total = (long long)N * 4 * H * W * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        long long r = i / cv;
        const int ow = (int)(r % (2 * W)); r /= (2 * W);
        const int oh = (int)(r % (2 * H));
        const long long n = r / (2 * H);
        const uint4 u = *reinterpret_cast<const uint4*>(x + (((n * H + oh / 2) * W + ow / 2) * C + v * 8));
        *reinterpret_cast<uint4*>(y + (((n * 2 * H + oh) * 2 * W + ow) * (long long)C + v * 8)) = u;
    }
}
// backward: dx[n,h,w,c] = sum_{a,b} dy[n, 2h+a, 2w+b, c]
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int N, int H, int W, int C,
                                      __nv_bfloat16* __restrict__ dx) {
    const int cv = C / 8;
    const long long total = (long long)N * H * W * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        long long r = i / cv;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const long long n = r / H;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float f[8];
                ld8e(dy + (((n * 2 * H + 2 * h + a) * 2 * W + 2 * w + b) * (long long)C + v * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        st8e(dx + (((n * H + h) * W + w) * (long long)C + v * 8), acc);
    }
}

// space-to-depth by 2: planes[(py*2+px)*N + n, i, j, c] = x[n, 2i+py, 2j+px, c]   (inverse = true: scatter back)
__global__ void phase_split_kernel(const __nv_bfloat16* __restrict__ src, int N, int H, int W, int C, int inverse,
                                   __nv_bfloat16* __restrict__ dst) {
    const int cv = C / 8;
    const int H2 = H / 2, W2 = W / 2;
    const long long total = (long long)N * H * W * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        long long r = i / cv;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const long long n = r / H;
        const long long full = (((n * H + h) * W + w) * (long long)C + v * 8);
        const int p = (h & 1) * 2 + (w & 1);
        const long long pl = ((((long long)p * N + n) * H2 + h / 2) * W2 + w / 2) * (long long)C + v * 8;
        if (inverse)
            *reinterpret_cast<uint4*>(dst + full) = *reinterpret_cast<const uint4*>(src + pl);
        else
            *reinterpret_cast<uint4*>(dst + pl) = *reinterpret_cast<const uint4*>(src + full);
    }
}

// column sums of a bf16 matrix (bias gradients): partial[blockIdx.y][c] over a row chunk
// partial[blockIdx.y][c] = sum of rows [blockIdx.y * rows_per_block, ...) of column c.  Thread = one 8-wide column vector
// (16-byte loads, four rows in flight), 4 row lanes per block combined through shared memory.
// kVec = false: generic fallback (C % 8 != 0 or unaligned rows), one column per thread.
template <bool kVec>
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, long long ld,
                                                          int rows_per_block, float* __restrict__ partial) {
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(M, r0 + rows_per_block);
    if (!kVec) {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        if (c >= C) return;
        float a = 0.f;
        for (long long r = r0; r < r1; ++r) a += __bfloat162float(x[r * ld + c]);
        partial[(size_t)blockIdx.y * C + c] = a;
        return;
    }
    __shared__ float sh[4][64][8];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int v = blockIdx.x * 64 + tx;
    const bool active = v * 8 < C;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (active) {
        const __nv_bfloat16* col = x + v * 8;
        long long r = r0 + ty;
        for (; r + 12 < r1; r += 16) {
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(col + (r + 4 * k) * ld);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f[8];
                up8e(u[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        }
        for (; r < r1; r += 4) {
            float f[8];
            up8e(*reinterpret_cast<const uint4*>(col + r * ld), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[ty][tx][j] = acc[j];
    __syncthreads();
    if (ty == 0 && active) {
        float* o = partial + (size_t)blockIdx.y * C + v * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (sh[0][tx][j] + sh[1][tx][j]) + (sh[2][tx][j] + sh[3][tx][j]);
    }
}
// out[c] (+)= sum_p partial[p, c]: 32 columns x 8 partial lanes per block
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int P, int C, int accumulate,
                                                           float* __restrict__ out) {
    __shared__ float sh[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (c < C)
        for (int p = ty; p < P; p += 8) a += partial[(size_t)p * C + c];
    sh[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += sh[k][tx];
        out[c] = accumulate ? out[c] + s : s;
    }
}

// ------------------------------------------------------------------------------------------------
// full fine-tuning support: convolution weight gradients as  dW = dY^T im2col(X)  (token-reduction GEMM)
// ------------------------------------------------------------------------------------------------
// cols[m, t*C + c] = x[n, ho*stride + dy_t - 1, wo*stride + dx_t - 1, c]  (0 outside the image), m = (n, ho, wo), t = 3*ky + kx
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int stride,
                                                        int Ho, int Wo, __nv_bfloat16* __restrict__ cols) {
    // thread = one 8-channel vector of one output pixel, all 9 taps (one index decomposition per 9 x 16 bytes written)
    const int cv = C >> 3;
    const long long total = (long long)N * Ho * Wo * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / cv;
        const int v = (int)(i - m * cv);
        const int wo = (int)(m % Wo);
        const long long r = m / Wo;
        const int ho = (int)(r % Ho), n = (int)(r / Ho);
        __nv_bfloat16* dst = cols + (size_t)m * (size_t)(9 * C) + v * 8;
        const __nv_bfloat16* src = x + ((size_t)n * H * W) * C + v * 8;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int hi = ho * stride + t / 3 - 1, wi = wo * stride + t % 3 - 1;
            uint4 u = make_uint4(0, 0, 0, 0);
            if (hi >= 0 && hi < H && wi >= 0 && wi < W) u = *reinterpret_cast<const uint4*>(src + ((size_t)hi * W + wi) * C);
            *reinterpret_cast<uint4*>(dst + (size_t)t * C) = u;
        }
    }
}
// Conv2d master weight W[Co, Ci, kk] (fp32, kk = kh*kw taps) -> the two bf16 GEMM operands of the implicit-GEMM convolution,
// zero padding included (every element of both outputs is written):
//   fwd  [co_p, kk * ci_p] : fwd[co, t * ci_p + ci]  = W[co, ci, t]
//   dgrad[Ci,  kk * cod_p] : dgrad[ci, t * cod_p + co] = W[co, ci, kk - 1 - t]   (correlation with the flipped kernel)
__global__ void __launch_bounds__(256) conv_pack_kernel(const float* __restrict__ W, int Co, int Ci, int kk, int ci_p, int co_p, int cod_p,
                                                        __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgrad) {
    // block = a (32 co x 32 ci) tile of the padded weight, staged through shared memory so that the reads (ci*kk contiguous
    // floats per co) and both writes (32 contiguous ci per (co, t); 32 contiguous co per (ci, t)) are coalesced
    extern __shared__ float tile[];  // [32 co][32 ci][kk] (+1 padding per co row)
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int row = 32 * kk + 1;
    for (int i = threadIdx.x; i < 32 * 32 * kk; i += blockDim.x) {
        const int co = i / (32 * kk), r = i - co * 32 * kk;  // r = ci_local * kk + t, contiguous in W for a fixed co
        const int ci = ci0 + r / kk;
        float v = 0.f;
        if (co0 + co < Co && ci < Ci) v = W[((size_t)(co0 + co) * Ci + ci0) * kk + r];
        tile[co * row + r] = v;
    }
    __syncthreads();
    // forward operand: fwd[co, t * ci_p + ci]
    for (int i = threadIdx.x; i < 32 * kk * 32; i += blockDim.x) {
        const int ci = i & 31, t = (i >> 5) % kk, co = (i >> 5) / kk;
        if (co0 + co < co_p && ci0 + ci < ci_p)
            fwd[(size_t)(co0 + co) * kk * ci_p + (size_t)t * ci_p + ci0 + ci] = __float2bfloat16(tile[co * row + ci * kk + t]);
    }
    // data-gradient operand: dgrad[ci, t * cod_p + co] = W[co, ci, kk - 1 - t]
    if (dgrad) {
        for (int i = threadIdx.x; i < 32 * kk * 32; i += blockDim.x) {
            const int co = i & 31, t = (i >> 5) % kk, ci = (i >> 5) / kk;
            if (ci0 + ci < Ci && co0 + co < cod_p)
                dgrad[(size_t)(ci0 + ci) * kk * cod_p + (size_t)t * cod_p + co0 + co] =
                    __float2bfloat16(tile[co * row + ci * kk + (kk - 1 - t)]);
        }
    }
}
// wgrad[co, ci, ky, kx] (+)= G[co, (3*ky + kx) * Cp + ci]   (torch Conv2d weight layout <- packed GEMM layout), kk = kh*kw
__global__ void __launch_bounds__(256) conv_wgrad_unpack_kernel(const float* __restrict__ G, long long ldg, int Co, int Ci, int Cp, int kk,
                                                                int accumulate, float* __restrict__ wgrad) {
    // warp = 32 consecutive ci of one co: per tap a contiguous 128-byte read of G; the 32 x kk results are transposed through
    // shared memory (stride kk <= 9 is odd or 1: conflict-free) so that the kk*32 contiguous floats of wgrad are written coalesced
    __shared__ float sh[8][32 * 9];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci_tiles = (Ci + 31) / 32;
    const int total = Co * ci_tiles;
    for (int w = blockIdx.x * 8 + warp; w < total; w += gridDim.x * 8) {
        const int co = w / ci_tiles, ci0 = (w - co * ci_tiles) * 32;
        const int ci = ci0 + lane;
        const float* g = G + (size_t)co * ldg + ci;
        for (int t = 0; t < kk; ++t) sh[warp][lane * kk + t] = ci < Ci ? g[(size_t)t * Cp] : 0.f;
        __syncwarp();
        const int nvalid = min(32, Ci - ci0) * kk;
        float* wout = wgrad + ((size_t)co * Ci + ci0) * kk;
        for (int e = lane; e < nvalid; e += 32) wout[e] = accumulate ? wout[e] + sh[warp][e] : sh[warp][e];
        __syncwarp();
    }
}
// out[n, c] (+)= sum_{r < rows} x[n * rows + r, c]   (per-image column sums: time-embedding gradients), one block per (n, 64 columns)
__global__ void __launch_bounds__(256) colsum_groups_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int rows, int C,
                                                            int accumulate, float* __restrict__ out) {
    __shared__ float sh[4][64];
    const int n = blockIdx.y, c = blockIdx.x * 64 + (threadIdx.x & 63), rr = threadIdx.x >> 6;
    float a = 0.f;
    if (c < C)
        for (int r = rr; r < rows; r += 4) a += __bfloat162float(x[((size_t)n * rows + r) * ldx + c]);
    sh[rr][threadIdx.x & 63] = a;
    __syncthreads();
    if (rr == 0 && c < C) {
        const float s = (sh[0][threadIdx.x] + sh[1][threadIdx.x]) + (sh[2][threadIdx.x] + sh[3][threadIdx.x]);
        float* o = out + (size_t)n * C + c;
        *o = accumulate ? *o + s : s;
    }
}

static int ew_grid(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace uwu

using namespace uwu;
typedef __nv_bfloat16 bf16;

// Row softmax of a bf16 matrix, in place: x[r, :] = softmax(x[r, :]).  Used by the frozen VAE encoder's single-head attention
// (head_dim 512 — outside the flash kernels' range; S = Q K^T and P V run on the GEMM kernel).  One block per row, the row is
// read three times from L2 (max, sum, write).
__global__ void __launch_bounds__(256) softmax_rows_kernel(__nv_bfloat16* __restrict__ x, long long ld, int N) {
    __nv_bfloat16* row = x + (size_t)blockIdx.x * ld;
    __shared__ float red[8];
    __shared__ float bc;
    const int nv = N >> 3;
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        float f[8];
        ld8e(row + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red[0];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
        bc = m;
    }
    __syncthreads();
    mx = bc;
    float sum = 0.f;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        float f[8];
        ld8e(row + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += __expf(f[j] - mx);
    }
    sum = warp_sum(sum);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        bc = 1.0f / t;
    }
    __syncthreads();
    const float inv = bc;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        float f[8];
        ld8e(row + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __expf(f[j] - mx) * inv;
        st8e(row + v * 8, f);
    }
}

extern "C" int uwu_softmax_rows(void* x, int64_t M, int32_t N, int64_t ld, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(M >= 0 && N > 0 && N % 8 == 0 && ld >= N && ld % 8 == 0, "uwu_softmax_rows: bad shape M=%lld N=%d ld=%lld",
                  (long long)M, N, (long long)ld);
    if (M == 0) return UWU_OK;
    UWU_CHECK_ARG(x && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && M < (1ll << 31), "uwu_softmax_rows: bad pointer / too many rows");
    softmax_rows_kernel<<<(unsigned)M, 256, 0, stream>>>(reinterpret_cast<bf16*>(x), ld, N);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_geglu_fwd(const void* in, int64_t M, int32_t F, void* out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(M >= 0 && F > 0 && F % 8 == 0, "uwu_geglu_fwd: bad shape M=%lld F=%d", (long long)M, F);
    if (M == 0) return UWU_OK;
    UWU_CHECK_ARG(in && out, "uwu_geglu_fwd: null pointer");
    geglu_fwd_kernel<<<(unsigned)(M < 16ll * sm_count() ? M : 16ll * sm_count()), 256, 0, stream>>>(reinterpret_cast<const bf16*>(in), M, F,
                                                                    reinterpret_cast<bf16*>(out));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_geglu_bwd(const void* in, const void* dout, int64_t M, int32_t F, void* din, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(M >= 0 && F > 0 && F % 8 == 0, "uwu_geglu_bwd: bad shape");
    if (M == 0) return UWU_OK;
    UWU_CHECK_ARG(in && dout && din, "uwu_geglu_bwd: null pointer");
    geglu_bwd_kernel<<<(unsigned)(M < 16ll * sm_count() ? M : 16ll * sm_count()), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(in), reinterpret_cast<const bf16*>(dout), M, F, reinterpret_cast<bf16*>(din));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_elementwise(const void* x, const void* a, int64_t n, int32_t mode, void* y, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(n >= 0 && n % 8 == 0, "uwu_elementwise: n=%lld must be a multiple of 8", (long long)n);
    UWU_CHECK_ARG(mode >= 0 && mode <= 7, "uwu_elementwise: bad mode %d", mode);
    if (n == 0) return UWU_OK;
    UWU_CHECK_ARG(x && y && (mode == 0 || mode == 3 || mode == 4 || mode == 6 || mode == 7 || a), "uwu_elementwise: null pointer");
    ew_kernel<<<ew_grid(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(a),
                                                       n / 8, mode, reinterpret_cast<bf16*>(y));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_nchw_to_nhwc(const void* src, int32_t src_dtype, int32_t N, int32_t C, int32_t HW, int32_t Cpad,
                                void* dst_bf16, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(N >= 0 && C > 0 && HW > 0 && Cpad >= C, "uwu_nchw_to_nhwc: bad shape");
    if (N == 0) return UWU_OK;
    UWU_CHECK_ARG(src && dst_bf16, "uwu_nchw_to_nhwc: null pointer");
    const int grid = ew_grid((long long)N * HW, 256);
    if (src_dtype == UWU_F32)
        nchw_to_nhwc_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(src), N, C, HW, Cpad,
                                                             reinterpret_cast<bf16*>(dst_bf16));
    else
        nchw_to_nhwc_kernel<bf16><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(src), N, C, HW, Cpad,
                                                            reinterpret_cast<bf16*>(dst_bf16));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_nhwc_to_nchw(const void* src, int32_t src_dtype, int32_t N, int32_t C, int32_t HW, int64_t ld,
                                float* dst, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(N >= 0 && C > 0 && HW > 0 && ld >= C, "uwu_nhwc_to_nchw: bad shape");
    if (N == 0) return UWU_OK;
    UWU_CHECK_ARG(src && dst, "uwu_nhwc_to_nchw: null pointer");
    const int grid = ew_grid((long long)N * HW, 256);
    if (src_dtype == UWU_F32)
        nhwc_to_nchw_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(src), N, C, HW, ld, dst);
    else
        nhwc_to_nchw_kernel<bf16><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(src), N, C, HW, ld, dst);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_upsample2x(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t backward, void* y,
                              void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(N >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "uwu_upsample2x: bad shape");
    if (N == 0) return UWU_OK;
    UWU_CHECK_ARG(x && y, "uwu_upsample2x: null pointer");
    if (!backward)
        upsample2x_kernel<<<ew_grid((long long)N * 4 * H * W * (C / 8), 256), 256, 0, stream>>>(
            reinterpret_cast<const bf16*>(x), N, H, W, C, reinterpret_cast<bf16*>(y));
    else
        upsample2x_bwd_kernel<<<ew_grid((long long)N * H * W * (C / 8), 256), 256, 0, stream>>>(
            reinterpret_cast<const bf16*>(x), N, H, W, C, reinterpret_cast<bf16*>(y));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_phase_split2(const void* src, int32_t N, int32_t H, int32_t W, int32_t C, int32_t inverse, void* dst,
                                void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(N >= 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, "uwu_phase_split2: bad shape");
    if (N == 0) return UWU_OK;
    UWU_CHECK_ARG(src && dst, "uwu_phase_split2: null pointer");
    phase_split_kernel<<<ew_grid((long long)N * H * W * (C / 8), 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(src), N, H, W, C, inverse, reinterpret_cast<bf16*>(dst));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int64_t uwu_colsum_workspace_floats(int64_t M, int32_t C) {
    if (M <= 0 || C <= 0) return 0;
    const int64_t P = (M + 511) / 512;
    return P * C;
}
extern "C" int uwu_colsum_bf16(const void* x, int64_t M, int32_t C, int64_t ld, int32_t accumulate, float* out,
                               float* workspace, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(M > 0 && C > 0 && ld >= C, "uwu_colsum_bf16: bad shape");
    UWU_CHECK_ARG(x && out && workspace, "uwu_colsum_bf16: null pointer");
    const int P = (int)((M + 511) / 512);
    const bool vec = C % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (vec)
        colsum_bf16_kernel<true><<<dim3((C / 8 + 63) / 64, P), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), M, C, ld, 512, workspace);
    else
        colsum_bf16_kernel<false><<<dim3((C + 255) / 256, P), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), M, C, ld, 512, workspace);
    UWU_CHECK_LAUNCH();
    colsum_final_kernel<<<(C + 31) / 32, 256, 0, stream>>>(workspace, P, C, accumulate, out);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}

extern "C" int uwu_im2col3x3(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride, void* cols, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(x && cols && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
                  "uwu_im2col3x3: bad arguments (C %% 8 == 0, stride 1 or 2)");
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    im2col3x3_kernel<<<ew_grid((long long)N * Ho * Wo * (C / 8), 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(x), N, H, W, C, stride, Ho, Wo, reinterpret_cast<bf16*>(cols));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_conv_pack(const float* W, int32_t Co, int32_t Ci, int32_t taps, int32_t ci_pad, int32_t co_pad, int32_t cod_pad,
                             void* fwd_bf16, void* dgrad_bf16, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(W && fwd_bf16 && Co > 0 && Ci > 0 && taps > 0 && ci_pad >= Ci && co_pad >= Co && (!dgrad_bf16 || cod_pad >= Co),
                  "uwu_conv_pack: bad arguments");
    UWU_CHECK_ARG(taps <= 16, "uwu_conv_pack: at most 16 taps");
    const int cmax = co_pad > cod_pad ? co_pad : cod_pad;
    dim3 grid((ci_pad + 31) / 32, (cmax + 31) / 32);
    conv_pack_kernel<<<grid, 256, (size_t)32 * (32 * taps + 1) * sizeof(float), stream>>>(
        W, Co, Ci, taps, ci_pad, co_pad, cod_pad, reinterpret_cast<bf16*>(fwd_bf16), reinterpret_cast<bf16*>(dgrad_bf16));
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_conv_wgrad_unpack(const float* G, int64_t ldg, int32_t Co, int32_t Ci, int32_t Ci_pad, int32_t taps,
                                     int32_t accumulate, float* wgrad, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(G && wgrad && Co > 0 && Ci > 0 && Ci_pad >= Ci && taps > 0 && ldg >= (int64_t)taps * Ci_pad,
                  "uwu_conv_wgrad_unpack: bad arguments");
    UWU_CHECK_ARG(taps <= 9, "uwu_conv_wgrad_unpack: at most 9 taps");
    conv_wgrad_unpack_kernel<<<ew_grid((long long)Co * ((Ci + 31) / 32) * 32, 256), 256, 0, stream>>>(G, ldg, Co, Ci, Ci_pad, taps, accumulate,
                                                                                                 wgrad);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
extern "C" int uwu_colsum_groups_bf16(const void* x, int64_t ldx, int32_t groups, int32_t rows, int32_t C, int32_t accumulate,
                                      float* out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    UWU_CHECK_ARG(x && out && groups > 0 && rows > 0 && C > 0 && ldx >= C, "uwu_colsum_groups_bf16: bad arguments");
    colsum_groups_kernel<<<dim3((C + 63) / 64, groups), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), ldx, rows, C, accumulate, out);
    UWU_CHECK_LAUNCH();
    return UWU_OK;
}
