/*
 * uwu_b200.h — C ABI of libuwu_b200.so: the sm_100a kernels behind the duwu diffusion training step.
 *
 * Conventions
 *   - plain C types only; device pointers are `void*` / typed pointers to DEVICE memory,
 *   - every entry point is asynchronous on the `stream` argument (a cudaStream_t passed as void*),
 *     never synchronises, never allocates, never keeps a pointer after returning,
 *   - return value: 0 = OK, <0 = invalid argument / unsupported shape, >0 = cudaError_t;
 *     the message is available through uwu_last_error() (thread-local),
 *   - there is NO CPU fallback: an unsupported shape is an error, a missing GPU is an error.
 *
 * Each function cites the reference call site it replaces (paths relative to /root/reference).
 */
#ifndef UWU_B200_H
#define UWU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UWU_OK 0
#define UWU_ERR_INVALID (-1)
#define UWU_ERR_UNSUPPORTED (-2)

/* dtype codes */
#define UWU_F32 0
#define UWU_BF16 1
#define UWU_I64 2

/* target / prediction types (src/duwu/loss/diffusion.py:84-98) */
#define UWU_TARGET_EPSILON 0
#define UWU_TARGET_V 1
#define UWU_TARGET_SAMPLE 2
#define UWU_TARGET_RF 3

/* loss-weight flags (src/duwu/loss/diffusion.py:141-167) */
#define UWU_WEIGHT_MIN_SNR 1
#define UWU_WEIGHT_DEBIASED 2
#define UWU_WEIGHT_EDM 4 /* lambda(sigma) = (sigma^2 + sigma_data^2) / (sigma * sigma_data)^2 in the first weight row (north_star a;
                            Karras et al. 2022 — the reference ships no EDM weighting, this is an additive mode) */

const char* uwu_last_error(void);
int uwu_version(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
int64_t uwu_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM / implicit-GEMM convolution (tcgen05 + TMEM + TMA)
 *   replaces torch.nn.functional.linear / conv2d under diffusers' UNet2DConditionModel, i.e. the
 *   `unet(noisy_latent, timesteps, **unet_kwargs)` call at src/duwu/loss/diffusion.py:172-176, and the
 *   LyCORIS `F.linear(x, W + dW)` forward patched in at src/duwu/trainer/trainer.py:152-154.
 * ------------------------------------------------------------------------------------------------ */
#define UWU_A_ROW 0  /* A[M,K], K contiguous                                  */
#define UWU_A_COL 1  /* A stored [K,M], M contiguous (reduction over rows)    */
#define UWU_A_CONV 2 /* A = NHWC activation [n_img_buf,H,W,Cin] + tap table   */
#define UWU_B_NK 0   /* B[N,K], K contiguous  (y = x W^T)                     */
#define UWU_B_KN 1   /* B stored [K,N], N contiguous                          */

typedef struct uwu_gemm_desc {
    const void* a;  /* bf16 */
    const void* a2; /* bf16, second channel-concatenated conv source or NULL */
    const void* b;  /* bf16 */
    int64_t M, N, K;
    int32_t a_layout, b_layout;
    int64_t lda, ldb; /* leading dimensions in elements (ignored for UWU_A_CONV) */
    /* conv geometry (UWU_A_CONV): output pixel m = ((n*H + h)*W + w) reads input pixel
       (n + tap_dn[t], h + tap_dh[t], w + tap_dw[t]) for tap t; out-of-range h/w read zeros.
       K index = t*(Cin1+Cin2) + c. */
    int32_t n_img_buf, H, W, Cin1, Cin2, ntaps;
    int32_t tap_dn[9], tap_dh[9], tap_dw[9];
    /* epilogue: out[m,n] = alpha*acc + bias[n] + bias_rows[m/rows_per_bias, n] + residual[m,n] */
    void* out;
    void* out2; /* columns >= n_split are written to out2[m, n - n_split] (or NULL) */
    int64_t ldo, ldo2;
    int32_t n_split;
    int32_t out_dtype; /* UWU_BF16 or UWU_F32 */
    const float* bias;
    const float* bias_rows;
    int32_t rows_per_bias;
    const void* residual; /* bf16 [M, ldr] or NULL */
    int64_t ldr;
    float alpha;
    int32_t accumulate; /* fp32 output only: out += result */
    int32_t block_n;    /* 0 = choose */
    int32_t stream_k;   /* fp32 output, plain epilogue: 1 = split the (tile, k-block) space evenly over the SMs and
                           reduce partial tiles with atomics; 2 = cut the reduction into slices (count chosen so that
                           slices x tiles fills whole waves) dealt round-robin in slice-major order, so concurrent CTAs
                           share one k-slice through L2; 0 = whole tiles per CTA; -1 = choose between 0 and 2 */
    /* segmented reduction (UWU_A_COL x UWU_B_KN): out = sum_{s < k_segs} A[:, s*a_seg_off + m]^T B[:, s*b_seg_off + n];
       K is the length of ONE segment. Used for the factored LoKr gradient dw2 = sum_l dY_l^T Z_l. 0/1 = off. */
    int32_t k_segs, a_seg_off, b_seg_off;
    /* grouped N (UWU_A_ROW): out[:, g*grp_n + n] = A[:, g*a_grp_koff : +K] B[:, n] with ONE right operand — [K, grp_n] (UWU_B_KN)
       or [grp_n, K] (UWU_B_NK) — shared by all N/grp_n groups (block-diagonal right operand). Used for V_l = dY_l w2 and
       T_j = X_j w2^T of the factored LoKr gradients. 0 = off. */
    int32_t grp_n, a_grp_koff;
    /* diagnostics: override shared-memory descriptor fields (0 = default) */
    int32_t dbg_a_lbo, dbg_a_sbo, dbg_a_kadv, dbg_b_lbo, dbg_b_sbo, dbg_b_kadv;
    /* fused GEGLU epilogues (diffusers GEGLU: hidden, gate = proj(x).chunk(2, -1); hidden * gelu(gate) — the FeedForward of
       BasicTransformerBlock, src/duwu/modules/rope_unet.py:395-404), CTA-pair kernel only (M >= 256), bf16 output:
       1 = forward  : N = 2F; `out` [M, 2F] receives the pre-activation (h | g), `out2` [M, F] receives h * gelu(g);
       2 = backward : N = F, the GEMM result is d = dL/d(h * gelu(g)); `aux` = saved pre-activation [M, ld_aux >= 2F];
                      `out` [M, 2F] receives (d * gelu(g) | d * h * gelu'(g)).
       Results are bit-identical to uwu_gemm followed by uwu_geglu_fwd / uwu_geglu_bwd. */
    int32_t epi_mode;
    const void* aux;
    int64_t ld_aux;
} uwu_gemm_desc;

int uwu_gemm(const uwu_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused noising: timestep sampling (or injected), sigma gather, x_t, target, loss weights, t-embedding.
 *   replaces DiffusionLoss.get_noise_noisy_latents_and_timesteps (src/duwu/loss/diffusion.py:74-82),
 *   sample_timesteps_and_sigmas (:64-72), get_target (:84-98), the weight factors of apply_snr_weight
 *   (:141-153) and apply_debiased_estimation (:155-167), and diffusers' Timesteps(dim, flip=True, shift=0).
 *   Arithmetic follows the reference op by op (one rounding per op, no FMA contraction) so that with
 *   injected eps/t the fp32 outputs are bit-identical to the PyTorch reference.
 * ------------------------------------------------------------------------------------------------ */
typedef struct uwu_noise_desc {
    const void* x0;       /* [B, n_per] latents, dtype below */
    const void* eps_in;   /* optional injected noise, same dtype/shape (NULL => in-kernel Philox4x32-10) */
    const int64_t* t_in;  /* optional injected timesteps [B] (NULL => Philox) */
    uint64_t seed, offset;
    const float* acp;     /* scheduler.alphas_cumprod [T] */
    const float* sigma_t; /* sigma of TIMESTEP t, [T] (= scheduler.sigmas[T-1-t]) */
    const float* snr;     /* scheduler.all_snr [T] (src/duwu/loss/diffusion.py:42-51) */
    int32_t T, B;
    int64_t n_per;
    int32_t dtype;        /* UWU_F32 | UWU_BF16: dtype of x0/eps/x_t/target */
    int32_t target_type;  /* UWU_TARGET_* */
    int32_t pred_type;    /* UWU_TARGET_* (selects the min-SNR formula) */
    int32_t weight_flags; /* UWU_WEIGHT_* */
    float gamma;          /* min_snr_gamma */
    void* x_t;            /* out */
    void* target;         /* out */
    void* eps_out;        /* out, optional */
    int64_t* t_out;       /* out [B] */
    float* sigma_out;     /* out [B] */
    float* w_out;         /* out [2,B]: min-SNR factor, debiased factor (1.0 when disabled) */
    void* temb_out;       /* out, optional bf16 [B, temb_dim] = [cos | sin] */
    int32_t temb_dim;
    const float* sigma_in; /* optional fp32 [B]: per-sample sigma used instead of sigma_t[t] (RectifiedFlowLoss "uniform_time"
                              sampling, src/duwu/loss/rectified_flow.py:27-45,63-71) */
    float sigma_data;      /* UWU_WEIGHT_EDM: sigma_data (0.5 in the EDM paper) */
    const uint64_t* step_dev; /* optional device counter ADDED to `offset` when the kernel runs: a launch captured in a CUDA graph
                                 keeps drawing fresh noise / timesteps on every replay */
} uwu_noise_desc;

int uwu_noise_fwd(const uwu_noise_desc* desc, void* stream);

/* Timesteps(dim)(vals) -> bf16 [n, dim]; SDXL add_time_proj on time_ids (diffusers, restated). */
int uwu_sincos_embed(const float* vals, int32_t n, int32_t dim, int32_t flip_sin_to_cos, void* out_bf16, void* stream);

/* Weighted MSE (src/duwu/loss/diffusion.py:179-193): losses[b] = w[1,b] * (mean_i (pred-target)^2 * w[0,b]),
 * loss = mean_b losses[b].  `workspace` holds uwu_wmse_workspace_floats(B, n_per) floats. Deterministic. */
int64_t uwu_wmse_workspace_floats(int32_t B, int64_t n_per);
int uwu_wmse_fwd(const void* pred, int32_t pred_dtype, const void* target, int32_t target_dtype, int32_t B,
                 int64_t n_per, const float* w, float* workspace, float* losses, float* loss, void* stream);
/* dpred = grad_scale * (*grad_scale_dev or 1) * w0*w1 * 2 (pred - target) / (n_per * B) */
/* Per-timestep validation statistics: counts[t] += 1, sums[t] += l, sqsums[t] += l*l for every sample (one scatter-add launch;
 * replaces the N_t boolean-mask loop of PlotValLossPerTimestep.on_validation_batch_end, src/duwu/trainer/callbacks.py:75-92).
 * timesteps: int64 (UWU_I64) or fp32 (UWU_F32, truncated like `.long()`); out-of-range values are ignored. */
int uwu_timestep_hist(const float* losses, const void* timesteps, int32_t t_dtype, int32_t B, int32_t T, float* counts,
                      float* sums, float* sqsums, void* stream);
int uwu_wmse_bwd(const void* pred, int32_t pred_dtype, const void* target, int32_t target_dtype, int32_t B,
                 int64_t n_per, const float* w, const float* grad_scale_dev, float grad_scale, void* dpred,
                 int32_t dpred_dtype, void* stream);

/* Prediction -> target space when prediction_type != target_type: get_x0_eps_from_pred_with_sigmas + get_target
 * (src/duwu/loss/diffusion.py:100-139, called at :177 with the CLEAN latents x — kept bug-compatible).  Per sample the map is
 * linear: result = A_b * out + C_b * x (forward), result = A_b * out (backward = gradient w.r.t. the model output).
 * sigma: fp32 [B] (the noising kernel's sigma_out), t: int64 [B], acp: alphas_cumprod table. */
int uwu_pred_convert(const float* out, const void* x, int32_t x_dtype, const float* sigma, const int64_t* t, const float* acp,
                     int32_t B, int64_t n_per, int32_t pred_type, int32_t target_type, int32_t backward, float* result,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Flash attention forward / backward (tcgen05 + TMEM + TMA), head_dim 64.
 *   replaces F.scaled_dot_product_attention in diffusers' AttnProcessor2_0 (in-tree copy of the flow:
 *   src/duwu/modules/rope_unet.py:76-175, SDPA call :151) for attn1 (self) and attn2 (cross, Lk = 77).
 *   q/o: bf16 [B*Lq, ld], k/v: bf16 [B*Lk, ld]; head h occupies columns [64h, 64h+64).
 *   lse: fp32 [B, heads, roundup(Lq,128)] (uwu_attn_lse_floats), natural-log-sum-exp of scale*QK^T.
 *   backward workspace: uwu_attn_bwd_workspace_floats floats, 16-byte aligned.
 * ------------------------------------------------------------------------------------------------ */
int64_t uwu_attn_lse_floats(int32_t B, int32_t heads, int32_t Lq);
int uwu_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t heads, int32_t Lq,
                 int32_t Lk, int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale,
                 void* stream);
/* forward-only attention with a causal mask and / or a per-(batch, key) padding mask, for the frozen CLIP text towers behind
 * ConcatTextEncoders (src/duwu/modules/text_encoders.py:139-200 calls transformers.CLIPTextModel(input_ids, attention_mask=...),
 * whose self-attention is causal + padding-masked).  Lk <= 128, head_dim <= 64.  key_mask: int32 [B, Lk], 0 = padding, or NULL. */
int uwu_attn_fwd_masked(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t heads, int32_t Lq,
                        int32_t Lk, int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale,
                        int32_t causal, const int32_t* key_mask, void* stream);
int64_t uwu_attn_bwd_workspace_floats(int32_t B, int32_t heads, int32_t Lq);
int uwu_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq,
                 void* dk, void* dv, int32_t B, int32_t heads, int32_t Lq, int32_t Lk, int32_t head_dim, int64_t ldq,
                 int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv, float scale,
                 float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bandwidth-bound glue over channels-last bf16 activations (fp32 statistics and parameters).
 *   replaces ATen group_norm / layer_norm / silu / gelu / add / upsample kernels under diffusers'
 *   ResnetBlock2D, Transformer2DModel, BasicTransformerBlock (in-tree copy of the block algebra:
 *   src/duwu/modules/rope_unet.py:288-415), GEGLU, Upsample2D, Downsample2D [third-party, restated in oracle/].
 * ------------------------------------------------------------------------------------------------ */
/* GroupNorm(+SiLU) on x[N, HW, C]; stats[N, G, 2] = {mean, rstd}; workspace: uwu_groupnorm_workspace_floats */
int64_t uwu_groupnorm_workspace_floats(int32_t N, int32_t HW, int32_t C, int32_t G);
int uwu_groupnorm_fwd(const void* x, int32_t N, int32_t HW, int32_t C, int32_t G, float eps, const float* gamma,
                      const float* beta, int32_t fuse_silu, void* y, float* stats, float* workspace, void* stream);
/* dx = GN'(dy) (+ dres); dgamma/dbeta (optional) are ACCUMULATED into */
int uwu_groupnorm_bwd(const void* x, const void* dy, int32_t N, int32_t HW, int32_t C, int32_t G, const float* gamma,
                      const float* beta, const float* stats, int32_t fuse_silu, const void* dres, void* dx,
                      float* dgamma, float* dbeta, float* workspace, void* stream);
/* One-launch variants: statistics pass -> per-image barrier between the co-resident blocks of the grid -> apply pass that
 * re-reads the block's own rows from L2 (one HBM pass less in forward, two less in backward).  `sync`: 2 * N uint32 in device
 * memory, ZERO before the first call; every call leaves them ready for the next (persistent per-device buffer).  The
 * backward variant is for a frozen affine (no dgamma / dbeta).  Return 1 = shape cannot run as one resident wave, nothing was
 * launched: call the three-kernel entry point instead.  Same reference call sites as uwu_groupnorm_fwd / _bwd. */
int uwu_groupnorm_fwd_fused(const void* x, int32_t N, int32_t HW, int32_t C, int32_t G, float eps, const float* gamma,
                            const float* beta, int32_t fuse_silu, void* y, float* stats, float* workspace, uint32_t* sync,
                            void* stream);
int uwu_groupnorm_bwd_fused(const void* x, const void* dy, int32_t N, int32_t HW, int32_t C, int32_t G, const float* gamma,
                            const float* beta, const float* stats, int32_t fuse_silu, const void* dres, void* dx,
                            float* workspace, uint32_t* sync, void* stream);
/* LayerNorm on x[M, C] with optional adaLN modulation y = LN(x) * (1 + mod_scale[m / rows_per_mod]) + mod_shift[..];
 * stats[M, 2] = {mean, rstd} (optional) */
int uwu_layernorm_fwd(const void* x, int32_t M, int32_t C, float eps, const float* gamma, const float* beta,
                      const float* mod_scale, const float* mod_shift, int32_t rows_per_mod, void* y, float* stats,
                      void* stream);
int64_t uwu_layernorm_bwd_workspace_floats(int32_t M, int32_t C);
int uwu_layernorm_bwd(const void* x, const void* dy, int32_t M, int32_t C, const float* gamma, const float* stats,
                      const void* dres, void* dx, float* dgamma, float* dbeta, int32_t accumulate, float* workspace,
                      void* stream);

/* in-place row softmax of a bf16 [M, ld >= N] matrix: the single-head (head_dim 512) attention of the frozen VAE encoder
 * (`AutoencoderKL.encode`, src/duwu/trainer/trainer.py:241-244), whose S = Q K^T and P V products run on uwu_gemm */
int uwu_softmax_rows(void* x, int64_t M, int32_t N, int64_t ld, void* stream);
/* GEGLU: out[m, f] = in[m, f] * gelu_erf(in[m, F + f])  (hidden, gate = proj.chunk(2)) */
int uwu_geglu_fwd(const void* in, int64_t M, int32_t F, void* out, void* stream);
int uwu_geglu_bwd(const void* in, const void* dout, int64_t M, int32_t F, void* din, void* stream);
/* mode 0: y = silu(x); 1: y = x * silu'(a); 2: y = x + a; 3: y = x; 4: y = gelu_tanh(x); 5: y = x * gelu_tanh'(a)
 * (bf16, n multiple of 8) */
int uwu_elementwise(const void* x, const void* a, int64_t n, int32_t mode, void* y, void* stream);
/* ------------------------------------------------------------------------------------------------
 * adaLN-Zero glue of DiT blocks (BASELINE.json configs[3]).  The reference tree has no DiT model; the algebra is the
 * `ada_norm_zero` branch of its patched transformer block: src/duwu/modules/rope_unet.py:306-309 (modulate), :344-345
 * (gate_msa), :395-398, :406-407 (MLP side).  x/y/dx: bf16 [M, C], M = B * rows_per_mod; mod: fp32 [B, ld_mod] rows holding
 * shift / scale / gate windows at the given column offsets; dmod: bf16 [B, ld_dmod] (per-sample sums over the rows).
 * ------------------------------------------------------------------------------------------------ */
/* y = LN(x) * (1 + scale[b]) + shift[b]  (no affine); stats[M, 2] = {mean, rstd} */
int uwu_adaln_fwd(const void* x, int64_t M, int32_t C, float eps, const float* mod, int64_t ld_mod, int32_t shift_off,
                  int32_t scale_off, int32_t rows_per_mod, void* y, float* stats, void* stream);
/* dx = LN'(dy * (1 + scale[b])) (+ dres); dmod[b, dshift_off + c] = sum_t dy, dmod[b, dscale_off + c] = sum_t dy * xhat */
int uwu_adaln_bwd(const void* x, const void* dy, int64_t M, int32_t C, const float* mod, int64_t ld_mod, int32_t scale_off,
                  const float* stats, int32_t rows_per_mod, const void* dres, void* dx, void* dmod_bf16, int64_t ld_dmod,
                  int32_t dshift_off, int32_t dscale_off, void* stream);
/* out = x + gate[b] * y */
int uwu_gate_residual_fwd(const void* x, const void* y, int64_t M, int32_t C, const float* mod, int64_t ld_mod,
                          int32_t gate_off, int32_t rows_per_mod, void* out, void* stream);
/* dy = gate[b] * dout; dmod[b, dgate_off + c] = sum_t dout * y */
int uwu_gate_residual_bwd(const void* dout, const void* y, int64_t M, int32_t C, const float* mod, int64_t ld_mod,
                          int32_t gate_off, int32_t rows_per_mod, void* dy, void* dmod_bf16, int64_t ld_dmod,
                          int32_t dgate_off, void* stream);
/* NCHW fp32 image <-> token rows [B * (H/p) * (W/p), ld].  order 0: column = c*p*p + ph*p + pw (patch-embedding conv
 * weight flattening); order 1: column = (ph*p + pw)*Ctok + c (DiT unpatchify).  patchify zero-fills channels >= Cimg and
 * columns >= Ctok*p*p; unpatchify takes the first Cimg of Ctok channels. */
int uwu_patchify(const float* img, int32_t B, int32_t Cimg, int32_t H, int32_t W, int32_t p, int32_t order, int32_t Ctok,
                 void* tok_bf16, int64_t ld, void* stream);
int uwu_unpatchify(const void* tok, int32_t tok_dtype, int64_t ld, int32_t B, int32_t Cimg, int32_t H, int32_t W, int32_t p,
                   int32_t order, int32_t Ctok, float* img, void* stream);
/* class-label embedding table: out[b] = table[idx[b]] (bf16 rows); dtable[idx[b]] += dout[b] */
int uwu_embed_gather(const float* table, const int64_t* idx, int32_t B, int32_t D, int32_t V, void* out_bf16, void* stream);
int uwu_embed_scatter_add(const void* dout_bf16, const int64_t* idx, int32_t B, int32_t D, int32_t V, float* dtable,
                          void* stream);

/* API boundary layout conversion: the reference tensors are NCHW (src/duwu/data/base.py:13) */
int uwu_nchw_to_nhwc(const void* src, int32_t src_dtype, int32_t N, int32_t C, int32_t HW, int32_t Cpad, void* dst_bf16,
                     void* stream);
int uwu_nhwc_to_nchw(const void* src, int32_t src_dtype, int32_t N, int32_t C, int32_t HW, int64_t ld, float* dst,
                     void* stream);
/* nearest 2x upsample (backward = 1: sum of the 2x2 children) */
int uwu_upsample2x(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t backward, void* y, void* stream);
/* space-to-depth by 2 into 4 phase planes [(py*2+px)*N + n, H/2, W/2, C] (stride-2 conv input); inverse = 1 scatters back */
int uwu_phase_split2(const void* src, int32_t N, int32_t H, int32_t W, int32_t C, int32_t inverse, void* dst, void* stream);
/* column sums of a bf16 matrix (bias gradients) */
int64_t uwu_colsum_workspace_floats(int64_t M, int32_t C);
int uwu_colsum_bf16(const void* x, int64_t M, int32_t C, int64_t ld, int32_t accumulate, float* out, float* workspace,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * LyCORIS adapters (LoRA / LoKr full_matrix / norm deltas) and the optimizer.
 *   replaces lycoris-lora's patched forward `F.linear(x, W + dW * multiplier)` and its autograd backward through
 *   kron / up@down (wrapper applied at src/duwu/trainer/trainer.py:148-169; preset configs/lycoris/sdxl-diffusers.toml),
 *   torch.optim.AdamW (src/duwu/trainer/trainer.py:52-74; configs/demo_training_lycoris.yaml:48-55) and Lightning's
 *   gradient_clip_val (configs/demo_training_lycoris.yaml:13).
 *   Shape bookkeeping is checked exactly: (out_l*out_k, in_m*in_n) must equal (N, K).
 * ------------------------------------------------------------------------------------------------ */
/* dst[N,K] bf16 = W + kron(w1[out_l,in_m], w2[out_k,in_n]) * multiplier   (w1 == NULL: plain fp32 -> bf16 cast) */
int uwu_fold_lokr(const float* W, const float* w1, const float* w2, int32_t N, int32_t K, int32_t out_l, int32_t out_k,
                  int32_t in_m, int32_t in_n, float multiplier, void* dst_bf16, void* stream);
/* dst[N,K] bf16 = W + (up[N,r] @ down[r,K]) * scale */
int uwu_fold_lora(const float* W, const float* up, const float* down, int32_t N, int32_t K, int32_t r, float scale,
                  void* dst_bf16, void* stream);
/* out = a + alpha * b (fp32; effective norm affine = gamma + w_norm * multiplier) */
int uwu_axpy_f32(const float* a, const float* b, float alpha, int32_t n, float* out, void* stream);

/* Every adapter fold of a step in one launch.  `entries` (device memory) describe the folds; chunk c of the grid works on
 * entry chunk_entry[c], elements [(c - chunk0) * chunk_elems, +chunk_elems) of its N*K weight (K % 4 == 0).
 *   kind 0: dst_bf16 = W                         kind 1: dst_bf16 = W + kron(a [N/p0, p2], b [p0, p1]) * scale   (LoKr)
 *   kind 2: dst_bf16 = W + (a [N, p0] @ b [p0, K]) * scale (LoRA)   kind 3: dst_f32 = W + a * scale  (norm deltas, N = 1)
 *   kind 4: dst_bf16 = W + ((a [N, p0] @ b [p0, K]) o ((a + p1) [N, p0] @ (a + p2) [p0, K])) * scale (LoHa; the second factor
 *           pair is addressed by element offsets from `a`: all adapter parameters live in one flat buffer)
 * Same arithmetic as uwu_fold_lokr / uwu_fold_lora / uwu_axpy_f32 (bit-identical results). */
typedef struct uwu_fold_entry {
    const float* W;
    const float* a;
    const float* b;
    void* dst;
    int32_t kind, N, K, p0, p1, p2;
    float scale;
    int32_t chunk0;
} uwu_fold_entry;
int uwu_fold_batch(const uwu_fold_entry* entries_dev, const int32_t* chunk_entry_dev, int32_t n_chunks, int32_t chunk_elems,
                   void* stream);

/* Full fine-tuning (trainable base weights, trainer.py:160-169 `self.unet.requires_grad_(True)` when lycoris_config is None):
 * the 3x3 convolution weight gradient is the token-reduction GEMM dW = dY^T im2col(X) (uwu_gemm, A_COL x B_KN, stream-K);
 *   uwu_im2col3x3        : cols[(n,ho,wo), t*C + c] = x[n, ho*stride + t/3 - 1, wo*stride + t%3 - 1, c] (zero padded), bf16 NHWC
 *   uwu_conv_wgrad_unpack: wgrad[co, ci, t] (+)= G[co, t*Ci_pad + ci]   (torch Conv2d layout <- packed GEMM layout)
 *   uwu_colsum_groups_bf16: out[g, c] (+)= sum_r x[g*rows + r, c]        (per-image sums: gradient of the time-embedding bias rows)
 * Replaces autograd's conv2d weight gradient / sum reductions under loss.backward() in the reference's full fine-tuning mode. */
/* Conv2d master weight [Co, Ci, kh, kw] fp32 -> bf16 implicit-GEMM operands (zero-padded): fwd [co_pad, taps*ci_pad] with
 * fwd[co, t*ci_pad + ci] = W[co, ci, t], and (optional) dgrad [Ci, taps*cod_pad] with dgrad[ci, t*cod_pad + co] =
 * W[co, ci, taps-1-t] (spatially flipped kernel, channels swapped).  One launch per convolution per step in full fine-tuning. */
int uwu_conv_pack(const float* W, int32_t Co, int32_t Ci, int32_t taps, int32_t ci_pad, int32_t co_pad, int32_t cod_pad,
                  void* fwd_bf16, void* dgrad_bf16, void* stream);
int uwu_im2col3x3(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride, void* cols, void* stream);
int uwu_conv_wgrad_unpack(const float* G, int64_t ldg, int32_t Co, int32_t Ci, int32_t Ci_pad, int32_t taps, int32_t accumulate,
                          float* wgrad, void* stream);
int uwu_colsum_groups_bf16(const void* x, int64_t ldx, int32_t groups, int32_t rows, int32_t C, int32_t accumulate, float* out,
                           void* stream);
/* The same contraction for many adapters in ONE launch (each adapter keeps its own G buffer until the flush).
 * Table entry: vec = 1 when in_n % 4 == 0, ldg % 4 == 0 and G / w2 are 16-byte aligned; target = block budget of the entry;
 * block0 = first block of the entry (prefix sum of uwu_lokr_grad_plan_blocks over the table). */
typedef struct uwu_lokr_grad_entry {
    const float* G;
    const float* w1;
    const float* w2;
    float* dw1;
    float* dw2;
    int64_t ldg;
    int32_t out_l, out_k, in_m, in_n;
    float multiplier;
    int32_t vec, target, block0;
} uwu_lokr_grad_entry;
int32_t uwu_lokr_grad_plan_blocks(int32_t out_l, int32_t out_k, int32_t in_m, int32_t in_n, int32_t vec, int32_t target);
int uwu_lokr_grad_batch(const uwu_lokr_grad_entry* entries_dev, int32_t n_entries, int32_t total_blocks, void* stream);
/* LoHa (lycoris `loha`): dW = ((w1a @ w1b) o (w2a @ w2b)) * scale, factors [N, r] / [r, K], r <= 16 */
int uwu_fold_loha(const float* W, const float* w1a, const float* w1b, const float* w2a, const float* w2b, int32_t N, int32_t K,
                  int32_t r, float scale, void* dst_bf16, void* stream);
int uwu_loha_grad(const float* G, int64_t ldg, const float* w1a, const float* w1b, const float* w2a, const float* w2b, int32_t N,
                  int32_t K, int32_t r, float scale, float* dw1a, float* dw1b, float* dw2a, float* dw2b, void* stream);
/* adapter gradients from G = dY^T X (fp32 [N, ldg]); dw1/dw2 (dup/ddown) are ACCUMULATED into */
int uwu_lokr_grad(const float* G, int64_t ldg, const float* w1, const float* w2, int32_t out_l, int32_t out_k, int32_t in_m,
                  int32_t in_n, float multiplier, float* dw1, float* dw2, void* stream);
int uwu_lora_grad(const float* G, int64_t ldg, const float* up, const float* down, int32_t N, int32_t K, int32_t r, float scale,
                  float* dup, float* ddown, void* stream);

/* Factored LoKr gradients: the adapter gradients of  y = x (W + kron(w1, w2))^T  without forming
 * G = dY^T X.  x: bf16 [M, in_m*in_n] (row stride ldx), w1: fp32 [out_l, in_m].
 *   uwu_lokr_z  : z[m, l*in_n + n] = sum_i w1[l, i] x[m, i*in_n + n]            (bf16 [M, out_l*in_n], contiguous;
 *                 w1_transposed: w1 is stored [in_m, out_l] and mixes as its transpose — the dY-side transform
 *                 U[m, j, p] = sum_i w1[i, j] dY[m, i, p] of the mirrored route, used when out_k < in_n)
 *   uwu_lokr_dw1: dw1[l, i] += multiplier * sum_{m, n} v[m, l*in_n + n] x[m, i*in_n + n]   (v: bf16 [M, out_l*in_n])
 * The two tensor-core steps in between (dw2 += sum_l dY_l^T Z_l and V_l = dY_l w2) are uwu_gemm calls with k_segs / grp_n.
 * Replaces autograd through lycoris' `make_kron` weight rebuild (forward patch installed at src/duwu/trainer/trainer.py:152-154). */
int uwu_lokr_z(const void* x, int64_t ldx, const float* w1, int64_t M, int32_t out_l, int32_t in_m, int32_t in_n, void* z,
               int32_t w1_transposed, void* stream);
int uwu_lokr_dw1(const void* v, const void* x, int64_t ldx, int64_t M, int32_t out_l, int32_t in_m, int32_t in_n,
                 float multiplier, float* dw1, void* stream);
/* One-pass LoKr gradients for attention projections (w2 64x64, w1 <= 32x32): reads x [M, in_m*64] and dY [M, out_l*64] (bf16, row
 * strides ldx / ldy) ONCE, forms Z = (w1 (x) I) X and V = X w2^T per 128-row tile on tcgen05, and accumulates
 *   dw2 += multiplier * sum dY^T Z,   dw1[i, j] += multiplier * sum_t <dY[t, i, :], V[t, j, :]>
 * in TMEM; results are added atomically into dw1 [out_l, in_m] / dw2 [64, 64] (fp32).  w1 / w2 are the fp32 masters (rounded
 * to bf16 as MMA operands).  Replaces autograd through lycoris' `make_kron(w1, w2)` rebuild for the `Attention -> lokr,
 * factor = 64` entries of configs/lycoris/sdxl-diffusers.toml (forward patch: src/duwu/trainer/trainer.py:152-154). */
int uwu_lokr_fused_supported(int32_t out_l, int32_t out_k, int32_t in_m, int32_t in_n);
int uwu_lokr_fused_grad(const void* x, int64_t ldx, const void* dy, int64_t ldy, int64_t M, int32_t out_l, int32_t in_m,
                        const float* w1, const float* w2, float* dw1, float* dw2, float multiplier, void* stream);
/* multi-tensor global grad norm: out2 = {||g||_2, min(1, max_norm/(norm+1e-6))}; tables are DEVICE arrays
 * (pointers as uint64, numels, and a chunk table (tensor index, chunk index) of n_chunks entries) */
int uwu_mt_gradnorm(const uint64_t* g_ptrs, const int64_t* numels, const int32_t* chunk_tensor, const int32_t* chunk_index,
                    int32_t n_chunks, int32_t chunk_elems, float max_norm, float* partial_ws, float* out2, void* stream);
/* torch.optim.AdamW step (decoupled weight decay, bias correction by `step`), gradients scaled by norm_clip[1] if given.
 * hyper_dev (optional, device fp32[3] = {lr, 1 - beta1^step, sqrt(1 - beta2^step)}) overrides lr / step when the kernel
 * runs, so that a launch captured in a CUDA graph follows the LR schedule (src/duwu/trainer/trainer.py:52-74) on replay */
int uwu_mt_adamw(const uint64_t* p_ptrs, const uint64_t* g_ptrs, const uint64_t* m_ptrs, const uint64_t* v_ptrs,
                 const int64_t* numels, const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t n_chunks,
                 int32_t chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                 const float* norm_clip, const float* hyper_dev, void* stream);
/* strided 2-D copy with cast to bf16 (skip-connection concat, conditioning staging) */
int uwu_copy2d_bf16(const void* src, int32_t src_dtype, int64_t lds, void* dst_bf16, int64_t ldd, int64_t rows, int32_t cols,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UWU_B200_H */
