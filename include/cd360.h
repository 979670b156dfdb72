/* cd360.h — C ABI of libcd360.so, the sm_100a (B200) kernels behind the pose-conditioned SDXL
 * UNet denoising step of customdiffusion360/custom-diffusion360.
 *
 * The reference has no FFI: its "plugin API" is Python (sgm module classes selected by `target:`
 * strings).  This header is the boundary WE define below that Python surface; every
 * entry point names the reference code whose arithmetic it replaces (paths relative to the
 * reference checkout).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says host;
 *   - the caller owns all buffers (inputs, outputs, workspaces); nothing is allocated, nothing
 *     synchronises, no global state is kept; launches go to `stream` and are CUDA-graph capturable;
 *   - activations are row-major bf16 "tokens x channels" ([B*H*W, C], i.e. NHWC) unless stated;
 *   - return value: CD360_OK or a negative CD360_ERR_* code; nothing is launched on error.
 */
#ifndef CD360_H_
#define CD360_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cd360_stream_t; /* cudaStream_t */

enum {
  CD360_OK = 0,
  CD360_ERR_SHAPE = -1,       /* unsupported / inconsistent shape */
  CD360_ERR_ALIGN = -2,       /* pointer or leading dimension not 16-byte aligned */
  CD360_ERR_UNSUPPORTED = -3, /* dtype / mode not built */
  CD360_ERR_LAUNCH = -4,      /* CUDA launch or driver error */
  CD360_ERR_NULL = -5         /* required pointer is NULL */
};

/* epilogue activations: SiLU (UNet), exact GELU (OpenCLIP text tower MLP), quick GELU x*sigmoid(1.702x)
 * (CLIP-L text tower MLP; sgm/modules/encoders/modules.py:377-517,622-772) */
enum { CD360_ACT_NONE = 0, CD360_ACT_SILU = 1, CD360_ACT_GELU = 2, CD360_ACT_QUICK_GELU = 3 };

/* Library / build info: returns the ABI version (bumped on any signature change). */
int cd360_abi_version(void);
/* Human-readable name of an error code (static string). */
const char* cd360_strerror(int code);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *
 *   out[M, N'] = epilogue( A[M, K] * W[N, K]^T )
 *
 * Replaces every nn.Linear / nn.Conv2d on the path: to_q/to_k/to_v/to_out
 * (sgm/modules/attention.py:323-329), GEGLU + FF out (attention.py:89-115), proj_in/proj_out
 * (attention.py:758,795), pose_emb_layers (attention.py:515), ResBlock/Downsample/Upsample 3x3 and
 * 1x1 convs (sgm/modules/diffusionmodules/openaimodel.py:142,215,283,320,337,720,971), FeatureNeRF
 * plane_coefs / decoder (sgm/modules/nerfsd_pytorch3d.py:40-51).
 *
 * A operand, linear mode (conv == 0): up to two K-segments, a0 [M, k0] (row stride lda0 elements)
 *   followed by a1 [M, k1] — the second segment is how a channel concatenation
 *   (th.cat([h, hs.pop()], 1), openaimodel.py:1074) is consumed without materialising it.
 * A operand, conv mode (conv == 1): a0 is an NHWC activation [B, H, W, C]; the kernel is an
 *   implicit GEMM over the 9 taps of a stride-1, pad-1 3x3 convolution, K = 9*C, W packed as
 *   [N, 9*C] with k = (ky*3+kx)*C + c.  Zero padding comes from TMA out-of-bounds fill.
 * W: bf16 [N, K] row-major (nn.Linear layout).  K segments must be multiples of 8.
 * Epilogue, in order: + bias[n] (fp32) ; + row_bias[m / rows_per_group, n] (fp32, the ResBlock
 *   timestep-embedding add, openaimodel.py:374) ; activation ; [GEGLU] ; + residual[m, n] (bf16) ;
 *   store bf16 or fp32.
 * GEGLU (geglu == 1): W/bias rows are pre-interleaved in blocks of `BN/2` so that every N tile
 *   holds [x-part | gate-part]; out[m, j] = x * gelu_erf(gate), out has N/2 columns
 *   (attention.py:94-96).  Use cd360_geglu_pack_block() to learn the block size.
 * --------------------------------------------------------------------------------------------- */
typedef struct cd360_gemm_args {
  const void* a0;
  int64_t lda0; /* elements; ignored in conv mode */
  int32_t k0;
  const void* a1;
  int64_t lda1;
  int32_t k1;
  const void* w;
  const float* bias;     /* [N] or NULL */
  const float* row_bias; /* [ceil(M / rows_per_group), N] or NULL */
  int32_t rows_per_group;
  int64_t ld_row_bias; /* elements; 0 = N */
  const void* residual; /* bf16 [M, N_out] or NULL */
  int64_t ldr;
  void* out; /* bf16 (out_fp32 == 0) or fp32 [M, N_out] */
  int64_t ldo;
  int32_t out_fp32;
  int32_t M, N;
  int32_t conv; /* 0 linear, 1 conv3x3 stride 1 pad 1 */
  int32_t B, H, W, C;
  int32_t act;   /* CD360_ACT_* */
  int32_t geglu; /* 0 / 1 */
  int32_t block_n; /* 0 = auto, 128 = single-CTA 128x128 tiles, 256/512 = CTA-pair 256x256 tiles, 1024 = pair tiles in 4-CTA clusters sharing W by TMA multicast */
  int32_t max_ctas; /* 0 = one per SM */
  /* LayerNorm folded into the contraction (linear mode, bf16 output).  With W' = W diag(gamma)
   * and bias' = bias + W beta supplied as `w` / `bias`:
   *     out[m, n] = rstd[m] * (acc[m, n] - mu[m] * ln_colsum[n]) + bias'[n]
   * where mu / rstd come from `ln_stats` = per-row partial (sum, sumsq) over 64-column slabs,
   * fp32 [M, ln_slabs, 2], written by the GEMM that produced A through `stats_out`
   * (fp32 [M, N_out/64, 2], moments of the bf16-rounded output rows; N_out % 64 == 0).
   * Replaces the three nn.LayerNorm of BasicTransformerBlock (attention.py:531-533, 609-636). */
  const float* ln_stats;
  int32_t ln_slabs;
  float ln_eps;
  const float* ln_colsum; /* [N] fp32: sum_k w[n, k] of the bf16 weight actually multiplied */
  float* stats_out;
  /* Split-K (0 / 1 = off; linear mode): the K loop of every output tile is divided among k_splits
   * CTAs (or CTA pairs); split s stores its fp32 partial tile into out + s * split_stride (plain
   * stores, so the result is deterministic) — `out` is an fp32 scratch [k_splits][M][ldo] and no
   * epilogue operand may be set; cd360_splitk_finish sums the slices in order and applies bias /
   * residual / conversion.  The number of slices actually written is
   * ceil(nkb / ceil(nkb / k_splits)) with nkb = ceil(k0/64) + ceil(k1/64) (cd360_splitk_slices).
   * For the small-M GEMMs of the training step (M = 256 tokens against 13-26 MB of weights: 10 CTAs
   * cannot pull the weights at HBM rate, 140 can) and weight gradients (K = 10^5 rows). */
  int32_t k_splits;
  int64_t split_stride; /* elements, >= M * ldo */
  /* tn != 0: out = A^T W with both operands stored contraction-outermost — a0 bf16 [k0, M] (row
   * stride lda0), w bf16 [k0, N] (row stride ldw), M, N, lda0, ldw multiples of 8; k0 arbitrary.
   * The weight gradients of the training step (dW = dY^T X over the token rows,
   * torch autograd of nn.Linear in the reference) without transposed copies of dY and X: the tiles
   * are consumed as MN-major tcgen05 operands.  Epilogue: bias / residual / act, or split-K. */
  int32_t tn;
  int64_t ldw;
} cd360_gemm_args;

int cd360_gemm_bf16(const cd360_gemm_args* args, cd360_stream_t stream);
/* Slices a split-K launch writes for a contraction of `k0 + k1` and a requested `k_splits`. */
int cd360_splitk_slices(int32_t k0, int32_t k1, int32_t k_splits);
/* Finish a split-K GEMM: out[m, n] = sum_s ws[s][m][n] (s ascending) + bias[n] + residual[m, n]
 * (bias / residual optional; out bf16 or fp32; row strides in elements; ws slices `split_stride`
 * elements apart with row stride ldw). */
int cd360_splitk_finish(const float* ws, int64_t ldw, int64_t split_stride, int32_t slices,
                        const float* bias, const void* residual, int64_t ldr, void* out, int64_t ldo,
                        int32_t out_fp32, int64_t M, int32_t N, cd360_stream_t stream);
/* Rows of the interleave block used by the GEGLU epilogue for a given N (= BN/2). */
int cd360_geglu_pack_block(int32_t n_total);

/* ---------------------------------------------------------------------------------------------
 * Attention, head dim 64, softmax(Q K^T / 8) V, bf16 in / fp32 softmax / bf16 out.
 * Replaces xformers.ops.memory_efficient_attention as called from
 * MemoryEfficientCrossAttention.forward (attention.py:393-418), including the head
 * split/merge permutes: Q/K/V are read in place from [B, n, heads*64] projections (row strides
 * ldq/ldk/ldv elements, so a fused QKV buffer works) and O is written as [B, nq, heads*64].
 * nkv need not be a multiple of the tile (77 text tokens): the tail is masked.
 * --------------------------------------------------------------------------------------------- */
int cd360_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                         int64_t ldv, void* o, int64_t ldo, int32_t batch, int32_t heads,
                         int32_t nq, int32_t nkv, cd360_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GroupNorm(32 groups) [+ SiLU] over NHWC bf16, fp32 statistics (GroupNorm32, util.py:309-311;
 * Normalize, attention.py:118).  Two launches inside: partial moments, then normalise(+SiLU).
 * The input may be the channel concatenation of two tensors (x0 [B*HW, c0] ‖ x1 [B*HW, c1]).
 * workspace: fp32, at least cd360_groupnorm_workspace_floats(B, HW) floats.
 * --------------------------------------------------------------------------------------------- */
int64_t cd360_groupnorm_workspace_floats(int32_t batch, int32_t hw);
int cd360_groupnorm_silu_bf16(const void* x0, int32_t c0, const void* x1, int32_t c1,
                              const float* gamma, const float* beta, void* out, float* workspace,
                              int32_t batch, int32_t hw, float eps, int32_t apply_silu,
                              cd360_stream_t stream);

/* LayerNorm over the channel dim of [rows, c] bf16 -> bf16 (nn.LayerNorm, attention.py:531-533). */
int cd360_layernorm_bf16(const void* x, const float* gamma, const float* beta, void* out,
                         int32_t rows, int32_t c, float eps, cd360_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Small-M linear for the embedding path: out[b, n] = act_out( sum_k act_in(x[b,k]) W[n,k] + bias[n] )
 * (+ add[b, n]).  x, out, add fp32; W bf16.  time_embed / label_emb (openaimodel.py:679-713,
 * 1026-1031) and ResBlock.emb_layers (openaimodel.py:307-313) for all blocks at once.
 * --------------------------------------------------------------------------------------------- */
int cd360_small_linear(const float* x, const void* w, const float* bias, const float* add,
                       float* out, int32_t batch, int32_t n, int32_t k, int32_t act_in,
                       int32_t act_out, cd360_stream_t stream);

/* timestep_embedding(t, dim) = [cos(t f_k) ‖ sin(t f_k)], f_k = exp(-ln(1e4) k / (dim/2))
 * (util.py:206-230).  t fp32 [batch]; out fp32 [batch, dim]. */
int cd360_timestep_embedding(const float* t, float* out, int32_t batch, int32_t dim,
                             cd360_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Layout / resampling helpers around the convolutions.
 * --------------------------------------------------------------------------------------------- */
/* x fp32 NCHW [src_batch, Cin, H, W] (row b reads image b % src_batch: the CFG replication
 * torch.cat([x]*rows), guiders.py:133, folded into the load) * scale[b] -> im2col bf16 [B*H*W, kpad] for the 3x3 pad-1 input conv
 * (openaimodel.py:720); k = (ky*3+kx)*Cin + c, zero padded to kpad.  scale may be NULL. */
int cd360_im2col3x3_nchw_f32(const float* x, const float* scale, void* out, int32_t batch,
                             int32_t src_batch, int32_t cin, int32_t h, int32_t w, int32_t kpad,
                             cd360_stream_t stream);
/* NHWC bf16 [B,H,W,C] -> im2col bf16 [B*(H/2)*(W/2), 9*C] of a stride-2 pad-1 3x3 conv
 * (Downsample, openaimodel.py:215-222). */
int cd360_im2col3x3_s2_bf16(const void* x, void* out, int32_t batch, int32_t h, int32_t w,
                            int32_t c, cd360_stream_t stream);
/* nearest x2 upsample, NHWC bf16 (Upsample.forward, openaimodel.py:161). */
int cd360_upsample_nearest2x_bf16(const void* x, void* out, int32_t batch, int32_t h, int32_t w,
                                  int32_t c, cd360_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Sampler-side elementwise math (fp32, NCHW latents).
 * --------------------------------------------------------------------------------------------- */
/* One guided Euler step, fused: eps is the UNet output for the G guidance rows of each image,
 * fp32 [G*N, hw, 4] (NHWC); x fp32 NCHW [N,4,h,w] updated in place.
 *   D_g   = x - sigma_q * eps_g   (EpsScaling c_out=-sigma_q, c_skip=1, denoiser.py:44; sigma_q is
 *                                  sigma after DiscreteDenoiser's table quantisation, :65-73)
 *   G==3: D = D_u + s (D_c - D_ic) + s_im (D_ic - D_u)   (ScheduledCFGImgTextRef, guiders.py:111-114)
 *   G==2: D = D_u + s (D_c - D_u)                        (VanillaCFGImgRef, guiders.py:147-150)
 *   x    += (x - D) / sigma * (sigma_next - sigma)       (to_d + Euler, sampling.py:103-106)
 * denoised_out (optional) receives D. */
int cd360_cfg_euler_step(float* x, const float* eps, float* denoised_out, int32_t n_img,
                         int32_t guidance_rows, int32_t hw, float sigma_q, float sigma,
                         float sigma_next, float scale, float scale_im, cd360_stream_t stream);

/* Same, with (sigma_q, sigma, sigma_next) read from device memory (fp32 [3]) so that one captured
 * CUDA graph replays for every step of the schedule. */
int cd360_cfg_euler_step_dev(float* x, const float* eps, float* denoised_out, int32_t n_img,
                             int32_t guidance_rows, int32_t hw, const float* sigmas3, float scale,
                             float scale_im, cd360_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * FeatureNeRF (sgm/modules/nerfsd_pytorch3d.py, sgm/modules/utils_cameraray.py).
 * Cameras are packed fp32 [b, n+1, 16] = R(9, row-major, PyTorch3D row-vector convention
 * X_cam = X_world R + T) | T(3) | focal(2) | principal point(2); index 0 = target camera.
 * --------------------------------------------------------------------------------------------- */
/* Per (b, ray, sample, view) geometry: projects the target-ray sample points into every reference
 * view and emits
 *   pe      bf16 [b, n, hw, d, kpe]   the 198 positional features of plane_coefs' input
 *                                     (nerfsd_pytorch3d.py:102-134), zero padded to kpe
 *   gidx    int32 [b, n, hw, d, 4]    flat pixel indices of the 4 bilinear corners (or -1)
 *   gwgt    fp32  [b, n, hw, d, 4]    bilinear weights (grid_sample align_corners=True, zeros
 *                                     padding, grid = clip(nan_to_num(-ndc), +-1.2); :79-98)
 *   vlogit  fp32  [b, n, hw, d]       geometry part of the nviews logit (:139-151), without the
 *                                     gathered-feature term
 * depths: fp32 [hw, d] sample depths along each target ray (Raymarcher, :308-330);
 * xy: fp32 [hw, 2] NDC ray positions (get_patch_raybundle, utils_cameraray.py:103-158).
 * w_nv_geo: fp32 [198] the non-feature columns of nviews.weight; b_nv: its bias, fp32 [1] in DEVICE
 * memory (NULL = 0) so that a captured CUDA graph follows optimiser updates of the bias. */
int cd360_nerf_points(const float* cams, const float* xy, const float* depths,
                      const float* w_nv_geo, const float* b_nv, void* pe, int32_t* gidx, float* gwgt,
                      float* vlogit, int32_t b, int32_t n, int32_t res, int32_t d, int32_t kpe,
                      cd360_stream_t stream);
/* Combine: for every (b, ray, sample): h_v = SiLU(hpre[b,v,p,:] + bilinear(G[b,v], gidx,gwgt)[:c]);
 * logit_v = vlogit + bilinear(G[b,v])[c]; a = softmax_v(logit); s = sum_v a_v h_v.
 * G: bf16 [b, n, hw, ldg] (first Linear hoisted through the gather: G = xref W1f^T ‖ xref w_nv_f),
 * hpre: bf16 [b, n, hw*d, c] (= pe W1p^T + b1).  Outputs s bf16 [b, hw*d, c] and the view softmax
 * fp32 [b, n, hw*d] (plane_features_attn, :139-155). */
int cd360_nerf_combine(const void* g, int64_t ldg, const void* hpre, const int32_t* gidx,
                       const float* gwgt, const float* vlogit, void* s, float* view_softmax,
                       int32_t b, int32_t n, int32_t hw, int32_t d, int32_t c,
                       cd360_stream_t stream);
/* Volume rendering (VolRender.forward, nerfsd_pytorch3d.py:170-231; trunc_exp attention.py:192-208;
 * sigmoid on rgb attention.py:594).  feats bf16 [b, hw, d, c]; raw fp32 [b, hw, d, 4] =
 * (rgb_raw 3, sigma_raw 1); dists fp32 [hw, d].  Outputs rendered bf16 [b, hw, c],
 * fg fp32 [b, hw], alphas fp32 [b, hw, d], rgb fp32 [b, hw, 3]. */
int cd360_nerf_volrender(const void* feats, const float* raw, const float* dists, void* rendered,
                         float* fg, float* alphas, float* rgb, int32_t b, int32_t hw, int32_t d,
                         int32_t c, cd360_stream_t stream);

/* Reference padding masks (FeatureNeRFEncoding.forward step 1, nerfsd_pytorch3d.py:61-70; supplied by every
 * training batch, data_co3d.py:485): out[(bn, y, x), :] = x[(bn, y, x), :] * mask[bn, floor(y*mh/res),
 * floor(x*mw/res)] — F.interpolate(mask_ref, [res, res], mode="nearest") fused with the multiply.
 * x / out bf16 [bn*res*res, c] (may alias), mask fp32 [bn, mh, mw]. */
int cd360_nerf_mask_ref(const void* x, const float* mask, void* out, int64_t bn, int32_t res,
                        int32_t mh, int32_t mw, int32_t c, cd360_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Text conditioner: FrozenCLIPEmbedder (HF CLIPTextModel, CLIP-L) and
 * FrozenOpenCLIPEmbedder (open_clip ViT-bigG-14 text tower), sgm/modules/encoders/modules.py:377-517,
 * 622-772.  Projections / MLPs use cd360_gemm_bf16 (CD360_ACT_QUICK_GELU / CD360_ACT_GELU epilogues),
 * LayerNorm cd360_layernorm_bf16, pooled projection cd360_small_linear.
 * --------------------------------------------------------------------------------------------- */
/* Token + positional embedding (modules.py:498-501 `text_model.embeddings(input_ids=tokens)`;
 * :716-728 `token_embedding(text) + positional_embedding`): out[b*ctx + t, :] = bf16(tok_emb[ids[b*ctx+t], :]
 * + pos_emb[t, :]).  ids int32 [rows], tok_emb fp32 [vocab, w] (the trainable `<new1>` row stays an fp32
 * master), pos_emb fp32 [ctx, w], out bf16 [rows, w]; rows % ctx == 0, w % 4 == 0. */
int cd360_embed_tokens(const int32_t* ids, const float* tok_emb, const float* pos_emb, void* out,
                       int32_t rows, int32_t ctx, int32_t w, int32_t vocab, cd360_stream_t stream);
/* Causal self-attention of the text towers (additive -inf mask above the diagonal, modules.py:447-453;
 * open_clip `attn_mask`), head dim 64, scale 1/8, n <= 128 tokens per sequence.  q/k/v/out are bf16
 * row-major views [batch*n, >= heads*64] with row strides ld* (slices of a fused QKV buffer work). */
int cd360_attention_causal_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                int64_t ldv, void* out, int64_t ldo, int32_t batch, int32_t heads,
                                int32_t n, cd360_stream_t stream);
/* Row gather for the end-of-text pooling (modules.py:737-743 `x[arange(B), text.argmax(-1)]`):
 * out[r, :] = fp32(x[idx[r], :]); x bf16 [src_rows, c] row stride ldx, idx int32 [nrows]. */
int cd360_gather_rows_bf16_f32(const void* x, int64_t ldx, const int32_t* idx, float* out, int32_t nrows,
                               int32_t c, int64_t src_rows, cd360_stream_t stream);

/* Utility: fp32 -> bf16 and bf16 -> fp32 contiguous conversion (weight prepack, I/O). */
int cd360_cast_f32_to_bf16(const float* x, void* out, int64_t n, cd360_stream_t stream);
int cd360_cast_bf16_to_f32(const void* x, float* out, int64_t n, cd360_stream_t stream);
/* NHWC bf16/fp32 [B, hw, C] -> NCHW fp32 [B, C, hw] (module outputs at the sgm boundary). */
int cd360_nhwc_to_nchw_f32(const void* x, int32_t x_is_fp32, float* out, int32_t batch,
                           int32_t hw, int32_t c, cd360_stream_t stream);

/* NCHW fp32 [B, C, hw] -> NHWC bf16 [B, hw, C] (module inputs at the sgm boundary). */
int cd360_nchw_f32_to_nhwc_bf16(const float* x, void* out, int32_t batch, int32_t hw, int32_t c,
                                cd360_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * VAE decode after the sampling loop: DiffusionEngine.decode_first_stage
 * (sgm/models/diffusion.py:207-212) -> AutoencoderKL.decode (sgm/models/autoencoder.py:313-316) ->
 * Decoder.forward (sgm/modules/diffusionmodules/model.py:715-757).  Convolutions, GroupNorm+swish,
 * nearest upsampling and the 1x1 q/k/v/proj_out reuse the entry points above; these two are new.
 * --------------------------------------------------------------------------------------------- */
/* out[b, o, p] = bias[o] + scale * sum_c w[o, c] x[b, c, p]; fp32 NCHW, cin, cout <= 8: the 1x1
 * post_quant_conv (autoencoder.py:314) with decode_first_stage's 1/scale_factor (diffusion.py:209)
 * folded into the weights.  bias may be NULL. */
int cd360_pointwise_conv_nchw_f32(const float* x, const float* w, const float* bias, float* out,
                                  int32_t batch, int32_t cin, int32_t cout, int64_t hw, float scale,
                                  cd360_stream_t stream);
/* out[r, :] = softmax(scale * s[r, :]) — fp32 scores [rows, n] (row stride lds) -> bf16 weights
 * (row stride ldo); n % 4 == 0.  The softmax of MemoryEfficientAttnBlock (model.py:231-266: ONE head
 * of width C over all H*W pixels, scale C^-0.5), between the score GEMM q k^T (fp32 out) and the
 * P v GEMM of cd360_gemm_bf16. */
int cd360_softmax_rows_f32_bf16(const float* s, int64_t lds, void* out, int64_t ldo, int64_t rows,
                                int32_t n, float scale, cd360_stream_t stream);

/* =============================================================================================
 * Training step: backward of the path above towards the pose weights
 * (trainkeys 'pose', sgm/models/diffusion.py:139-144), the loss of
 * StandardDiffusionLossImgRef (sgm/modules/diffusionmodules/loss.py:140-216) and AdamW.
 * The reference obtains all of these from torch.autograd; here each is an explicit kernel.  The
 * data gradients of every Linear / conv reuse cd360_gemm_bf16 with transposed / tap-flipped
 * weight packs (dX = dY W), weight gradients are cd360_gemm_bf16 over operands transposed by
 * cd360_transpose_to_bf16 (dW = dY^T X, fp32 out).
 * ============================================================================================= */

/* Backward of cd360_attention_bf16 (autograd of xformers memory_efficient_attention,
 * attention.py:406).  o = forward output, dout = its gradient, all [B, n, heads*64] bf16 with row
 * strides in elements.  dq always; dk/dv both or neither (NULL for text cross-attention, whose K/V
 * are projections of the constant context).  lse / dsum: fp32 scratch [batch, heads, nq] (opaque per-query
 * statistics; only cd360_attention_bwd_kv_split_bf16 may consume them).  Runs on tcgen05 / TMEM
 * (csrc/attention_bwd_tcgen05.cu: 128 x 128 tiles, S / dP by SS MMAs, P / dS handed to the second MMA
 * through TMEM). */
int cd360_attention_bwd_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                             int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo,
                             void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                             float* lse, float* dsum, int32_t batch, int32_t heads, int32_t nq,
                             int32_t nkv, cd360_stream_t stream);

/* dK / dV of an attention with FEW keys and VERY MANY queries — reference_attn's attn2 over the
 * hw*24 ray samples against the 77 text tokens (attention.py:571-598), whose gradient the conditioner
 * needs (sgm/models/diffusion.py:343-356).  The query tiles are divided among `nsplit` CTAs per
 * (key tile, head, batch); partial sums meet in kv_acc (fp32 [2][batch*nkv][heads*64], ZEROED by the
 * caller) and are converted to bf16 dk / dv.  lse / dsum: the per-query statistics written by a
 * preceding cd360_attention_bwd_bf16 call on the same operands (its dk / dv NULL). */
int cd360_attention_bwd_kv_split_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk,
                                      const void* v, int64_t ldv, const void* dout, int64_t lddo,
                                      const float* lse, const float* dsum, float* kv_acc,
                                      void* dk, int64_t lddk, void* dv, int64_t lddv,
                                      int32_t batch, int32_t heads, int32_t nq, int32_t nkv,
                                      int32_t nsplit, cd360_stream_t stream);

/* LayerNorm backward w.r.t. the input (nn.LayerNorm, attention.py:531-533; gamma/beta are frozen):
 * dx = dLN(x)^T dy [+ add].  x, dy, add, dx bf16 [rows, c]; c <= 1280. */
int cd360_layernorm_bwd_bf16(const void* x, const float* gamma, const void* dy, const void* add,
                             void* dx, int32_t rows, int32_t c, float eps, cd360_stream_t stream);

/* GroupNorm(32)[+SiLU] backward w.r.t. the input (GroupNorm32 + SiLU, openaimodel.py:280-283;
 * Normalize, attention.py:118).  Input = virtual concat [x0 | x1] as in the forward; dy bf16
 * [B*HW, c0+c1]; optional add0/add1 (bf16, row strides ld_add0/ld_add1) are accumulated into the
 * result (the gradient arriving through the ResBlock skip path); outputs dx0 [B*HW, c0],
 * dx1 [B*HW, c1].  workspace: cd360_groupnorm_bwd_workspace_floats(B) floats. */
int64_t cd360_groupnorm_bwd_workspace_floats(int32_t batch);
int cd360_groupnorm_silu_bwd_bf16(const void* x0, int32_t c0, const void* x1, int32_t c1,
                                  const float* gamma, const float* beta, const void* dy,
                                  const void* add0, int64_t ld_add0, const void* add1,
                                  int64_t ld_add1, void* dx0, void* dx1, float* workspace,
                                  int32_t batch, int32_t hw, float eps, int32_t apply_silu,
                                  cd360_stream_t stream);

/* GEGLU backward (attention.py:94-96).  raw bf16 [rows, 2f] = pre-activation [a | gate] interleaved
 * in blocks of `block` columns (block = f: chunk(2) layout; block = cd360_geglu_pack_block(2f): the
 * packed-weight layout), dh bf16 [rows, f] -> draw bf16 [rows, 2f] in the same column order. */
int cd360_geglu_bwd_bf16(const void* raw, const void* dh, void* draw, int64_t rows, int32_t f,
                         int32_t block, cd360_stream_t stream);

/* GEGLU forward from a KEPT pre-activation (training: raw is saved for cd360_geglu_bwd_bf16 instead of
 * being recomputed): raw bf16 [rows, 2f] in the same block layout -> h bf16 [rows, f] = a * gelu(gate)
 * (attention.py:94-96, exact erf). */
int cd360_geglu_fwd_bf16(const void* raw, void* h, int64_t rows, int32_t f, int32_t block,
                         cd360_stream_t stream);

/* out = a + b, bf16, n % 8 == 0 (sum of the gradients arriving at a skip connection,
 * openaimodel.py:1074). */
int cd360_add_bf16(const void* a, const void* b, void* out, int64_t n, cd360_stream_t stream);

/* out = dy * silu'(pre), fp32, n elements: backward of the SiLU inside time_embed / label_emb /
 * ResBlock.emb_layers (openaimodel.py:679-713, 270-276) on the way to dL/d(vector) — the gradient the
 * reference's conditioner needs for its trainable token rows (sgm/models/diffusion.py:343-356). */
int cd360_silu_bwd_f32(const float* pre, const float* dy, float* out, int64_t n, cd360_stream_t stream);

/* in [rows, cols] (bf16, or fp32 if in_is_fp32; row stride ld_in) -> out bf16 [cols, ld_out] with
 * out[c][r] = in[r][c], zero for r in [rows, ld_out): the K-major operand of a weight-gradient GEMM. */
int cd360_transpose_to_bf16(const void* in, int32_t in_is_fp32, int64_t ld_in, void* out,
                            int64_t ld_out, int32_t rows, int32_t cols, cd360_stream_t stream);

/* out[c] += sum_r x[r, c] (bias gradients); x bf16 [rows, c] row stride ld; out fp32, zeroed by the caller. */
int cd360_colsum_bf16(const void* x, int64_t ld, float* out, int64_t rows, int32_t c,
                      cd360_stream_t stream);

/* Backward of cd360_im2col3x3_s2_bf16 (Downsample, openaimodel.py:215-222):
 * dcol bf16 [B*(H/2)*(W/2), 9*C] -> dx bf16 [B*H*W, C]. */
int cd360_col2im3x3_s2_bf16(const void* dcol, void* dx, int32_t batch, int32_t h, int32_t w,
                            int32_t c, cd360_stream_t stream);
/* Backward of cd360_upsample_nearest2x_bf16: g bf16 [B, 2h, 2w, C] -> dx [B, h, w, C]. */
int cd360_upsample_nearest2x_bwd_bf16(const void* g, void* dx, int32_t batch, int32_t h, int32_t w,
                                      int32_t c, cd360_stream_t stream);

/* Backward of cd360_nerf_volrender (VolRender, nerfsd_pytorch3d.py:170-231; trunc_exp backward
 * attention.py:201-205).  d_rendered bf16 [b*hw, c]; dfg [b,hw], dalphas [b,hw,d], drgb [b,hw,3]
 * fp32 or NULL.  Outputs dfeats bf16 [b,hw,d,c] and draw bf16 [b,hw,d,8] = gradient of the raw
 * decoder outputs (rgb 3, sigma 1, 4 zero columns: the K-padded GEMM operand). */
int cd360_nerf_volrender_bwd(const void* feats, const float* raw, const float* dists,
                             const void* d_rendered, const float* dfg, const float* dalphas,
                             const float* drgb, void* dfeats, void* draw, int32_t b, int32_t hw,
                             int32_t d, int32_t c, cd360_stream_t stream);
/* Backward of cd360_nerf_combine.  ds bf16 [b, hw*d, c] -> dhpre bf16 [b, n, hw*d, c], dlogit fp32
 * [b, n, hw*d], and dg fp32 [b, n, hw, ldg] (zeroed by the caller; bilinear scatter by atomics:
 * columns [0,c) gradient of the hoisted first Linear's output, column c of the nviews feature term). */
int cd360_nerf_combine_bwd(const void* g, int64_t ldg, const void* hpre, const int32_t* gidx,
                           const float* gwgt, const float* vlogit, const void* ds, void* dhpre,
                           float* dlogit, float* dg, int32_t b, int32_t n, int32_t hw, int32_t d,
                           int32_t c, cd360_stream_t stream);
/* Gradient of the geometry columns of nviews.weight (nerfsd_pytorch3d.py:139-151): dw fp32 [198],
 * zeroed by the caller; dlogit fp32 [b, n, pts]. */
int cd360_nerf_nviews_geo_bwd(const float* cams, const float* dlogit, float* dw, int32_t b, int32_t n,
                              int64_t pts, cd360_stream_t stream);

/* Denoising loss and its gradient (loss.py:173-181 'l2', EpsWeighting, EpsScaling):
 * eps fp32 tokens [b*hw, 4]; x_noisy, target fp32 NCHW [b,4,hw]; sigma [b]; mask fp32 [b,hw] or NULL.
 * loss[b]; mask_sum[b] (optional); deps bf16 [b*hw, ldd] = d(coef * sum_b loss_b)/d eps, zero padded. */
int cd360_diffusion_loss(const float* eps, const float* x_noisy, const float* target,
                         const float* sigma, const float* mask, float coef, float* loss,
                         float* mask_sum, void* deps, int32_t batch, int32_t hw, int32_t ldd,
                         cd360_stream_t stream);
/* FeatureNeRF supervision of one pose block and its gradients (loss.py:183-206,
 * diffusion.py:221-236).  loss3 fp32 [b, 3] = (fg, bg, rgb); gradients pre-multiplied by the
 * per-image weights wfg / wbg / wrgb.  rgb NULL = no rgb term. */
int cd360_nerf_aux_loss(const float* fg, const float* alphas, const float* rgb, const float* op,
                        const float* mask_s, const float* tgt, const float* mask_sum,
                        const float* wfg, const float* wbg, const float* wrgb, float* loss3,
                        float* dfg, float* dalphas, float* drgb, int32_t batch, int32_t hw, int32_t d,
                        cd360_stream_t stream);
/* F.interpolate(mode='bilinear', antialias=True) of the supervision maps (loss.py:186,199-200):
 * in fp32 [planes, ih, iw] -> out [planes, oh, ow] = out_scale * resized + out_shift. */
int cd360_resize_bilinear_aa(const float* in, float* out, int32_t planes, int32_t ih, int32_t iw,
                             int32_t oh, int32_t ow, float out_scale, float out_shift,
                             cd360_stream_t stream);
/* AdamW on flat fp32 buffers (torch.optim.AdamW, the default optimizer_config; diffusion.py:310-373).
 * step >= 1 is the 1-based step count; g is multiplied by grad_scale first. */
int cd360_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                     cd360_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CD360_H_ */
