/* npvp_b200 C-ABI: sm_100a kernels for the NPVP inference hot path.
 *
 * Drop-in boundary: the reference (XiYe20/NPVP) is pure PyTorch and owns no
 * native code, so there is no existing FFI to mirror; each entry point below replaces the
 * eager torch op sequence cited next to it (file:line into the reference tree).  The Python
 * modules in npvp_b200/ (same class names / signatures as models/Predictor.py and
 * models/ResNetAutoEncoder.py) bind these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - every function returns NPVP_OK (0) or a negative error code; the message is available
 *     from npvp_last_error() (thread-local);  no exceptions cross the boundary;
 *   - token layout: activations are channels-last, a "frame" is 64 tokens (8x8 grid) x C;
 *     16-bit buffers are GEMM/conv operands, fp32 buffers carry residual streams and statistics;
 *   - "bf16" in a parameter name means "16-bit operand buffer": entry points with an `fp16` flag (and the
 *     `fp16` field of npvp_epilogue_t) read/write IEEE half instead of bfloat16 when it is 1.  The autoencoder
 *     runs in half (its rounding errors land directly in pixels), the predictor in bfloat16 (range safety).
 */
#ifndef NPVP_B200_H
#define NPVP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPVP_OK 0
#define NPVP_ERR_INVALID (-1)
#define NPVP_ERR_CUDA (-2)

#define NPVP_ACT_NONE 0
#define NPVP_ACT_RELU 1
#define NPVP_ACT_GELU 2 /* exact erf GELU, nn.GELU() default (models/VidHRFormer.py:73,337) */
#define NPVP_ACT_TANH 3
#define NPVP_ACT_SIGMOID 4

#define NPVP_GEMM_AUTO 0
#define NPVP_GEMM_TCGEN05 1 /* TMA-fed tcgen05.mma, TMEM accumulators */
#define NPVP_GEMM_SIMT 2    /* CUDA-core reference path for debugging / odd shapes */
#define NPVP_GEMM_TCGEN05_2CTA 4 /* cluster of 2 CTAs, tcgen05.mma.cta_group::2, 256x256 tiles */

#define NPVP_PAD_ZERO 0
#define NPVP_PAD_REFLECT 1
#define NPVP_PAD_REPLICATE 2

#define NPVP_ATTN_SPATIAL_WINDOW 0 /* 4x4 windows on the 8x8 grid, L = 16 */
#define NPVP_ATTN_TEMPORAL 1       /* per-pixel sequences over time, Lq = Tq, Lk = Tk */

/* Fused GEMM epilogue:  v = acc + bias[n];  v = act(v);  v *= alpha;  v += res1[m,n];  v += res2[m,n];
 * v = post_relu ? max(v,0) : v;  stored to out_f32 and/or out_bf16 (row stride ld_out elements). */
typedef struct npvp_epilogue {
  const void* bias; /* fp32 [N] or NULL */
  const void* res1; /* [M, ld_res] or NULL */
  const void* res2; /* [M, ld_res] or NULL */
  void* out_f32;    /* fp32 [M, ld_out] or NULL */
  void* out_bf16;   /* bf16 [M, ld_out] or NULL */
  float alpha;
  int32_t act;
  int32_t res1_bf16; /* 1: res1 is bf16, 0: fp32 */
  int32_t res2_bf16;
  int32_t post_relu;
  int32_t fp16;      /* 16-bit type of the operands A, W: 0 = bfloat16, 1 = IEEE half */
  int32_t out16;     /* 16-bit type of out_bf16 and of 16-bit residuals: 0 = same as the operands, 1 = IEEE half, 2 = bfloat16 */
  int64_t ld_out;
  int64_t ld_res;
  /* optional (NULL = off): per-frame partial statistics of the stored values, a frame being 64 consecutive rows.
   * fp32 [M/64][P = 4 * N/256][2] = (sum, sum of squares) over one 32-row x 128-column block each; reduce them with
   * npvp_ffn_stats_finalize.  Requires M % 64 == 0, N % 256 == 0, 16-bit output only, act NONE, no residuals
   * (the conv-FFN's fc1: LayerNorm((Ch,8,8)) statistics for free instead of a second pass over h1). */
  float* frame_stats;
} npvp_epilogue_t;

const char* npvp_last_error(void);
int npvp_version(void);
/* number of kernels launched through this library since the last reset (bench accounting) */
int64_t npvp_launch_count(void);
void npvp_reset_launch_count(void);
/* library-wide switches: "gemm_2cta": 1 = N >= 256 GEMMs of the default back-end always use the 2-CTA cluster kernel,
 * 0 = never, -1 (default) = when K >= 1024 (where the main loop dominates and the halved operand traffic pays);
 * "gemm_epi_direct": 1 = 16-bit outputs without residuals are stored straight from the accumulator layout with 32-byte
 * sector stores, 0 (default) = transposed through shared memory for row-coalesced stores (A/B switch, identical results) */
int npvp_set_option(const char* name, int value);

/* ---- dense contractions -------------------------------------------------------------------
 * D[M,N] = A[M,K] (bf16, row stride lda) x W[N,K]^T (bf16, row stride ldw), fp32 accumulate.
 * Replaces every nn.Linear / 1x1 conv / MHA in- and out-projection on the path
 * (models/VidHRFormer.py:104,111,221,225,239,298,380,387; models/submodules.py:148-162,396-403)
 * and, through npvp_conv_gemm_bf16, the 3x3 / strided / transposed convs of the autoencoder
 * (models/ResNetAutoEncoder.py:75-87,169-183,241,254; models/submodules.py:25). */
int npvp_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                   const npvp_epilogue_t* ep, int backend, void* stream);
/* Convolution as an IMPLICIT GEMM on the same tcgen05 kernel: rows = output pixels (f,oy,ox), K = (ky,kx,ci); the A operand
 * is gathered from the channels-last activations by producer warps (cp.async into the 128B-swizzled smem layout), so no
 * im2col buffer is written or read.  x 16-bit [frames,H,W,C] (phase_major=1: [frames,H/2,W/2,4,C]); Wt 16-bit [N, KH*KW*C];
 * C must be 32 or a multiple of 64.  Replaces the 3x3 / stride-2 / transposed convolutions of the autoencoder and the
 * event encoder (models/ResNetAutoEncoder.py:75-87,169-183,241,254; models/submodules.py:25,376). */
int npvp_conv_gemm_bf16(const void* x, int64_t frames, int H, int W, int C, int KH, int KW, int stride, int pad,
                        int pad_mode, int Ho, int Wo, int phase_major, const void* Wt, int64_t ldw, int64_t N,
                        const npvp_epilogue_t* ep, void* stream);
/* (r02) When stride == 1, the padding is zero padding of (K-1)/2, C % 64 == 0 and 128-pixel tiles are whole rows or whole
 * frames (W <= 128, 128 % W == 0, H*W % 128 == 0 or 128 % (H*W) == 0), npvp_conv_gemm_bf16 loads the A operand as SHIFTED
 * WINDOWS of a 4-D tensor map (C, W, H, frames) with cp.async.bulk.tensor.4d - the TMA unit zero-fills the padding - and
 * runs without gather warps; npvp_set_option("conv_tma", 0) switches back to the gather (A/B).
 *
 * Transposed convolution nn.ConvTranspose2d(Cin, Cout, 3, stride 2, padding 1, output_padding 1) + folded BatchNorm + ReLU
 * (models/ResNetAutoEncoder.py:169-183) as ONE implicit GEMM over the 2x2 input neighbourhood:
 *   out[f, 2y+py, 2x+px, co] = act( bias[(q,co)] + sum_{dy,dx,ci} x[f, y+dy, x+dx, ci] * Wt[(q,co), (dy,dx,ci)] ),
 * output phases ordered q = (py,px) = (0,0), (0,1), (1,1), (1,0) so that every tap feeds a contiguous column range: only the
 * 9 live (phase, tap) blocks of the 16 are loaded and multiplied.  x [frames,H,W,Cin] and out [frames,2H,2W,Cout] are plain
 * channels-last 16-bit tensors (A tiles by TMA as above, pixel-shuffled stores in the epilogue).  Wt 16-bit [4*Cout, 4*Cin],
 * bias fp32 [4*Cout]; Cin % 64 == 0, Cout % 32 == 0, W a power of two <= 128; ep: 16-bit output only, no residuals. */
int npvp_convt_gemm_bf16(const void* x, int64_t frames, int H, int W, int Cin, const void* Wt, int64_t ldw, int Cout,
                         const npvp_epilogue_t* ep, void* stream);
/* fp32 CUDA-core GEMM for the tiny, precision-critical NRMLP (models/submodules.py:299-314). */
int npvp_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                  const float* bias, int act, float* out, int64_t ldo, void* stream);

/* ---- predictor: positional code, per-token / per-frame normalisation --------------------------
 * Fourier features [cos(2 pi x B^T), sin(2 pi x B^T)]  (NRMLP.gaussian_mapping, submodules.py:317-327).
 * coor fp32 [rows,3], B fp32 [half,3], out fp32 [rows, 2*half]. */
int npvp_fourier_features(const float* coor, const float* B, float* out, int64_t rows, int half, void* stream);

/* Fused  a = LayerNorm_C(x) ; u = a + qe ; fused = GroupNorm1(u over the frame) * (1+gamma) + beta
 * (VidHRFormer.py:87-88,95-96,210-212,218-219,229,236 + PosFeatFuser submodules.py:432-454).
 * x fp32 [n_clips*T, 64, 512]; ln_w/ln_b fp32 [512] or NULL (no LayerNorm: a = x);
 * qe fp32 [n_clips, 64, 512] or NULL; beta fp32 [pos_frames,64,512]; gamma fp32 [pos_frames,64,512] or NULL;
 * pos_frames = T (or 0): one set of timestamps shared by the batch, like the reference's per-module coordinates
 * (Predictor.py:352-359); pos_frames = n_clips * T: every clip has its own timestamps (a batch mixing prediction,
 * interpolation and arbitrary continuous-time queries);
 * out_ln (bf16, optional) receives a; out_fused (bf16, optional) receives fused. */
int npvp_ln_posfuse(const float* x, const float* ln_w, const float* ln_b, const float* qe, const float* beta,
                    const float* gamma, void* out_ln_bf16, void* out_fused_bf16, int64_t n_clips, int64_t T,
                    int64_t pos_frames, void* stream);
/* LayerNorm over C=512 per token (VidHRFormer.py:91,110,214,224,243; final norm :48,:151 with relu=1 for :159).
 * Outputs optional fp32 and/or bf16. */
int npvp_layernorm_rows(const float* x, const float* w, const float* b, float* out_f32, void* out_bf16,
                        int64_t rows, int relu, int fp16, void* stream);
/* Deferred residual add: a residual branch's GEMM leaves its output `delta` in bf16 [rows,512] instead of read-modify-writing
 * the fp32 stream in its (latency-sensitive) epilogue; the LayerNorm kernel that reads the stream next performs
 * x += delta (written back, fp32) before normalising.  Same arguments as the functions above plus x in-out and delta. */
int npvp_add_layernorm_rows(float* x, const void* delta_bf16, const float* w, const float* b, float* out_f32, void* out_bf16,
                            int64_t rows, int relu, int fp16, void* stream);
int npvp_add_ln_posfuse(float* x, const void* delta_bf16, const float* ln_w, const float* ln_b, const float* qe, const float* beta,
                        const float* gamma, void* out_ln_bf16, void* out_fused_bf16, int64_t n_clips, int64_t T, int64_t pos_frames,
                        void* stream);
/* y += GELU(LayerNorm_(C,8,8)(h))  - MlpDWBN norm3 + act3 + the block's residual add
 * (VidHRFormer.py:388-389 with :91/:214/:243).  h [frames,64,512] fp32 (h_is_bf16 = 0) or bf16 (1: the fc2 GEMM's 16-bit
 * output; h is normalised right here, so its rounding is harmless); w,b fp32 [64,512] (hw-major). */
int npvp_frame_ln_gelu_residual(const void* h, int h_is_bf16, const float* w_hwc, const float* b_hwc, float* y, int64_t frames,
                                void* stream);
/* The same, fused with the consumer that follows it in every block: after y is updated in registers the kernel also runs
 * npvp_ln_posfuse on the new y (VidHRFormer.py:91 -> :95-96, :214 -> :218-219, :243 -> next layer's :210-212). */
int npvp_frame_ln_gelu_residual_posfuse(const void* h, int h_is_bf16, const float* w_hwc, const float* b_hwc, float* y, const float* ln_w,
                                        const float* ln_b, const float* qe, const float* beta, const float* gamma,
                                        void* out_ln_bf16, void* out_fused_bf16, int64_t n_clips, int64_t T, int64_t pos_frames,
                                        void* stream);
/* mean over time of the memory (Predictor.py:346): mem fp32 [n,T,64*512] -> evt fp32 [n,64*512]. */
int npvp_temporal_mean(const float* mem, float* evt, int64_t n_clips, int64_t T, int64_t frame_elems, void* stream);

/* ---- predictor: conv-FFN middle  (MlpDWBN norm1/act1/dw3x3/norm2/act2, VidHRFormer.py:381-385) ----
 * step 1: per-frame (sum, sumsq) of h1 bf16 [frames,64,Ch] -> stats fp32 [frames,2] = (mean, rstd) */
int npvp_ffn_frame_stats(const void* h_bf16, float* stats, int64_t frames, int64_t Ch, void* stream);
/* step 1 fused into the producing GEMM (npvp_epilogue_t.frame_stats): partial fp32 [frames,P,2] -> stats fp32 [frames,2] */
int npvp_ffn_stats_finalize(const float* partial, int64_t P, float* stats, int64_t frames, int64_t elems_per_frame, void* stream);
/* step 2: y = dw3x3(GELU(LN1(h1))) + b; also per-(frame,chunk) partial (sum,sumsq) of y.
 * n1w/n1b fp32 [64,Ch] (hw-major elementwise affine), dw_w fp32 [9,Ch], dw_b fp32 [Ch];
 * y bf16 [frames,64,Ch]; partial fp32 [frames, Ch/128, 2] (Ch a multiple of 128). */
int npvp_ffn_dwconv(const void* h_bf16, const float* stats1, const float* n1w, const float* n1b,
                    const float* dw_w, const float* dw_b, void* y_bf16, float* partial2, int64_t frames,
                    int64_t Ch, void* stream);
/* step 3: out = GELU(LN2(y)) bf16, statistics from the partials of step 2. */
int npvp_ffn_norm2(const void* y_bf16, const float* partial2, const float* n2w, const float* n2b, void* out_bf16,
                   int64_t frames, int64_t Ch, void* stream);
/* steps 1b + 2 + 3 in ONE pass and in packed half arithmetic (Ch = 2048 only; replaces VidHRFormer.py:377-381, i.e.
 * norm1 -> act1 -> dw3x3 -> norm2 -> act2 of MlpDWBN.forward):  out = GELU(LN2(dw3x3(GELU(LN1(h1))) + b)).
 *   h1_f16    IEEE half [frames,64,Ch]: fc1 output (npvp_gemm_bf16 with out16 = 1)
 *   part1     fp32 [frames,32,2]: the fc1 epilogue's partial (sum, sumsq) (npvp_epilogue_t.frame_stats); reduced here
 *   ln_wb_f16 half [2 norms][64 px][Ch/2 pairs][(w_c, w_c+1), (b_c, b_c+1)]: both LayerNorm affines, pair-interleaved
 *   dw_w_f16  half [9,Ch], dw_b_f16 half [Ch]
 *   out_f16   half [frames,64,Ch] (!= h1): feeds fc2 (npvp_gemm_bf16 with fp16 = 1)
 *   xch       fp32 [>= frames,32,2], 256-byte aligned, and cnt uint32 [>= frames]: statistics exchange between the 32 warps
 *             (on 32 SMs) that share a frame.  PERSISTENT scratch owned by the caller: xch must hold the all-ones bit
 *             pattern and cnt zeros before the first call; every call leaves them in that state again (no memset per call).
 * All blocks of the launch are co-resident (npvp_ffn_mid16_lanes() x 32 blocks; 0 = the kernel does not fit the device).
 * Statistics are fp32 / fp64; element-wise math is half2 (max |GELU error| 1.0e-3 = half an ulp at |x| ~ 3). */
int npvp_ffn_mid16(const void* h1_f16, const float* part1, const void* ln_wb_f16, const void* dw_w_f16, const void* dw_b_f16,
                   void* out_f16, float* xch, unsigned int* cnt, int64_t frames, int64_t Ch, void* stream);
int npvp_ffn_mid16_lanes(void);

/* ---- predictor: attention cores (8 heads x 64) ---------------------------------------------
 * softmax(Q K^T / 8 [+mask]) V for the short sequences of the factorised attention
 * (nn.MultiheadAttention cores at VidHRFormer.py:104,221,239,298 + window permutes :447-475).
 * q/k/v/out are bf16 token matrices (row = token, 512 used columns, row strides ld*).
 * SPATIAL_WINDOW: q,k,v have n_clips*Tq*64 rows; TEMPORAL: q/out have n_clips*Tq*64 rows,
 * k/v have n_clips*Tk*64 rows.  mask_last=1 reproduces the encoder quirk (:100-102):
 * queries 0..Tq-2 may not attend to key Tk-1. */
int npvp_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                   int64_t ldo, int mode, int64_t n_clips, int Tq, int Tk, int mask_last, void* stream);

/* ---- predictor: event encoder / latent (models/submodules.py:388-410) --------------------------
 * depthwise 3x3 (zero pad) on the 8x8 grid, channels-last, folded BN scale in w, shift + optional ReLU.
 * x fp32 [frames,64,C]; w fp32 [9,C]; shift fp32 [C]; out bf16 [frames,64,C]. */
int npvp_dwconv3x3_tokens(const float* x, const float* w, const float* shift, void* out_bf16, int64_t frames,
                          int64_t C, int relu, void* stream);
/* z = mu + exp(0.5*logvar) * eps.  mulv fp32 [n*64, 2*C] = (mu | logvar) token-major; eps fp32 NCHW [n,C,8,8]
 * (the reference's layout, submodules.py:408-410) or NULL (z = mu); z fp32 [n,64,C] token-major. */
int npvp_latent_reparam(const float* mulv, int64_t ld, const float* eps_nchw, float* z, int64_t n_clips, int64_t C,
                        void* stream);

/* ---- layout changes at the module boundary --------------------------------------------------
 * (N,T,C,H,W) fp32  <->  channels-last tokens.  frames = N*T, HW = H*W. */
int npvp_nchw_to_tokens(const float* x, float* out_f32, void* out_bf16, int64_t frames, int64_t C, int64_t HW,
                        int fp16, void* stream);
int npvp_tokens_to_nchw(const float* x_f32, const void* x_bf16, float* out, int64_t frames, int64_t C, int64_t HW,
                        int relu, int fp16, void* stream);

/* ---- autoencoder ---------------------------------------------------------------------------
 * 7x7 stem: reflect-pad 3, conv (no bias) + folded BN + ReLU  (ResNetAutoEncoder.py:70-73).
 * x fp32 NCHW [frames,Cin,H,W]  OR  x_u8 uint8 pixels [frames,Cin,H,W] with the dataset's VidNormalize mean / std (host pointers,
 * Cin values each): the stem then applies VidToTensor + VidNormalize itself, ((u8 / 255) - mean) / std in the reference's
 * operation order (utils/dataset.py:835-858) - exactly one of x and x_u8 is non-NULL;
 * w fp32 [49*Cin, Cout] (tap-major, BN scale folded); shift fp32 [Cout]; out bf16 NHWC [frames,H,W,Cout].
 * Cout 32 / 64 with W <= 250 runs on tcgen05 (csrc/head_tc.cu: im2col row blocks built once per input row in a shared-memory ring,
 * 7 taps x 2 MMAs of M 128 x N Cout per 128 positions); otherwise the mma.sync tile kernel.  npvp_set_option("stem_tc", 0) forces the latter. */
int npvp_conv7x7_stem(const float* x, const float* w, const float* shift, void* out_bf16, int64_t frames, int Cin,
                      int Cout, int H, int W, int fp16, const void* x_u8, const float* norm_mean, const float* norm_std, void* stream);
/* 7x7 head: reflect-pad 3, conv + bias + Tanh|Sigmoid (ResNetAutoEncoder.py:184-189).
 * x 16-bit [frames,H,W,Cin] (phase_major=1: stored [frames,H/2,W/2,4,Cin], the ConvT GEMM's native output), Cin 32 or 64;
 * w 16-bit, pre-packed as mma.sync B fragments [Cin/32][14 k-steps = (ky, 16-channel half)][NT = ceil(7 Cout / 8)][32 lanes][4]
 * of the matrix B[(ky,ci), n = kx*Cout + co] (see pack_head_weights in npvp_b200/_lib.py); Cout in [1,3]; bias fp32 [Cout];
 * out fp32 NCHW [frames,Cout,H,W] (model space) and / or out_u8 uint8 NCHW: the pixel-space frame the reference writes to image
 * files - VidReNormalize, clamp to [0,1], ToPILImage's trunc(255 v) (utils/dataset.py:860-886, utils/train_summary.py:243-248) -
 * fused into the epilogue; pix_inv_std / pix_inv_mean (host pointers, Cout values: 1/std and -mean as npvp_frames_to_pixels
 * takes them) are required with out_u8.  At least one of out / out_u8 is non-NULL.
 * Plain NHWC input (phase_major = 0) with W <= 250 runs on tcgen05 (csrc/head_tc.cu: whole reflect-padded rows streamed once
 * through a shared-memory ring by TMA, 7 taps x Cin/16 MMAs of M 128 x N 32 per 128 pixels); otherwise the mma.sync tile kernel.
 * Same packed weights, same results to 2e-6; npvp_set_option("head_tc", 0) forces the tile kernel. */
int npvp_conv7x7_head(const void* x_bf16, const void* w, const float* bias, float* out, int64_t frames, int Cin,
                      int Cout, int H, int W, int phase_major, int act, int fp16, void* out_u8, const float* pix_inv_std,
                      const float* pix_inv_mean, void* stream);
/* 2x2/stride-2 max-pool of a column slice of a token matrix (NonLocalAttenion2D k/v pooling, submodules.py:151,158).
 * x bf16 [frames*H*W, ldx], columns [col0, col0+Cn) -> out bf16 [frames*(H/2)*(W/2), Cn]. */
int npvp_maxpool2x2_cols(const void* x_bf16, int64_t ldx, int col0, int Cn, void* out_bf16, int64_t frames, int H,
                         int W, int fp16, void* stream);
/* Non-local attention core: softmax(q k^T) v, UNSCALED (submodules.py:153-160).
 * q bf16 [frames*HW, ldq] (first dq columns); kv bf16 [frames*HWk, dq+dv] (k | v); out bf16 [frames*HW, dv]. */
int npvp_nonlocal_attention(const void* q, int64_t ldq, const void* kv, void* out, int64_t frames, int HW, int HWk,
                            int dq, int dv, int fp16, void* stream);

/* ---- pixel-space post-processing and evaluation metrics (the consumers of the predicted frames) ----------------
 * Model space -> [0,1] pixel space, fp32 NCHW [n_images,C,HW]:  v = clamp(x / inv_std[c] - inv_mean[c], 0, 1), exactly
 * the fp32 operation order of VidReNormalize + clamp (utils/dataset.py:860-886, utils/train_summary.py:243-245) with
 * inv_std = 1/std, inv_mean = -mean (HOST arrays of C floats).  out_f32 and / or out_u8 (= trunc(v * 255), what ToPILImage
 * writes, train_summary.py:246-248) may be NULL. */
int npvp_frames_to_pixels(const float* frames, const float* inv_std, const float* inv_mean, float* out_f32, void* out_u8,
                          int64_t n_images, int C, int64_t HW, void* stream);
/* uint8 pixels NCHW -> model space fp32: ((u / 255) - mean[c]) / std[c]  (VidToTensor + VidNormalize, utils/dataset.py:835-858);
 * mean / std are HOST arrays of C floats. */
int npvp_pixels_to_frames(const void* in_u8, const float* mean, const float* std, float* out, int64_t n_images, int C,
                          int64_t HW, void* stream);
/* PSNR per image (utils/metrics.py:12-30, mean_flag=False): -10 log10(mean((x/r - y/r)^2) + 1e-8); x, y fp32 [n_images, elems]. */
int npvp_psnr(const float* x, const float* y, float* out, int64_t n_images, int64_t elems, float data_range, void* stream);
/* SSIM per image (utils/metrics.py:47-109, mean_flag=False): depthwise 11x11 Gaussian window, zero padding 5, mean over
 * (C,H,W).  window11: HOST array, the normalised 1-D Gaussian (the reference's 2-D window is its outer product). */
int npvp_ssim(const float* x, const float* y, const float* window11, float* out, int64_t n_images, int C, int H, int W,
              void* stream);
/* Best-of-K evaluation of stochastic samples (NPVP-S; utils/metrics.py:12-109 applied per sample): per-frame PSNR
 * (window11 == NULL) or SSIM (window11 = host pointer to the 11 normalised Gaussian taps) of samples fp32
 * [n_clips][K][T][C,H,W] against gt fp32 [n_clips][T][C,H,W] (read in place, not replicated) -> scores fp32 [n_clips][K][T]. */
int npvp_sample_scores(const float* samples, const float* gt, const float* window11, float* scores, int64_t n_clips, int K, int T,
                       int C, int H, int W, float data_range, void* stream);
/* mean_scores[n][k] = mean_t scores[n][k][t]; best_idx[n] = argmax_k (ties: lowest k); best (optional, NULL to skip)
 * receives the winner's frames: clip_elems = T*C*H*W fp32 values per sample (multiple of 4, 16-byte aligned buffers). */
int npvp_best_of_k(const float* scores, const float* samples, int64_t n_clips, int K, int T, int64_t clip_elems, int32_t* best_idx,
                   float* mean_scores, float* best, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NPVP_B200_H */
