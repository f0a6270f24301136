/* tvs_b200.h - C ABI of the B200-native TuneVLSeg prompt-tuning hot path (libtvs_b200.so).
 *
 * The reference (naamiinepal/tunevlseg) has NO native / FFI layer: its hot path is Python calling ATen
 * library kernels through transformers / monai / torchmetrics.  This header is
 * therefore the boundary a maintainer would bind from the reference's Python wrappers (ctypes stub in
 * INTEGRATION.md); every entry cites the reference call site(s) (relative to /root/reference, or to
 * site-packages/transformers/models/clipseg/modeling_clipseg.py = "hf:") whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     never allocates, and returns 0 on success or a negative code (message: tvs_last_error());
 *   - row-major everywhere; "ld*" are leading dimensions in ELEMENTS;
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t), f32 = float, i64 = int64_t, u8 = uint8_t;
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute entry fails.
 */
#ifndef TVS_B200_H
#define TVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVS_ABI_VERSION 3

int tvs_version(void);
const char* tvs_last_error(void);
/* number of kernels launched by this library in the calling process since load (bench.py: gpu_launches) */
int64_t tvs_launch_count(void);
/* 0 when the current device is sm_100 and the library can run on it */
int tvs_device_check(void);
/* template instance chosen by the calling thread's last tvs_gemm_bf16 launch:
 * (tile_n << 16) | (pipeline stages << 8) | (epilogue << 5) | (tf32 << 4) | cta_group (1 = independent CTAs, 2 = cta_group::2
 * pairs); epilogue: 0 = generic (runtime flags), 1-4 = the straight-line epilogues of the hot tower shapes (gemm_sm100.cu).
 * Lets tests assert that the benched shapes really run the pair kernel. */
int32_t tvs_gemm_last_variant(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = epilogue( A[M,K] * W[N,K]^T )  on tcgen05 / TMEM, operands by TMA, fp32 accumulate.
 * Replaces every nn.Linear on the path (hf::297-300 q/k/v/out_proj, hf::345-353 fc1/fc2, decoder
 * reduces / film hf::551-552,582-584, text/visual projection) and, with a transposed weight copy,
 * their dgrad (frozen backbone: there is no wgrad).
 *
 * epilogue, in this order (each step optional):
 *     v  = acc (+ bias[n])
 *     pre_bf16[m,n] = v                      save the pre-activation for the backward pass
 *     v  = act(v)                            act: TVS_ACT_*
 *     v += residual_f32[m,n]
 *     out_f32[m,n] = v ; out_bf16[m,n] = v
 * TVS_ACT_RES_RELU applies the ReLU AFTER the residual add: relu(acc + bias + residual) - the tail of a ResNet
 * bottleneck (cris_model/clip.py:66-75).
 * TVS_ACT_DQGELU / TVS_ACT_DRELU multiply v by act'(aux_bf16[m,n]) (aux = saved pre-activation, or the
 * post-ReLU activation for DRELU) - the dgrad-through-activation epilogue.
 * Requirements: K % 8 == 0, lda % 8 == 0, ldw % 8 == 0, A and W 16-byte aligned.
 * ------------------------------------------------------------------------------------------------ */
/* operand formats.  kind::f16 MMAs take IEEE fp16 or bf16 operands at the same rate: the vision tower's FORWARD operands
 * (LayerNorm / attention / GELU outputs and the frozen weights) are fp16 - three more significand bits than bf16 cut the
 * logit error 2.5x (tools/precision_attribution.py) - while gradients stay bf16 (range).
 * A and W must have the SAME format: although the instruction descriptor has one format field per operand, a B200 raises
 * "illegal instruction" for a mixed fp16 x bf16 kind::f16 MMA (measured, round 2).
 * The 16-bit pointers below are named *_bf16 for history; their format follows these flags. */
enum { TVS_AB_BF16 = 0, TVS_AB_TF32 = 1, TVS_AB_F16 = 2 /* A, W IEEE fp16 */ };
enum { TVS_GEMM_ROUND_OUT_TF32 = 1, TVS_GEMM_OUT16_F16 = 2,
       TVS_GEMM_PRE_DGELU = 8 /* with TVS_ACT_QGELU and pre_bf16: store QuickGELU'(pre-activation) instead of the pre-activation
                                 (the forward has the sigmoid in registers anyway; the dgrad then only multiplies, TVS_ACT_MULAUX) */,
       TVS_GEMM_STREAM_K = 4 /* opt-in: cut the (tile, k-block) space into one contiguous range per CTA pair (specialised pair
                                kernels only; measured SLOWER than whole-tile round robin on the tower shapes) */ };     /* tvs_gemm_args.reserved (flags) */
enum { TVS_ACT_NONE = 0, TVS_ACT_QGELU = 1, TVS_ACT_RELU = 2, TVS_ACT_DQGELU = 3, TVS_ACT_DRELU = 4, TVS_ACT_RES_RELU = 5,
       TVS_ACT_MULAUX = 6 /* v *= aux_bf16[m,n]: the dgrad through an activation whose DERIVATIVE the forward saved (TVS_GEMM_PRE_DGELU) */ };

typedef struct tvs_gemm_args {
    const void* A;  int64_t lda;           /* bf16 [M,K] */
    const void* W;  int64_t ldw;           /* bf16 [N,K] */
    int32_t M, N, K;
    const float* bias;                     /* f32 [N] or NULL */
    const float* residual; int64_t ldr;    /* f32 [M,N] or NULL (may alias out_f32) */
    float* out_f32;  int64_t ldo32;        /* f32 [M,N] or NULL */
    void*  out_bf16; int64_t ldo16;        /* bf16 [M,N] or NULL */
    void*  pre_bf16; int64_t ldpre;        /* bf16 [M,N] or NULL */
    const void* aux_bf16; int64_t ldaux;   /* bf16 [M,N], required by TVS_ACT_D* */
    int32_t act;
    int32_t tile_n;                        /* 0 = auto, else 64 / 128 / 256 */
    int32_t ab_dtype;                      /* TVS_AB_BF16: A, W are bf16 (kind::f16).  TVS_AB_TF32: A, W are f32 and the
                                              MMA runs kind::tf32 (10-bit mantissa) - used for the small text tower and
                                              decoder, whose rounding dominates the logit error; K, lda, ldw % 4 == 0 */
    int32_t reserved;                      /* flags; TVS_GEMM_ROUND_OUT_TF32: round out_f32 to nearest tf32 (cvt.rna) - for outputs
                                              that only feed further TVS_AB_TF32 GEMMs, whose MMA truncates its operands;
                                              TVS_GEMM_OUT16_F16: out_bf16 is written as IEEE fp16 (pre_bf16 stays bf16) */
    int32_t conv_h, conv_w;                /* != 0: implicit-GEMM 3x3 convolution, stride 1, pad 1 (cris_model/layers.py:14-26,
                                              clip.py:26-27).  A = zero-bordered channels-last image [B, H+2, W+2, C] (from
                                              tvs_pad_nhwc), lda = C, M = B*(H+2)*(W+2), K = 9*C, W = [N, 9*C] with K ordered
                                              (ky, kx, c).  Outputs / residual / aux are UNPADDED [B*H*W, N] matrices: the
                                              epilogue maps rows and skips the border.  No im2col matrix ever exists: each k-block
                                              is a TMA load of the A rows shifted by its tap's offset. */
    /* Deep-prompt overwrite fused into the residual epilogue (base_multimodal_clipseg.py:394-398, base_visual_learner.py:18-23:
     * the prompt rows of the block output are REPLACED by the next depth's context).  When ovr_ctx != NULL the output rows whose
     * position inside their sample (row % ovr_S) lies in [ovr_row0, ovr_row0 + ovr_n) receive
     * ovr_ctx[(row / ovr_S) * ovr_batch_stride + (row % ovr_S - ovr_row0) * N + col] instead of the GEMM result
     * (ovr_batch_stride = 0: one shared context; n * N: per-sample contexts, CoCoOp).  Implemented for the 16-bit
     * bias + f32-residual -> f32 GEMM (fc2 of a vision block); other configurations are rejected - use tvs_prompt_overwrite. */
    const float* ovr_ctx; int64_t ovr_batch_stride;
    int32_t ovr_S, ovr_row0, ovr_n, ovr_reserved;
} tvs_gemm_args;

int tvs_gemm_bf16(const tvs_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension D (eps inside the sqrt), one warp per row.
 * Replaces nn.LayerNorm at hf::362-365 (layer_norm1/2), hf::782 (pre_layrnorm), hf::636
 * (final_layer_norm), hf::784 (post_layernorm), decoder post-norms hf::395-398.
 * fwd: y = (x - mean) * rstd * gamma + beta ; saves mean/rstd (f32 [M]) when non-NULL.
 * bwd: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma   (no dgamma/dbeta: frozen)
 *      dx_out_f32 = dx (+ dx_add_f32) ; dx_out_bf16 = same, rounded.  dy is bf16 (dy_bf16) or f32 (dy_f32).
 * D % 4 == 0, D <= 2048 (2048: the CRIS decoder FFN norm, layers.py:303-309).
 * ------------------------------------------------------------------------------------------------ */
enum { TVS_LN_ROUND_TF32 = 1, TVS_LN_Y16_F16 = 2 };
int tvs_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int64_t M, int32_t D,
                      float* y_f32, void* y_bf16, float* mean, float* rstd, int32_t round_tf32, void* stream);
/* round_tf32 (flags): bit 0 (TVS_LN_ROUND_TF32): y_f32 is rounded to nearest tf32 - it feeds a TVS_AB_TF32 GEMM, whose MMA
 * truncates its operands; bit 1 (TVS_LN_Y16_F16): the 16-bit output y_bf16 is written as IEEE fp16.
 * (tvs_attn_fwd's out_f32 and tvs_cross_attn_fwd's out are always rounded that way: they only feed such GEMMs.) */
int tvs_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma,
                      const float* mean, const float* rstd, const float* dx_add_f32, int64_t M, int32_t D,
                      float* dx_out_f32, void* dx_out_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-head self-attention, flash style (scores never leave the SM), fp32 softmax.
 * Replaces eager_attention_forward hf::256-276 as called from CLIPSegAttention hf::302-338.
 * qkv: bf16 [B*S, 3*H*hd]  (columns: Q | K | V, each head-major; the d^-0.5 scale is pre-folded into Wq)
 * out: bf16 [B*S, H*hd] ; lse: f32 [B,H,S] (natural-log-sum-exp of the row, saved for the backward)
 * causal != 0: key j > query i is masked (text tower, base_multimodal_clipseg.py:205-209);
 * key_mask: u8 [B,S] 1 = attend, 0 = padding, or NULL (base_multimodal_clipseg.py:212-222).
 * hd is 64 (towers) or 16 (decoder).  bwd writes dqkv (same layout as qkv); delta is f32 [B,H,S] scratch.
 * ------------------------------------------------------------------------------------------------ */
/* flags: TVS_ATTN_O_F16 - `out` (written by fwd, read by bwd for delta = rowsum(dO o O)) is IEEE fp16 instead of bf16: it is
 * the A operand of the out-projection GEMM (TVS_AB_F16).  tcgen05 path only (hd = 64, no masks); qkv, dout, dqkv stay bf16. */
enum { TVS_ATTN_O_F16 = 1 };
int tvs_attn_fwd(const void* qkv, int32_t B, int32_t S, int32_t H, int32_t hd, int32_t causal,
                 const uint8_t* key_mask, void* out, float* out_f32 /* optional f32 copy of out */, float* lse,
                 int32_t flags, void* stream);
int tvs_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int32_t B, int32_t S,
                 int32_t H, int32_t hd, int32_t causal, const uint8_t* key_mask, float* delta, void* dqkv,
                 int32_t flags, void* stream);
/* Bottom block of a prompted tower: below it only the prompt rows (the last n of every sample) still carry a gradient
 * (base_visual_learner.py:18-23 appends them last; the patch / class embeddings are frozen), so dqkv is only needed for
 * rows >= row_begin.  Rows of dqkv below the 128-row tile that contains row_begin are left untouched (hd = 64 without
 * masks); other configurations compute every row like tvs_attn_bwd. */
int tvs_attn_bwd_tail(const void* qkv, const void* out, const void* dout, const float* lse, int32_t B, int32_t S,
                      int32_t H, int32_t hd, int32_t causal, const uint8_t* key_mask, float* delta, void* dqkv,
                      int32_t row_begin, int32_t flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Vision embedding pieces (hf::196-212 CLIPSegVisionEmbeddings.forward; base_multimodal_clipseg.py:449-465)
 * im2col: image f32 [B,3,H,W] -> bf16 [B*g*g, 3*P*P] (column order c,py,px = Conv2d weight flattening)
 * assemble: h[b,0] = cls + pos[0]; h[b,1+p] = patches[b*g*g+p] + pos[1+p]; h[b,1+g*g+j] = ctx[(b),j]
 *           (ctx_batch_stride = 0 for a shared prompt, n*D for per-sample prompts); n may be 0.
 * ------------------------------------------------------------------------------------------------ */
int tvs_im2col_patches(const float* image, int32_t B, int32_t C, int32_t H, int32_t W, int32_t P, void* out_bf16,
                       int32_t out_f16 /* != 0: IEEE fp16 columns (TVS_AB_F16 patch-embedding GEMM) */, void* stream);
int tvs_vision_assemble(const float* patches, const float* cls, const float* pos, const float* ctx,
                        int64_t ctx_batch_stride, int32_t B, int32_t G2, int32_t n, int32_t D, float* h,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Deep-prompt row replacement and its backward (base_visual_learner.py:18-23: h[:, -n:] = ctx;
 * coop_context_learner.py:124-134: h[:, 1:n+1] = ctx).  x: f32 [B,S,D]; rows row0..row0+n-1.
 * grad: dctx[(b),j,:] (+)= sum_b dx[b,row0+j,:]  then dx rows are zeroed (the overwritten values have no
 * upstream).  ctx_batch_stride = 0 -> reduced over the batch, else per-sample.
 * ------------------------------------------------------------------------------------------------ */
int tvs_prompt_overwrite(float* x, void* x_bf16, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t n,
                         const float* ctx, int64_t ctx_batch_stride, void* stream);
int tvs_prompt_grad(float* dx, void* dx_bf16 /* optional bf16 mirror of dx whose rows are zeroed too */, int32_t B,
                    int32_t S, int32_t D, int32_t row0, int32_t n, float* dctx, int64_t ctx_batch_stride,
                    int32_t zero_rows, void* stream);

/* Row slicing around the decoder (base_clipseg.py:132-142: output[:, 1:-n] drops the CLS and prompt rows):
 * slice:   y[b, r, :] = x[b, row0 + r, :], r < nrows   (f32 and/or bf16 copy)
 * unslice: dx[b, s, :] = dy[b, s - row0, :] inside the window, 0 elsewhere (the backward of slice) */
int tvs_slice_rows(const float* x, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t nrows, float* y_f32,
                   void* y_bf16, void* stream);
int tvs_unslice_rows(const float* dy, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t nrows, float* dx,
                     void* stream);
/* dw[n,k] += sum_m dy[m,n] * x[m,k] for a small trainable weight (N*K <= 4096): wgrad of the additive
 * Conv2d(64,1,k) of base_clipseg.py:63-70 contracted at low resolution.  dw must be zero-initialised. */
int tvs_wgrad_small(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, int64_t M, int32_t N, int32_t K,
                    float* dw, void* stream);

/* f32 -> bf16 copy (n elements, n % 4 == 0) */
int tvs_cast_bf16(const float* x, void* y_bf16, int64_t n, void* stream);
/* y (+)= x for f32 buffers (gradient merge at the decoder taps) */
int tvs_add_f32(float* y, const float* x, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FiLM conditioning (base_clipseg.py:111-115): y[b,s,:] = mul[b,:] * x[b,s,:] + add[b,:]  (f32)
 * bwd: dx = mul * dy ; dmul[b,:] = sum_s dy * x ; dadd[b,:] = sum_s dy
 * ------------------------------------------------------------------------------------------------ */
int tvs_film_fwd(const float* x, const float* mul, const float* add, int32_t B, int32_t S, int32_t D, float* y,
                 void* y_bf16, void* stream);
int tvs_film_bwd(const float* dy, const float* x, const float* mul, int32_t B, int32_t S, int32_t D, float* dx,
                 float* dmul, float* dadd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder head (base_clipseg.py:132-157, vpt_clipseg.py:287-304; hf::573-575 ConvTranspose2d(64,1,P,P)):
 *   tconv: f32 [B*g*g, P*P]  = feat @ Wt^T (from tvs_gemm_bf16; no bias)     -> pixel shuffle + bias_t
 *   addmap: f32 [B*g*g, KK]  = feat @ Wa^T  (KK = k*k taps of the additive Conv2d, contracted over the
 *           channels at LOW resolution; bilinear upsampling and the k x k replicate-padded stencil commute
 *           with that contraction, so the (B,64,H,W) upsampled tensor is never materialised)
 *   logits[b,Y,X] = wa * (tconv + bias_t) + wb * (sum_k up(addmap_k)[clamp(Y+ky-k/2), clamp(X+kx-k/2)] + bias_a)
 *   blend: 0 none (wa=1, wb=0), 1 ratio (wa = 1-r, wb = r), 2 add (wa = wb = 1).  r read from *ratio.
 * bwd: dtconv (bf16 [B*g*g, P*P]) = wa * dlogits ; daddmap (f32 [B*g*g, KK]) ; dbias_a, dratio (f32 scalars,
 *      accumulated with atomics into zero-initialised outputs).
 * ------------------------------------------------------------------------------------------------ */
int tvs_head_fwd(const float* tconv, int64_t ld_tconv, const float* addmap, int64_t ld_addmap, const float* bias_t,
                 const float* bias_a, const float* ratio, int32_t blend, int32_t B, int32_t G, int32_t P,
                 int32_t ksize, float* logits, float* add_out /* f32 [B,H,W] additive branch incl. bias, or NULL */,
                 void* stream);
int tvs_head_bwd(const float* dlogits, const float* tconv, int64_t ld_tconv, const float* add_out,
                 const float* bias_t, const float* ratio, int32_t blend, int32_t B, int32_t G, int32_t P,
                 int32_t ksize, void* dtconv_bf16, int64_t ld_dtconv, float* daddmap, int64_t ld_daddmap,
                 float* dbias_a, float* dratio, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused Dice+BCE loss and Dice / IoU counters in ONE pass over logits and mask.
 * Replaces monai.losses.DiceCELoss(sigmoid=True, lambda_dice, lambda_ce) (configs/model/maple_clipseg.yaml:29-33),
 * torch.sigmoid + mask.long() (src/models/image_text_mask_module.py:103-107) and the update step of
 * torchmetrics Dice(threshold, average="samples") / JaccardIndex(task="binary") (:118-119, :284-298).
 *   logits, mask: f32 [B, N]   (N = H*W; mask in [0,1], target = (int64)mask)
 *   parts:  f64 [B,4]  = { sum p*y, sum p, sum y, sum bce }      (p = sigmoid(logit))
 *   counts: i64 [B,3]  = { tp, fp, fn } with p >= threshold       (Dice, bit-exact)
 *   confmat:i64 [4]    = { tn, fp, fn, tp } with p >  threshold, ACCUMULATED into the buffer (IoU state)
 *   loss:   f32 [1]    = lambda_dice * mean_b(1 - (2I+1e-5)/(P+G+1e-5)) + lambda_ce * mean bce
 *   scratch: >= tvs_dicebce_scratch_bytes(B, N) bytes
 * bwd: dlogits[b,i] = gscale[0] * dloss/dlogit  (gscale: f32 [1] device scalar = upstream grad)
 * ------------------------------------------------------------------------------------------------ */
int64_t tvs_dicebce_scratch_bytes(int32_t B, int64_t N);
int tvs_dicebce_metrics_fwd(const float* logits, const float* mask, int32_t B, int64_t N, float threshold,
                            float lambda_dice, float lambda_ce, double* parts, int64_t* counts, int64_t* confmat,
                            float* loss, void* scratch, void* stream);
/* counters only, from probabilities the caller already holds (metric(preds, target) API of
 * src/models/image_text_mask_module.py:118-119); scratch as above */
int tvs_metrics_from_probs(const float* preds, const float* mask, int32_t B, int64_t N, float threshold,
                           int64_t* counts, int64_t* confmat, void* scratch, void* stream);
int tvs_dicebce_bwd(const float* logits, const float* mask, const double* parts, const float* gscale, int32_t B,
                    int64_t N, float lambda_dice, float lambda_ce, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AdamW over one flat f32 buffer (torch.optim.AdamW semantics; configs/model/maple_clipseg.yaml:36-39).
 * grad is multiplied by grad_scale first (1/world after the NCCL sum).  step is 1-based.
 * step_dev / lr_dev: optional DEVICE scalars that override step / lr, so that a captured CUDA graph can be
 * replayed with an advancing step (tvs_counter_inc bumps the device counter from inside the graph).
 * ------------------------------------------------------------------------------------------------ */
int tvs_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                   const int32_t* step_dev, const float* lr_dev, void* stream);
int tvs_counter_inc(int32_t* counter_dev, void* stream);

/* ================================================================================================
 * CRIS path (src/models/components/cris_model, src/models/core_models/coop/coop_cris.py).
 * Activations are channels-last matrices [B*H*W, C]; a k x k convolution = tvs_im2col_nhwc + tvs_gemm_bf16
 * (BatchNorm folded into the weight rows, ReLU / residual in the GEMM epilogue); its dgrad = GEMM with the
 * transposed weight + tvs_col2im_nhwc.  Replaces nn.Conv2d / nn.BatchNorm2d / F.avg_pool2d / F.interpolate /
 * nn.MultiheadAttention / F.conv2d(groups=B) at the cited lines.
 * ================================================================================================ */
/* col[(b,oy,ox), (ky,kx,c)] = x[b, oy*stride - pad + ky, ox*stride - pad + kx, c] (0 outside); columns
 * [ksize*ksize*C, ldcol) are zero-filled (K padding for the GEMM).  elem_bytes 2 (bf16) or 4 (f32).
 * clip.py:199-218 (stem), :26-27 (bottleneck conv2), layers.py:14-26 (conv_layer 3x3). */
int tvs_im2col_nhwc(const void* x, int32_t elem_bytes, int32_t B, int32_t H, int32_t W, int32_t C, int32_t ksize,
                    int32_t stride, int32_t pad, void* col, int64_t ldcol, int32_t round_tf32, void* stream);
/* xp[b, y+1, x+1, :] = x[b, y, x, :], zero border: the operand layout of the implicit-GEMM convolution above.
 * elem_bytes 2 or 4; round_tf32 rounds f32 elements to nearest tf32 on the way. */
int tvs_pad_nhwc(const void* x, int32_t elem_bytes, int32_t B, int32_t H, int32_t W, int32_t C, void* xp, int32_t round_tf32,
                 void* stream);
/* y = x rounded to nearest tf32 (cvt.rna), [M, C] f32 views.  The kind::tf32 MMA truncates its operands (a
 * systematic -3.4e-4 relative bias per GEMM that compounds over the ~60 layers of CLIP-RN50), so forward operands are
 * rounded on their way into tvs_gemm_bf16: here, or inside tvs_im2col_nhwc (round_tf32 != 0) for k x k convolutions. */
int tvs_round_tf32(const float* x, int64_t ld_x, int64_t M, int32_t C, float* y, int64_t ld_y, void* stream);
/* dgrad of a stride-1 'same' k x k conv from dcol = dy @ W: dx[b,y,x,c] = sum_taps dcol[...] for c < Cx <= Ccol,
 * times (relu_mask > 0) when given (the ReLU of the layer that produced x). */
int tvs_col2im_nhwc(const float* dcol, int64_t ldcol, int32_t B, int32_t H, int32_t W, int32_t Ccol, int32_t Cx,
                    int32_t ksize, const float* relu_mask, int64_t ld_mask, float* dx, int64_t ld_dx, void* stream);
/* out = y > 0 ? dy : 0 on [M, C] views with row strides (ReLU backward in front of a 1x1-conv dgrad GEMM) */
int tvs_relu_mask(const float* dy, int64_t ld_dy, const float* y, int64_t ld_y, int64_t M, int32_t C, float* out,
                  int64_t ld_out, void* stream);
/* 2x2 / stride-2 average pooling (clip.py:32,47,222 nn.AvgPool2d(2); layers.py:432 F.avg_pool2d) */
/* is_bf16: bit 0 = bf16 elements (else f32); bit 1 = round the f32 result to nearest tf32 (feeds a tf32 GEMM) */
int tvs_avgpool2_nhwc(const void* x, int32_t is_bf16, int32_t B, int32_t H, int32_t W, int32_t C, void* y,
                      int64_t ld_out, void* stream);
/* bilinear x2, align_corners=False (layers.py:85-88 nn.Upsample, :427,:441 F.interpolate); f32.
 * fwd writes rows of ld_out floats (a column slice of a concat buffer); bwd reads dy rows of ld_dy floats. */
int tvs_upsample2x_fwd(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, float* y, int64_t ld_out, void* stream);
int tvs_upsample2x_bwd(const float* dy, int64_t ld_dy, int32_t B, int32_t H, int32_t W, int32_t C, float* dx, void* stream);
/* Attention over a SHORT key set in fp32: Sq queries x Sk <= 80 keys, head dim 64.  Used for the decoder's cross
 * attention (layers.py:341-349 multihead_attn with key_padding_mask) and, with causal != 0 (Sq == Sk, key j > query i
 * masked), for the <= 77-token CLIP text encoder of CRIS (clip.py:291-343), whose output steers the dynamic
 * convolution and needs more than bf16 operands.  q: [B*Sq, H*64] rows of ld_q (scale pre-folded), k / v: [B*Sk, H*64] rows of ld_kv,
 * key_mask u8 [B,Sk] 1 = attend or NULL; out [B*Sq, H*64]; lse, delta f32 [B,H,Sq]. */
int tvs_cross_attn_fwd(const float* q, int64_t ld_q, const float* k, const float* v, int64_t ld_kv,
                       const uint8_t* key_mask, int32_t B, int32_t Sq, int32_t Sk, int32_t H, int32_t hd, int32_t causal,
                       float* out, int64_t ld_o, float* lse, void* stream);
int tvs_cross_attn_bwd(const float* q, int64_t ld_q, const float* k, const float* v, int64_t ld_kv,
                       const uint8_t* key_mask, const float* out, const float* dout, int64_t ld_o, const float* lse,
                       int32_t B, int32_t Sq, int32_t Sk, int32_t H, int32_t hd, int32_t causal, float* dq, int64_t ld_dq,
                       float* dk, float* dv, int64_t ld_dkv, float* delta, void* stream);
/* Projector tail (layers.py:96-119): out[b,p] = bias[b] + sum_{c,t} x[b, p+off_t, c] * w[b, c*9+t], 3x3, zero pad.
 * x f32 [B*H*W, C]; w rows of ld_w floats (the txt Linear output), bias[b] at bias + b*ld_bias; taps f32 [B*H*W, 9]
 * scratch kept for nothing (recomputed in bwd).  bwd: dx, and dw_part f32 [chunks, B, C*9] partial sums over
 * pixel chunks (the caller adds them: deterministic). */
int tvs_dynconv_fwd(const float* x, const float* w, int64_t ld_w, const float* bias, int64_t ld_bias, int32_t B,
                    int32_t H, int32_t W, int32_t C, float* taps, float* out, void* stream);
int tvs_dynconv_bwd(const float* dout, const float* x, const float* w, int64_t ld_w, int32_t B, int32_t H, int32_t W,
                    int32_t C, float* dx, float* dw_part, int32_t chunks, void* stream);
/* Separable table-driven resampling of single-channel maps (coop_cris.py:235 bicubic, align_corners=True).
 * fwd: out[b,y,x] = sum_a sum_c wy[y,a] wx[x,c] in[b, iy[y,a], ix[x,c]]   (tables [Ho,ntaps] / [Wo,ntaps])
 * bwd: din[b,y,x] = sum over the transposed tables ty/twy [Hi,max_taps] with counts cy [Hi] (same for x).
 * tile > 0: the big map is stored in tvs_head_fwd's [B*(Ho/tile)*(Wo/tile), tile*tile] layout. */
int tvs_resample2d_fwd(const float* in, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, const int32_t* iy,
                       const float* wy, const int32_t* ix, const float* wx, int32_t ntaps, int32_t tile, float* out,
                       void* stream);
/* Predict tail (src/utils/save_utils.py:96-104): TF.resize(pred, mask_shape, BICUBIC, antialias=False) fused with
 * torchvision.utils.save_image's quantisation u8 = clamp(v * 255 + 0.5, 0, 255); tables as above (align_corners=False). */
int tvs_resample2d_u8(const float* in, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, const int32_t* iy,
                      const float* wy, const int32_t* ix, const float* wx, int32_t ntaps, uint8_t* out, void* stream);
int tvs_resample2d_bwd(const void* dout, int32_t dout_is_bf16, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho,
                       int32_t Wo, const int32_t* ty, const float* twy, const int32_t* cy, const int32_t* tx,
                       const float* twx, const int32_t* cx, int32_t max_taps, int32_t tile, float* din, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused feed-forward block of the CLIPSeg decoder layer (reduce_dim 64): CLIPSegMLP inside CLIPSegDecoderLayer,
 * hf::341-354 called from hf::421-431, with the residual of the post-LN block.  The [M, F] hidden activation stays in
 * tensor memory.   fwd: out = x + relu(x W1^T + b1) W2^T + b2       bwd (dgrad only): dx = g + ((g W2) o [x W1^T + b1 > 0]) W1
 * x, g, out, dx: f32 [M, 64]; w1_hi / w1_lo: bf16 [F, 64] head and tail of fc1.weight (w = hi + lo);
 * w2t_hi / w2t_lo: bf16 [F, 64] head and tail of fc2.weight^T; both tails NULL = single bf16 products.
 * With the tails every product is hi*hi + lo*hi + hi*lo in fp32 accumulation (~2^-17 relative).
 * F must be a multiple of 64, <= 4096.
 * ------------------------------------------------------------------------------------------------ */
int tvs_ffn64_fwd(const float* x, const void* w1_hi, const void* w1_lo, const void* w2t_hi, const void* w2t_lo,
                  const float* b1, const float* b2, int64_t M, int32_t D, int32_t F, float* out, void* stream);
int tvs_ffn64_bwd(const float* x, const float* g, const void* w1_hi, const void* w1_lo, const void* w2t_hi,
                  const void* w2t_lo, const float* b1, int64_t M, int32_t D, int32_t F, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Eval input transforms on the device (configs/experiment/coop/clipseg.yaml:113-127 `_eval_transforms`, applied by
 * src/data/core_datasets/image_text_mask_dataset.py:52-84): cv2.resize(INTER_CUBIC) of the uint8 HWC image in OpenCV's
 * 8-bit fixed-point arithmetic (four int taps scaled by 2048 per output column / row, replicated borders,
 * (v + 2^21) >> 22, saturation) + albumentations.Normalize + ToTensorV2 in one kernel; cv2.INTER_NEAREST for the mask.
 * img: device uint8 [Hi, Wi, 3] rows of ld_bytes; xofs / yofs int32 [Wo] / [Ho] first-tap source index + 1 (tap k reads
 * clamp(ofs + k - 1)); xcoef / ycoef int32 [Wo, 4] / [Ho, 4]; mean255 / inv_std255: HOST pointers to 3 floats
 * (255 mean, 1 / (255 std)).  out_chw f32 [3, Ho, Wo] and / or out_u8_hwc uint8 [Ho, Wo, 3] (the resized image itself).
 * tvs_resize_nearest_f32: in f32 [Hi, Wi] rows of ld floats; xofs / yofs int32 source indices; out f32 [Ho, Wo].
 * ------------------------------------------------------------------------------------------------ */
int tvs_preproc_image_u8(const uint8_t* img, int32_t Hi, int32_t Wi, int64_t ld_bytes, const int32_t* xofs,
                         const int32_t* xcoef, const int32_t* yofs, const int32_t* ycoef, const float* mean255,
                         const float* inv_std255, int32_t Ho, int32_t Wo, float* out_chw, uint8_t* out_u8_hwc,
                         void* stream);
int tvs_resize_nearest_f32(const float* in, int32_t Hi, int32_t Wi, int64_t ld, const int32_t* xofs,
                           const int32_t* yofs, int32_t Ho, int32_t Wo, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Train-time augmentations on the device (configs/experiment/coop/clipseg.yaml:84-103 `train_transforms`):
 * albumentations.Affine(interpolation=INTER_CUBIC, mode=BORDER_REPLICATE) = cv2.warpAffine on the resized uint8 image
 * (masks: INTER_NEAREST), RandomBrightnessContrast = a 256-entry uint8 look-up table, then Normalize + ToTensorV2.
 * warpAffine is OpenCV's fixed-point walk (imgproc/src/imgwarp.cpp): adelta / bdelta int32 [Wo] and x0 / y0 int32 [Ho]
 * are the tables cv::warpAffine builds from the inverted matrix (1/1024 pixel; x0 / y0 include the rounding term: 16 for
 * INTER_CUBIC, 512 for INTER_NEAREST); tab = int16 [32][32][4][4], cv::initInterTab2D(INTER_CUBIC, fixed point, 2^15),
 * device memory, 16-byte aligned.  lut: device uint8 [256] applied to the warped bytes, or NULL.  Outputs as in
 * tvs_preproc_image_u8 (mean255 / inv_std255: HOST pointers to 3 floats).  All device arithmetic is integer: the uint8
 * result equals cv2's bit for bit (oracle/augment.py, pinned to cv2).
 * tvs_lut_normalize_u8: img uint8 [H, W, 3] -> lut (or NULL) -> (x - mean255) * inv_std255 -> f32 [3, H, W].
 * ------------------------------------------------------------------------------------------------ */
int tvs_warp_affine_u8(const uint8_t* img, int32_t Hi, int32_t Wi, int64_t ld_bytes, const int32_t* adelta,
                       const int32_t* bdelta, const int32_t* x0, const int32_t* y0, const int16_t* tab,
                       const uint8_t* lut, const float* mean255, const float* inv_std255, int32_t Ho, int32_t Wo,
                       float* out_chw, uint8_t* out_u8_hwc, void* stream);
int tvs_warp_affine_nearest_f32(const float* in, int32_t Hi, int32_t Wi, int64_t ld, const int32_t* adelta,
                                const int32_t* bdelta, const int32_t* x0, const int32_t* y0, int32_t Ho, int32_t Wo,
                                float* out, void* stream);
int tvs_lut_normalize_u8(const uint8_t* img, int32_t H, int32_t W, int64_t ld_bytes, const uint8_t* lut,
                         const float* mean255, const float* inv_std255, float* out_chw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TVS_B200_H */
