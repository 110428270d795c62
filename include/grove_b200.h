/* grove_b200 — C ABI of the B200-native grounding path of GROVE (ekazakos/grove).
 *
 * The reference has no FFI: its boundary is the Python nn.Module API of model/SAM/modeling/*.py and the
 * grounding methods of model/GROVE.py.  grove_b200/modeling/*.py mirrors that API and
 * lowers every call onto the entry points below through ctypes (grove_b200/_lib.py); INTEGRATION.md
 * shows the binding a reference maintainer would add.  Conventions:
 *   - plain device pointers + sizes, no framework types; every function returns 0 on success or a
 *     GROVE_ERR_* code, with a message retrievable from grove_last_error();
 *   - no allocation inside: callers pass outputs and workspaces; all work is enqueued on `stream`;
 *   - "bf16" = __nv_bfloat16 bits; token-major activations are [frames * G*G, channels] (NHWC).
 * Each entry cites the reference code it replaces (file:line under the reference root).
 */
#ifndef GROVE_B200_H_
#define GROVE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__CUDACC__) || defined(__CUDA_RUNTIME_H__)
typedef cudaStream_t grove_stream_t;
#else
typedef void* grove_stream_t;
#endif

#define GROVE_B200_ABI_VERSION 5

/* ---- library ---------------------------------------------------------------------------------- */
int grove_abi_version(void);
const char* grove_last_error(void);
/* kernels launched by this library since the last reset (bench.py's gpu_launches) */
long long grove_launch_count(void);
void grove_reset_launch_count(void);
/* kernels launched by replaying a captured CUDA graph are reported by the host side (a replay does not pass through the entry points) */
void grove_add_launch_count(long long n);

/* ---- dense contractions on tcgen05 (gemm_tcgen05.cu) ------------------------------------------ */
typedef struct grove_gemm_epilogue {
  const float* bias;       /* [N] fp32 or NULL */
  const float* resid;      /* fp32 [*, N] or NULL (may alias an fp32 `out`: in-place residual stream) */
  int resid_row_mod;       /* >0: residual row = m % resid_row_mod (abs-pos embedding, image_encoder.py:176-177) */
  const float* gate_alpha; /* non-NULL: multiply by tanh(*gate_alpha) (adapter gate, image_encoder.py:45,54) */
  int act;                 /* 0 none | 1 exact-erf GELU (common.py:18) | 2 ReLU */
  int out_f32;             /* 1: `out` is fp32, 0: bf16 */
  void* out2_bf16;         /* optional second, bf16 copy of the output, or NULL */
  int max_ctas;            /* 0 = one persistent CTA per SM; >0 caps the grid (tests) */
  int force_ctas;          /* 0 = auto; 1 = single-CTA tiles; 2 = CTA-pair (cta_group::2) tiles when N % 256 == 0 (tests) */
  /* ---- training (ABI v2) ---- */
  int out2_pre_act;        /* 0: out2 = copy of out | 1 (bf16 out): out2 = value BEFORE the activation (MLP lin1 pre-GELU, saved for
                              backward) | 2 (fp32 out): out2 = value after the activation, before gate / residual (adapter ReLU output) */
  const void* dact_pre;    /* bf16 [M,N] or NULL: multiply the result by act'(dact_pre) — backward of the fused activation (bf16 out) */
  int dact;                /* with dact_pre: 1 = exact-GELU derivative, 2 = ReLU mask (pre > 0) */
  int splits;              /* >1: split-K; `out` is fp32 [splits, M, N] raw partial sums (no other epilogue field allowed);
                              finish with grove_reduce_partials_f32 */
  /* ---- ABI v3 ---- */
  void* workspace;         /* optional device scratch (16-byte aligned) or NULL.  With it, a residual-stream GEMM whose last wave of
                              256x256 tiles would be mostly empty computes its trailing rows as split-K partial planes here and
                              finishes them with a fix-up kernel (needs splits * rows * N * 4 bytes, <= 32 MB for the encoder) */
  long long workspace_bytes;
  /* ---- ABI v5 ---- */
  const void* resid_bf16;  /* bf16 [M, N] or NULL (exclusive with `resid`; bf16 `out`, may alias it): the residual stream kept in bf16 —
                              out = bf16(resid_bf16 + gate * act(acc + bias)), the sum formed in fp32 and rounded once */
  float* ln_stats_out;     /* with resid_bf16 (plain GEMM, N % 256 == 0): fp32 [M, N/128, 2] partial (sum, sum of squares) of every output
                              row, one slot per 128-column slab — the LayerNorm statistics of the NEXT op, produced for free */
  const float* ln_stats;   /* LayerNorm folded into this GEMM (nn.LayerNorm of the A rows, image_encoder.py:245,257): fp32
                              [M, ln_parts, 2] partial sums of row m of A (= a producer's ln_stats_out); A is the RAW bf16 stream, W must
                              carry gamma (W * gamma), bias must be b + W.beta, and ln_colsum[n] = sum_k (bf16(W*gamma))[n,k].
                              out = rstd_m * (acc - mu_m * ln_colsum[n]) + bias[n], then the activation.  bf16 out, no residual. */
  const float* ln_colsum;
  int ln_parts;
  float ln_eps;
} grove_gemm_epilogue;

/* out[M,N] = resid + gate * act(A[M,K] . W[N,K]^T + bias).  A, W bf16 row-major (nn.Linear layout).
 * Replaces nn.Linear / 1x1 conv calls: image_encoder.py:304 (qkv), :324 (proj), common.py:25-26 (MLP),
 * image_encoder.py:484-491 (patch embed after grove_im2col_patch16), :153-158 (neck 1x1),
 * GROVE.py:77-79 (text_hidden_fcs), transformer.py:205-208,227-229 (decoder projections).
 * Requires N % 128 == 0, K % 8 == 0, 16-byte aligned pointers. */
int grove_gemm_bf16(const void* A, const void* W, void* out, int M, int N, int K, const grove_gemm_epilogue* epi,
                    grove_stream_t stream);

/* Implicit-GEMM 'same' convolution over token-major X[V,T,G,G,C] (bf16): kt=3 -> Conv3d 3x3x3
 * (SpatioTemporalConvAdapter, image_encoder.py:40-59), kt=1 -> Conv2d 3x3 pad 1 (neck, image_encoder.py:160-166).
 * Wp[N, taps*C] is the weight repacked tap-major ((kd,)kh,kw,C).  out[V*T*G*G, N] with the same epilogue. */
int grove_conv_gemm_bf16(const void* X, const void* Wp, void* out, int V, int T, int G, int C, int N, int kt,
                         const grove_gemm_epilogue* epi, grove_stream_t stream);

/* ---- encoder element-wise / attention kernels (encoder_ops.cu, attention.cu) ------------------- */
/* images[V,3,T,H,W] bf16 ('b c t h w', GROVE.py:162) -> patches[(V*T)*(H/16)*(W/16), 768], k=(c,py,px)
 * (the im2col of PatchEmbed's Conv2d 16x16 stride 16, image_encoder.py:484-491). */
int grove_im2col_patch16(const void* images, void* patches, int V, int T, int H, int W, grove_stream_t stream);
/* y = LayerNorm(x) over the last dim D (image_encoder.py:245,257 eps 1e-6; common.py:31-43 as token-major rows).
 * x fp32 [rows,D]; y bf16 or fp32 [rows,D]; gamma/beta fp32.  D % 128 == 0, D <= 1280. */
/* the same with a bf16 input row (bf16 residual stream); statistics in fp32 */
int grove_layernorm_bf16in(const void* x, const float* gamma, const float* beta, void* y, int y_f32, int rows, int D, float eps,
                           grove_stream_t stream);
int grove_layernorm(const float* x, const float* gamma, const float* beta, void* y, int y_f32, int rows, int D, float eps,
                    grove_stream_t stream);
/* Windowed attention with decomposed rel-pos bias on the UNPARTITIONED token-major qkv[F,G,G,3,heads,hd] (bf16), tcgen05/TMEM/TMA
 * (attention_win_tc.cu): window_partition's zero padding after norm1 (image_encoder.py:245-249,344-348) is reproduced by giving pad
 * tokens k = b_k, v = b_v (qkv_bias, bf16) — they receive softmax mass like in the reference — and window_unpartition's crop (:382-383)
 * by not computing pad queries.  Persistent CTAs over (frame, window, head) units, one 4-D TMA box per operand, pad tokens patched in
 * shared memory, P kept in tensor memory.  rel_table: bf16 [64, hd] with rows 0..26 = rel_pos_h, rows 32..58 = rel_pos_w, other rows
 * zero.  out[F,G,G,heads*hd] bf16.  Replaces Attention.forward :301-326 + add_decomposed_rel_pos :420-458 on windowed blocks. */
int grove_attn_window_relpos_tc_fwd(const void* qkv, const void* qkv_bias_bf16, const void* rel_table, void* out, int F, int G, int heads,
                                    int hd, int ws, grove_stream_t stream);
/* Same, additionally writing lse[F*G*G, heads] fp32 = log2-domain log-sum-exp of every real query row (scale, bias and the pad keys of
 * its window included) -- saved by the training forward so that grove_attn_relpos_bwd_lse can skip its own log-sum-exp sweep. */
int grove_attn_window_relpos_tc_fwd_lse(const void* qkv, const void* qkv_bias_bf16, const void* rel_table, void* out, float* lse, int F, int G,
                                        int heads, int hd, int ws, grove_stream_t stream);
/* Global attention over one frame's G*G tokens with decomposed rel-pos bias (tables [2G-1, hd] bf16): tcgen05/TMEM/TMA
 * kernel (attention_tc.cu), single-pass softmax with a lazily raised running maximum, scores never leave the SM.  qkv [F,G,G,3,heads,hd], out [F,G,G,heads*hd]. */
int grove_attn_global_relpos_fwd(const void* qkv, const void* rel_pos_h, const void* rel_pos_w, void* out, int F, int G,
                                 int heads, int hd, grove_stream_t stream);
/* Same, additionally writing lse[F*G*G, heads] fp32 = log2-domain log-sum-exp of every query row (max + log2 sum, scale and bias
 * included) — saved by the training forward so that grove_attn_relpos_bwd can skip its own log-sum-exp sweep (lse may be NULL). */
int grove_attn_global_relpos_fwd_lse(const void* qkv, const void* rel_pos_h, const void* rel_pos_w, void* out, float* lse, int F, int G,
                                     int heads, int hd, grove_stream_t stream);
/* fp32 -> bf16 cast (n % 8 == 0) and token-major [F,N,C] bf16 -> NCHW [F,C,N] transposition helpers */
int grove_cast_f32_bf16(const float* x, void* y, long long n, grove_stream_t stream);
int grove_tokens_to_nchw_bf16(const void* tok, void* nchw, int F, int N, int C, grove_stream_t stream);
int grove_nchw_to_tokens_bf16(const void* nchw, void* tok, int F, int N, int C, grove_stream_t stream);
/* AdaptiveAvgPooling3D of the CLIP video features (model/llava/model/multimodal_encoder/pooling.py:6-25):
 * x [(B*T), H*W, C] token-major (bf16, or fp32 when is_f32) -> out [B, OT*OH*OW, C] (same dtype), windows as
 * nn.AdaptiveAvgPool3d((OT, OH, OW)); the reference's two einops rearranges are folded into the indexing.  C % 8 == 0. */
int grove_adaptive_avgpool3d_tokens(const void* x, void* out, int is_f32, int B, int T, int H, int W, int C, int OT, int OH, int OW,
                                    grove_stream_t stream);

/* ---- text projection / prompt encoder / box decoder (decoder_ops.cu) --------------------------- */
/* dst[i,:] = bf16(src[row_idx[i],:]) — gathers the [DET] rows BEFORE projecting them (GROVE.py:249-257 projects all
 * L tokens and gathers afterwards; identical values for the kept rows, L/P times less work). */
int grove_gather_rows_bf16(const void* src, int src_is_f32, const int* row_idx, void* dst, int n_rows, int D, grove_stream_t stream);
/* PositionEmbeddingRandom.forward (prompt_encoder.py:203-229): pe[G*G, 2*F2] fp32 token-major, gauss [2,F2] fp32. */
int grove_dense_pe(const float* gauss, float* pe, int G, int F2, grove_stream_t stream);
/* y[r,:] = bf16(x[r,:] + vec[:]) — src = image_embeddings + dense no-mask embedding (mask_decoder.py:183,
 * prompt_encoder.py:182-184), done once per FRAME: keys only become per-phrase after the first image-to-token update. */
int grove_add_rowvec_bf16(const void* x, const float* vec, void* y, long long rows, int C, grove_stream_t stream);
/* Token-to-image attention (transformer.py:231-240 inside cross_attn_token_to_image / final_attn_token_to_image):
 * q fp32 [B,T,H*dh] (already projected), k,v bf16 [*,N,H*dh] (projected keys; instance b reads row block src_of[b],
 * or b when src_of is NULL), out fp32 [B,T,H*dh] (before out_proj).  Built for T=6, dh=16.
 * lse_out (optional, fp32 [B,T,H]): log2-domain log-sum-exp of the scaled scores, saved for grove_decoder_t2i_attention_bwd. */
int grove_decoder_t2i_attention(const float* q, const void* k, const void* v, const int* src_of, float* out, float* lse_out, int B, int T,
                                int N, int heads, int dh, grove_stream_t stream);
/* Image-to-token attention (cross_attn_image_to_token, transformer.py:173-179): qi bf16 [*,N,H*dh] = q_proj(keys+pe),
 * kt, vt fp32 [B,T,H*dh]; out bf16 [B,N,H*dh] (before out_proj). */
int grove_decoder_i2t_attention(const void* qi, const float* kt, const float* vt, const int* src_of, void* out, int B, int T, int N,
                                int heads, int dh, grove_stream_t stream);
/* keys_out[b,n,:] = bf16(LayerNorm(keys_in[src_of[b],n,:] + delta[b,n,:]))  (norm4, transformer.py:180), C = 256. */
int grove_decoder_keys_add_ln(const void* keys_in, const int* src_of, const float* delta, const float* g, const float* b, void* keys_out,
                              int B, int N, int C, float eps, grove_stream_t stream);
/* ---- fused token side of a TwoWayAttentionBlock (transformer.py:151-182), ABI v4.  All weights fp32 and TRANSPOSED ([in][out]);
 * 6 tokens x 256 channels per instance, cross-attention width 128, MLP width 2048. ---- */
typedef struct grove_twoway_a_params {
  const float *wq_t, *bq, *wk_t, *bk, *wv_t, *bv, *wo_t, *bo;   /* self_attn q/k/v/out_proj, [256][256] each */
  const float *ln_g, *ln_b;                                      /* norm1 */
  float ln_eps;
  const float *wq2_t, *bq2;                                      /* cross_attn_token_to_image.q_proj, [256][128] */
  int skip_pe;                                                   /* skip_first_layer_pe: no query_pe, no residual (layer 0) */
} grove_twoway_a_params;
typedef struct grove_twoway_b_params {
  const float *wo_t, *bo;                                        /* cross_attn_token_to_image.out_proj, [128][256] */
  const float *ln2_g, *ln2_b;                                    /* norm2 */
  float ln2_eps;
  const float *w1_t, *b1, *w2_t, *b2;                            /* mlp.lin1 [256][mlp_dim], mlp.lin2 [mlp_dim][256] */
  int mlp_dim;
  const float *ln3_g, *ln3_b;                                    /* norm3 */
  float ln3_eps;
  const float *wk_t, *bk, *wv_t, *bv;                            /* cross_attn_image_to_token.k_proj / v_proj, [256][128] */
  const float *wqf_t, *bqf;                                      /* optional: q_proj of the NEXT token->image attention, [256][128] */
} grove_twoway_b_params;
/* Part A (one launch): queries fp32 [B,6,256], tokens (= query_pe) fp32 [B,6,256] ->
 *   queries_out = norm1(self_attn(q = k = queries + pe, v = queries) (+ queries)),  qt_out [B,6,128] = q_proj(queries_out + pe)
 * (transformer.py:155-164). */
int grove_twoway_block_tokens_a_fwd(const float* queries, const float* tokens, const grove_twoway_a_params* p, float* queries_out, float* qt_out,
                                    int B, int T, int C, grove_stream_t stream);
/* Part B (one launch, a cluster of 8 CTAs per instance): queries = part A's output, att fp32 [B,6,128] = grove_decoder_t2i_attention output ->
 *   x = norm2(queries + out_proj(att)); queries_out = norm3(x + lin2(relu(lin1(x))));
 *   kt_out = k_proj(queries_out + pe), vt_out = v_proj(queries_out) [B,6,128] (image->token attention, transformer.py:164-177);
 *   qf_out (optional) = wqf(queries_out + pe): the query projection of the following token->image attention (transformer.py:99-101). */
int grove_twoway_block_tokens_b_fwd(const float* queries, const float* att, const float* tokens, const grove_twoway_b_params* p,
                                    float* queries_out, float* kt_out, float* vt_out, float* qf_out, int B, int T, int C, grove_stream_t stream);
/* grove_decoder_t2i_attention with 256 threads per (instance, head) and a shuffle-first merge (same arguments and results up to fp32
 * summation order). */
int grove_decoder_t2i_attention_wide(const float* q, const void* k, const void* v, const int* src_of, float* out, float* lse_out, int B, int T,
                                     int N, int heads, int dh, grove_stream_t stream);
/* The last step of the box decoder for the prompt token `tok` (= 1 + num_mask_tokens) of every instance, one launch:
 *   hs = LayerNorm_final(queries[b,tok,:] + out_proj(att[b,tok,:]))   (final_attn_token_to_image + norm_final_attn, transformer.py:99-104)
 *   records[b] = (sigmoid(W2 relu(W0 hs + b0) + b2), Wt hs + bt)      (bbox_prediction_head / temporal_objectness_head, mask_decoder.py:80-85,191-203)
 * queries fp32 [B,T,C], att fp32 [B,T,CI] (grove_decoder_t2i_attention output), records fp32 [B,5] = (cx, cy, w, h, objectness logit) --
 * the packed per-(frame, phrase) record that replaces the reference's pickled all_gather_object (infer_iground.py:290-293).
 * Wt/bt NULL (use_temp_objectness=False) writes logit 0.  hs_out (optional, fp32 [B,C]) keeps the head input.  C = 256, CI = 128. */
int grove_decoder_heads_fwd(const float* queries, const float* att, const float* Wo, const float* bo, const float* ln_g, const float* ln_b,
                            float eps, const float* W0, const float* b0, const float* W2, const float* b2, const float* Wt, const float* bt,
                            float* records, float* hs_out, int B, int T, int tok, int C, int CI, grove_stream_t stream);
/* Token-side dense layer, all fp32: y[R,N] = act(x[R,K] . W[N,K]^T + b) (+ resid).  act: 0 none, 1 GELU, 2 ReLU, 3 sigmoid.
 * (6-token projections / MLP of transformer.py:151-182, heads mask_decoder.py:80-85,198-203, PE.W products.) */
int grove_small_linear_f32(const float* x, const float* W, const float* b, const float* resid, float* y, int R, int N, int K, int act,
                           grove_stream_t stream);
/* Self-attention among T <= 8 tokens (transformer.py:155-161): q,k,v,out fp32 [B,T,H*dh]. */
int grove_token_self_attention(const float* q, const float* k, const float* v, float* out, int B, int T, int heads, int dh,
                               grove_stream_t stream);
/* y = LN(x (+ r)) rows of fp32 [R,C] (norm1-3, norm_final_attn; eps 1e-5); optional y2 = y + add2. */
int grove_add_layernorm_f32(const float* x, const float* r, const float* g, const float* b, float* y, const float* add2, float* y2,
                            int R, int C, float eps, grove_stream_t stream);

/* ---- heads post-process, losses, box-IoU utilities (box_ops.cu) -------------------------------- */
/* (cx,cy,w,h) in (0,1) -> scale by the video's (w,h) -> xyxy; keep = sigmoid(logit) > thr
 * (GROVE.py:307-315, utils/bbox_utils.py:25-62).  size_wh fp32 [B,2] per box. */
int grove_box_postprocess(const float* boxes, const float* logits, const float* size_wh, float thr, float* xyxy, uint8_t* keep, int B,
                          grove_stream_t stream);
/* Sums for _compute_loss_components_video (GROVE.py:339-381): sel[b] marks predictions that have a GT box gt[b]
 * (cxcywh); labels[b] = objectness target.  sums[0..2] = sum GIoU loss (torchvision arithmetic, eps 1e-7), sum L1,
 * sum BCE-with-logits. */
int grove_box_losses_fwd(const float* boxes, const float* logits, const float* gt, const uint8_t* sel, const float* labels, float* sums,
                         int B, grove_stream_t stream);
/* IoU matrix out[n,m] of a[n, lda>=4] vs b[m, ldb>=4] (xyxy in the first 4 columns), fp64 when f64 != 0 else fp32,
 * bit-exact with the reference's numpy / torch arithmetic:
 *   mode 0: eval_vidstg.py:13-63 np_box_iou (no +1);  mode 1: eval_iground.py:39-56 compute_iou (+1, 0.0 on empty union);
 *   mode 2: eval_anet.py:22-119 bbox_overlaps_batch 3-D branch (+1, frm_mask[n,m] 1 = different frame (may be NULL),
 *           zero-area gt -> 0, zero-area anchor -> -1). */
int grove_box_iou(const void* a, int lda, const void* b, int ldb, const uint8_t* frm_mask, void* out, int n, int m, int mode, int f64,
                  grove_stream_t stream);
/* Greedy one-to-one matching (eval_iground.py:85-96): iou, sim fp64 [n,m] (clobbered); pairs int32 [min(n,m),2]; count int32[1]. */
int grove_greedy_match(double* iou, double* sim, double iou_thr, double sim_thr, int* pairs, int* count, int n, int m,
                       grove_stream_t stream);
/* Centre-in-box decision (eval_youcookinteractions.py:43-48): pred, gt fp64 [n,4] xyxy; correct[i] = 1 iff the centre of pred[i]
 * (Python-float arithmetic, (x1+x2)/2) lies inside gt[i] with INCLUSIVE bounds. */
int grove_center_in_box(const double* pred, const double* gt, uint8_t* correct, int n, grove_stream_t stream);
/* Video IoU and recall decisions of one video (eval_vidstg.py:157-178): pred, gt fp64 [n,4] (one row per ground-truth frame, in frame
 * order); ious[i] = np_box_iou (mode 0) or 0 when pred[i] is all zeros; viou[0] = (sum in frame order) / max(n,1);
 * over[j] = viou > thr[j] (strict) for the k thresholds. */
int grove_viou_decisions(const double* pred, const double* gt, const double* thr, int n, int k, double* ious, double* viou, uint8_t* over,
                         grove_stream_t stream);
/* Validation metrics (train.py:826-835): giou_sum[0] = sum over rows with sel[i] of torchvision's GIoU loss of boxes[i] vs gt[i] ON THE
 * COORDINATES AS GIVEN (fp32; the reference passes cxcywh there -- reproduced); acc[0] = #{i : (sigmoid(logits[i]) > 0.5) == labels[i]}.
 * giou_each (optional, fp32 [B]) receives the per-row loss (0 where not selected). */
int grove_val_metrics(const float* boxes, const float* logits, const float* gt, const uint8_t* sel, const int* labels, double* giou_sum, int* acc,
                      float* giou_each, int B, grove_stream_t stream);


/* ==== frame pre-processing fused into the patch embed's operand ===============================
 * ResizeLongestSide.apply_image (model/SAM/utils/transforms.py:27-34 — PIL bilinear, bit-exact restatement of Pillow's
 * Resample.c: 22-bit fixed-point coefficients, horizontal then vertical pass, uint8 in between) followed by
 * grounding_enc_processor (HowTo100M.py:168-178: (x - mean) / std, zero pad) and .bfloat16() (train.py:751-753).
 * bounds [out, 2] = (first source index, tap count), coeffs [out, ksize] int32: built by the host exactly like precompute_coeffs. */
/* horizontal pass over `rows` image rows: in [rows, w_in, 3] u8 -> out [rows, w_out, 3] u8 */
int grove_resize_rows_u8(const uint8_t* in, uint8_t* out, const int* bounds, const int* coeffs, int ksize, long long rows, int w_in, int w_out,
                         grove_stream_t stream);
/* vertical pass + normalise + pad + patchify: in [F, h_in, w, 3] u8 -> patches [F*(img/16)^2, 768] bf16 (k = c*256 + py*16 + px), the A
 * operand of the PatchEmbed GEMM (image_encoder.py:484-491).  Rows >= h_out and columns >= w are the zero padding; identity
 * coefficients (count 1, 1 << 22) when the height is kept.  mean3 / std3: HOST pointers to three floats. */
int grove_frames_to_patches_u8(const uint8_t* in, const int* vbounds, const int* vcoeffs, int vksize, void* patches, int F, int h_in, int w,
                               int h_out, int img, const float* mean3, const float* std3, grove_stream_t stream);

/* ==== training step of the grounding branch (BASELINE config 4; SURVEY.md §8a row "4-bwd") ========================
 * The reference trains this branch through autograd (train.py:761-782: model(**batch); model.backward(loss)); trainable
 * there: the Conv3d adapters, the whole mask decoder with its heads, text_hidden_fcs (train.py:279-296) — the ViT blocks
 * are frozen but sit between the adapters, so activations' gradients flow through every block after the first adapter.
 * The contractions of the backward pass reuse grove_gemm_bf16 / grove_conv_gemm_bf16 with transposed / flipped weights
 * (input gradients) and transposed activations (weight gradients, K = tokens, optional split-K). */

/* dWp[N, taps*C] fp32 (tap-major, [splits, ...] partial planes when splits > 1) = sum_tokens dY[token, n] * X[token + shift(tap), c]:
 * weight gradient of grove_conv_gemm_bf16 (Conv3d adapter, image_encoder.py:40-59).  dYt = dY transposed [N, tokens] bf16,
 * Xt3 = X channel-major in three w-shifted planes [3, C, V,T,G,G] bf16 (grove_transpose_shift3_to_bf16; a TMA box cannot start at an
 * odd element of the innermost dimension); the B operand is a 5-D TMA box whose origin carries the tap's (h,t) shift. */
int grove_conv_wgrad_bf16(const void* dYt, const void* Xt3, float* dWp, int V, int T, int G, int C, int N, int kt, int splits,
                          grove_stream_t stream);
/* out[C,R] bf16 = transpose(in[R,C]) (in fp32 or bf16): operands of the weight-gradient GEMMs */
int grove_transpose_to_bf16(const void* in, int in_is_f32, void* out, int R, int C, grove_stream_t stream);
/* out[3, C, R] bf16: plane d = transpose(in[R,C]) shifted by d-1 along the grid's w axis (R = frames*G*G, zero outside [0,G)) */
int grove_transpose_shift3_to_bf16(const void* in, int in_is_f32, void* out, int R, int C, int G, grove_stream_t stream);
/* out[i] = (accumulate ? out[i] : 0) + scale * sum_s partials[s, i]: finishes a split-K GEMM */
int grove_reduce_partials_f32(const float* partials, int splits, long long n, float* out, int accumulate, float scale, grove_stream_t stream);
/* LayerNorm backward over rows of u = x (+ r) [fp32], or u = keys[src_of[row/N]*N + row%N] (bf16) + r (fp32) when x_is_keys_bf16
 * (norm4 of the two-way block, transformer.py:180).  dx_out = (dx_in ? dx_in : 0) + dLN(dy); optional bf16 copy; optional
 * d gamma / d beta (+=, atomics).  D % 128 == 0, D <= 1280. */
int grove_layernorm_bwd(const void* x, const float* r, const int* src_of, int N, int x_is_keys_bf16, const float* gamma, const void* dy,
                        int dy_is_f32, const float* dx_in, float* dx_out, void* dx_bf16, float* dgamma, float* dbeta, long long rows, int D,
                        float eps, grove_stream_t stream);
/* Adapter gate backward (image_encoder.py:54): dyc = dy * tanh(alpha) * [relu_out > 0] (bf16); dbias += colsum(dyc);
 * dalpha += (1 - tanh^2(alpha)) * sum(dy * relu_out).  relu_out = relu(conv + b) saved by the forward GEMM (out2_pre_act = 2). */
int grove_adapter_gate_bwd(const float* dy, const void* relu_out, const float* alpha, void* dyc, float* dbias, float* dalpha, long long rows,
                           int D, grove_stream_t stream);
/* out[c] += sum_r x[r,c] (bias gradients) ; out[f,:] = (+=) sum_{b in [off[f],off[f+1])} x[b,:] (phrases sharing a frame's keys) */
int grove_colsum(const void* x, int x_is_f32, float* out, long long R, int C, grove_stream_t stream);
int grove_segment_sum_f32(const float* x, const int* offsets, float* out, int segments, long long n, int accumulate, grove_stream_t stream);
/* token-side fp32: dW[N,K] += dY[R,N]^T . X[R,K] ; dx = dy * act'(y) (kind 2 ReLU, 3 sigmoid, from the activation's OUTPUT y) */
int grove_small_wgrad_f32(const float* dy, const float* x, float* dw, int R, int N, int K, grove_stream_t stream);
int grove_act_bwd_f32(const float* dy, const float* y, float* dx, long long n, int kind, grove_stream_t stream);
int grove_token_self_attention_bwd(const float* q, const float* k, const float* v, const float* dout, float* dq, float* dk, float* dv, int B,
                                   int T, int heads, int dh, grove_stream_t stream);
/* Cross-attention backward of the two-way block (transformer.py:164-180, 99-104); layouts as the forward entry points.
 * t2i: dk, dv bf16 [B,N,H*dh] per instance.  i2t: dkt, dvt fp32 [B,T,H*dh] must be zero-initialised (accumulated). */
int grove_decoder_t2i_attention_bwd(const float* q, const void* k, const void* v, const int* src_of, const float* att, const float* datt,
                                    const float* lse, float* dq, void* dk, void* dv, int B, int T, int N, int heads, int dh,
                                    grove_stream_t stream);
int grove_decoder_i2t_attention_bwd(const void* qi, const float* kt, const float* vt, const int* src_of, const void* dout, void* dqi, float* dkt,
                                    float* dvt, int B, int T, int N, int heads, int dh, grove_stream_t stream);
/* out[n] fp32 = sum_b x[b, n] (bf16) */
int grove_batch_sum_bf16(const void* x, float* out, int B, long long n, grove_stream_t stream);
/* d(qkv) of Attention.forward + add_decomposed_rel_pos (image_encoder.py:301-326, 420-458) for frozen blocks: ws = 14 windowed
 * (on the unpartitioned tensors, qkv_bias_bf16 = the k/v of window_partition's zero-padded tokens) or ws = 0 global.
 * att = the forward output (before proj), datt its gradient, both bf16 [F,G,G,heads*hd]; dqkv bf16 [F,G,G,3,heads,hd].
 * rel_pos_h/w: bf16 [2S-1, hd] with S = ws or G.  workspace: grove_attn_relpos_bwd_workspace_bytes() bytes, 16-byte aligned. */
long long grove_attn_relpos_bwd_workspace_bytes(int F, int G, int heads, int hd, int ws);
int grove_attn_relpos_bwd(const void* qkv, const void* qkv_bias_bf16, const void* rel_pos_h, const void* rel_pos_w, const void* att,
                          const void* datt, void* dqkv, void* workspace, int F, int G, int heads, int hd, int ws, grove_stream_t stream);
/* Same with the forward kernel's log-sum-exp (grove_attn_global_relpos_fwd_lse / grove_attn_window_relpos_tc_fwd_lse; fp32
 * [F*G*G, heads], log2 domain, may be NULL).  With it the backward runs on tcgen05 / TMEM: global layers on 32x32 / 64x64 grids as a
 * query-side and a key-side kernel (attention_bwd_tc.cu; the query side also computes the bias rows and back-projects the bias
 * cotangents through the rel-pos tables), windowed layers as one persistent kernel over (frame, window, head) units
 * (attention_win_bwd_tc.cu); head dims 64 and 80.  Without it (or on 16x16 grids) the warp-level kernels of attention_bwd.cu run.
 * rel_pos_h / rel_pos_w must then be 16-byte aligned (TMA). */
int grove_attn_relpos_bwd_lse(const void* qkv, const void* qkv_bias_bf16, const void* rel_pos_h, const void* rel_pos_w, const void* att,
                              const void* datt, void* dqkv, void* workspace, const float* lse_fwd, int F, int G, int heads, int hd, int ws,
                              grove_stream_t stream);
/* d boxes [B,4] (cxcywh) and d logits [B] of the loss of _compute_loss_components_video (GROVE.py:339-381):
 * cg = upstream * giou_weight / (n_gt + 1e-8) (GIoU and L1 share it, GROVE.py:375), co = upstream * objectness_weight / (n_pred + 1e-8). */
int grove_box_losses_bwd(const float* boxes, const float* logits, const float* gt, const uint8_t* sel, const float* labels, float cg, float co,
                         float* dboxes, float* dlogits, int B, grove_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GROVE_B200_H_ */
