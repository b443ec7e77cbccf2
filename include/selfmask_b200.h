/*
 * selfmask_b200 — C-ABI of the B200-native SelfMask inference + evaluation hot path.
 *
 * Drop-in boundary.  Every entry point takes plain device pointers, sizes and a
 * CUDA stream (passed as void* = cudaStream_t), returns 0 on success or a negative smk_status, never
 * throws, never allocates device memory (the caller owns every buffer, including the workspace whose
 * size it queries first) and never synchronises the stream.  One host thread per GPU/process.
 *
 * What each group replaces in the reference (paths under /root/reference):
 *   smk_model_*        networks/maskformer/maskformer.py:164-251  MaskFormer.forward
 *                      (networks/vision_transformer.py:269-304 encoder,
 *                       networks/maskformer/transformer_decoder.py:112-150,260-297 decoder,
 *                       maskformer.py:144-162,223 pixel decoder + mask einsum + sigmoid,
 *                       maskformer.py:229-239,254-268 objectness MLP), reached through
 *                      base_structure.py:18-24 BaseStructure._forward
 *   smk_eval_batch     evaluator.pyc@L199-226 (last-layer slice, x4 bilinear, upper-bound query,
 *                      objectness top-1) + the integer/float reductions behind
 *                      metrics/iou.py:22-31, f_measure.py:24-81, mae.py:9, pixel_acc.py:10-14,
 *                      s_measure.py:11-124
 *   smk_upsample_bilinear  evaluator.pyc@L209-211  F.interpolate(scale_factor=4, 'bilinear')[..., :h, :w]
 *   smk_mask_metrics   the same reductions for one full-resolution mask (the five metric callables)
 *   smk_gemm_*, smk_layernorm, smk_attention, ...   single kernels, exported for unit tests
 *
 * Mapping to the export list sketched in SURVEY.md §8b (the survey proposed per-stage entry points; the library keeps the stages in
 * ONE forward call so that no intermediate tensor crosses the boundary, and exposes them through taps and single-kernel entries):
 *   smk_encoder_forward(h, x, tokens_out)          → smk_model_forward(..., mask_pred = objectness = features = NULL) + smk_model_tap(h, 1, tokens_out)
 *                                                     (the Python mirror's `model(x, encoder_only=True)`)
 *   smk_decoder_forward(h, tokens, queries, obj)    → inside smk_model_forward; queries via smk_model_tap(h, 2, ...), objectness is an output
 *   smk_mask_head(h, queries, tokens, mask_pred)    → inside smk_model_forward (all_layers selects 6-layer / last-layer masks); the
 *                                                     contraction alone is smk_gemm_batched
 *   smk_upsample_sigmoid / smk_hist_metrics / smk_smeasure → fused in smk_eval_batch (x4 upsample + every metric reduction in two
 *                                                     launches) and, for one full-resolution mask, smk_mask_metrics; the upsample alone
 *                                                     is smk_upsample_bilinear
 *   smk_layernorm, smk_gemm_*, smk_attention*       → exported as sketched (smk_gemm_tokens = the patch-embed form with token assembly,
 *                                                     smk_gemm_q8 / smk_split_q8 = the fp8-corrected split form and its operand layout,
 *                                                     smk_attention_tc_multi = attention beyond one key tile)
 *
 * Threading contract: one host thread drives a device at a time.  Per-device one-time state (kernel attributes, SM count, the
 * persisting-L2 carve-out) is keyed by the CUDA device ordinal, so one process may use several GPUs in turn; two host threads on
 * the SAME device must serialise their calls (the optional event profiler and launch counter are process-wide).
 */
#ifndef SELFMASK_B200_H_
#define SELFMASK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SMK_OK = 0,
  SMK_ERR_INVALID = -1,   /* bad argument / unsupported geometry */
  SMK_ERR_CUDA = -2,      /* a CUDA call failed; see smk_last_error() */
  SMK_ERR_WORKSPACE = -3, /* workspace too small */
  SMK_ERR_UNSUPPORTED = -4
} smk_status;

/* numeric modes */
enum { SMK_MODE_FP32 = 0,  /* validation: every contraction in fp32 on CUDA cores          */
       SMK_MODE_BF16 = 1,  /* throughput: tcgen05 bf16 operands, fp32 accumulate/residual/LN */
       SMK_MODE_BF16X3 = 2,/* parity on tensor cores: every GEMM as a 3-term bf16 split (hi·hi + hi·lo + lo·hi, K' = 3K) on
                              tcgen05 with fp32 accumulate; LayerNorm / softmax attention / residual in fp32 */
       SMK_MODE_FP16S = 3  /* parity AND throughput (the benchmarked mode): fp16 tcgen05 operands with as many split terms per
                              contraction as the 2e-2 logit budget needs (patch embed / proj / fc1 / fc2: hi·hi in fp16 + the two correction
                              products hi·lo + lo·hi on e4m3 operands at the fp8 tensor rate, smk_gemm_q8; qkv, memory K/V: A_hi·(W_hi + W_lo);
                              encoder attention: single-pass fp16 on tcgen05; decoder self-attention fp32; decoder tail as in
                              SMK_MODE_BF16), fp32 accumulate / residual / LayerNorm / softmax */ };

typedef struct {
  int32_t patch;        /* 16 (or 8)                                    */
  int32_t dim;          /* 384                                          */
  int32_t depth;        /* 12 encoder blocks                            */
  int32_t heads;        /* 6                                            */
  int32_t mlp_dim;      /* 1536                                         */
  int32_t n_queries;    /* 10 / 20                                      */
  int32_t dec_layers;   /* 6                                            */
  int32_t dec_ffn;      /* 1536                                         */
  int32_t scale_factor; /* pixel-decoder bilinear factor (4 for ViT-S/16) */
  int32_t pos_grid;     /* side of the learned position grid (14)       */
} smk_config;

const char* smk_last_error(void);
int smk_version(void);

/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t smk_launch_count(void);
/* per-category device timing with CUDA events on the launching stream.  enable: 1 start recording (resets), 0 stop.
 * smk_prof_read synchronises the recorded events and fills ms[8], work[8] (algorithmic FLOPs or bytes), launches[8];
 * categories: 0 tcgen05 GEMM, 1 CUDA-core attention, 2 fp32 GEMM, 3 LayerNorm, 4 evaluation, 5 mask head, 6 other,
 * 7 tcgen05 attention. */
int smk_prof_enable(int enable);
int smk_prof_read(double* ms, double* work, int64_t* launches);
/* per-launch timeline of the recorded launches in launch order (tuning aid): duration, category and start time relative to the
 * first recorded launch; returns the number of entries written (<= cap) or a negative status */
int smk_prof_timeline(float* ms, int* cat, float* start_ms, int cap);
/* the same with each launch's sub-category tag (1 patch embed, 2 qkv, 3 encoder attention, 4 proj, 5 fc1, 6 fc2, 7 encoder LayerNorm,
 * 8 memory K/V, 9 decoder GEMM, 10 decoder attention, 11 decoder LayerNorm, 12 mask logits, 13 mask upsample, 14 objectness,
 * 15 per-query IoU, 16 mask metrics, 17 im2col, 0 other), its ALGORITHMIC work (FLOPs or bytes) and the work it issued (split
 * GEMMs issue 2-3 tensor-core terms per algorithmic FLOP) */
int smk_prof_timeline2(float* ms, int* cat, int* tag, double* work, double* issued, int cap);

/* ---- weights: one fp32 device blob in a canonical order --------------------------------------
 * The table maps the reference's state_dict keys (SURVEY.md §8b, 267 tensors) to blob offsets. */
int smk_weight_count(const smk_config* cfg);
int smk_weight_entry(const smk_config* cfg, int index, char* name, int name_cap, int64_t* offset, int64_t* numel);
int64_t smk_weights_numel(const smk_config* cfg);

/* ---- model ----------------------------------------------------------------------------------- */
typedef struct smk_model smk_model;

/* bytes of persistent device memory the model needs for repacked weights + activations at
 * (max_batch, img_h, img_w) in `mode`. */
int64_t smk_model_workspace_bytes(const smk_config* cfg, int mode, int max_batch, int img_h, int img_w);

/* `weights` is the fp32 blob (device, must stay alive); `workspace` is caller-owned device memory.
 * Repacks weights (bf16 copies, position-embedding resample, constant-folded decoder layer-0
 * self-attention) on `stream`. */
int smk_model_create(const smk_config* cfg, int mode, const float* weights, void* workspace, int64_t workspace_bytes,
                     int max_batch, int img_h, int img_w, void* stream, smk_model** out);
int smk_model_destroy(smk_model* m);

/* x [B,3,H,W] fp32 NCHW (ImageNet-normalised).  Outputs (any may be NULL):
 *   mask_pred  [B, L, nq, h', w'] fp32 probabilities, L = dec_layers if all_layers else 1 (last)
 *   objectness [B, L, nq]         fp32 probabilities
 *   features   [B, dim]           mean over queries of the last decoder layer
 * h' = ceil(H/patch)*scale_factor. */
int smk_model_forward(smk_model* m, const float* x, int B, int H, int W, int all_layers,
                      float* mask_pred, float* objectness, float* features, void* stream);

/* Same, from raw pixels: x [B,3,H,W] uint8 NCHW; the reference's host-side normalisation
 * (datasets/base_dataset.py:250, torchvision to_tensor + normalize: ((float)u / 255 - mean[c]) / std[c]) is fused into
 * the patch im2col with IEEE fp32 division, bit-identical to smk_model_forward on the host-normalised image.
 * mean_std: HOST pointer to 6 floats {mean[3], std[3]}. */
int smk_model_forward_u8(smk_model* m, const uint8_t* x, const float* mean_std, int B, int H, int W, int all_layers,
                         float* mask_pred, float* objectness, float* features, void* stream);

/* debug taps for stage-level parity tests: copies an internal activation (fp32) into `out`.
 * what: 1 final-LN encoder tokens [B,N,D]; 2 decoder queries after the shared final norm [L,B,nq,D];
 *       3 residual stream after the last encoder block [B,N,D] */
int smk_model_tap(smk_model* m, int what, float* out, int64_t out_numel, void* stream);
/* when set (non-NULL), the next forward passes also store the pre-sigmoid mask logits, same shape as mask_pred */
int smk_model_debug_logits(smk_model* m, float* logits);

/* ---- evaluation ------------------------------------------------------------------------------ */
#define SMK_QCOUNT_STRIDE 2      /* per query: intersection, union at threshold 0.5            */
#define SMK_MCOUNT_STRIDE 528    /* per evaluated mask, int32: see layout below                */
#define SMK_MSUM_STRIDE 32       /* per evaluated mask, double                                 */
/* int32 layout of one evaluated mask:
 *   [0,256)   histogram over foreground pixels (gt==1) of bin(p) = #{k : t_k < p}, t_k = float(k/255)
 *   [256,512) the same over background pixels
 *   512 tp@0.5   513 (tp+fp)@0.5   514 sum(gt)   515 tp@tau   516 (tp+fp)@tau   (tau = 2*mean(p))
 *   517 centroid X   518 centroid Y   519 H*W   520 query index of this mask
 * double layout:
 *   0 sum p   1 sum |p-g|   2 tau   3 fg sum p   4 fg sum p^2   5 bg sum (1-p)   6 bg sum (1-p)^2
 *   8+5*q+{0..4}, q = LT,RT,LB,RB: pixel count, sum p, sum p^2, sum g, sum p*g                    */

/* mask_pred: probabilities at mask resolution, element (b,q,y,x) at
 * mask_pred[b*batch_stride + q*hp*wp + y*wp + x]; objectness (b,q) at objectness[b*obj_stride + q];
 * gt uint8 {0,1} [B,H,W] with H <= hp*up, W <= wp*up (crop, evaluator.pyc@L211).
 * Outputs: q_counts int32 [B,nq,2]; idx int32 [B,2] = (objectness top-1, upper-bound query);
 *          m_counts int32 [B,2,528]; m_sums double [B,2,32]  (index 0 = selected, 1 = upper bound). */
int smk_eval_batch(const float* mask_pred, int64_t batch_stride, const float* objectness, int64_t obj_stride,
                   const uint8_t* gt, int B, int nq, int hp, int wp, int up, int H, int W,
                   int32_t* q_counts, int32_t* idx, int32_t* m_counts, double* m_sums, void* stream);

/* one or more full-resolution masks [n,H,W] fp32 against gt uint8 [n,H,W] → m_counts [n,528], m_sums [n,32] */
int smk_mask_metrics(const float* pred, const uint8_t* gt, int n, int H, int W,
                     int32_t* m_counts, double* m_sums, void* stream);

/* Records → metric values on the device, bit-identical to the host finalisation (metrics.py::finalize, i.e. iou.py:31,
 * pixel_acc.py:10-14, f_measure.py:24-81, mae.py:9 in float32 and s_measure.py:108-124 in float64).
 * m_counts int32 [n_masks,528], m_sums double [n_masks,32] → out double [n_masks,8] =
 * {iou, pixel_acc, F@0.5, F-max, F@2·mean, MAE (each an exact float32 value), S-measure (float64), 0}.
 * c1 = float32(1 + beta_square**2), c2 = float32(beta_square**2), eps = float32(1e-7) (f_measure.py:49,80). */
int smk_finalize_records(const int32_t* m_counts, const double* m_sums, int64_t n_masks, float c1, float c2, float eps,
                         double* out, void* stream);

/* in [n,h,w] fp32 → out [n,H,W], H <= h*scale, W <= w*scale; ATen bilinear, align_corners=False */
int smk_upsample_bilinear(const float* in, float* out, int64_t n, int h, int w, int scale, int H, int W, void* stream);

/* ---- single kernels (unit tests) ------------------------------------------------------------- */
/* epilogue flags */
enum { SMK_EPI_NONE = 0, SMK_EPI_GELU = 1, SMK_EPI_RELU = 2, SMK_EPI_RESIDUAL = 4 /* C += existing C (fp32) */ };

/* C[M,N] = A[M,K] · W[N,K]^T + bias[N], fp32 on CUDA cores.  lda/ldc in elements. */
int smk_gemm_f32(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc,
                 int M, int N, int K, int epilogue, void* stream);
/* bf16 tcgen05 GEMM: A [M,K] bf16 (lda), W [N,K] bf16; fp32 accumulate.  out_f32: 0 = bf16 output, 1 = fp32 output,
 * 2 = bf16x3 split output (row = [hi | hi | lo], 3N columns, ldc >= 3N: the A operand of a following split GEMM).
 * K % 64 == 0, N % 128 == 0. */
int smk_gemm_bf16(const void* A, int64_t lda, const void* W, const float* bias, void* C, int64_t ldc,
                  int M, int N, int K, int epilogue, int out_f32, void* stream);
/* Split-operand tcgen05 GEMM: C[M,N] = sum over t < n_terms of A[:, a_off[t] : a_off[t]+K] · W[:, w_off[t] : w_off[t]+K]^T + bias.
 * A [M, lda], W [N, ldw]: 16-bit operands (bf16, or fp16 when f16 != 0), offsets in elements (multiples of 64).  With operands
 * stored as [hi | lo] rows, terms (0,0),(0,K),(K,0) give the ~fp32 3-term product, (0,0),(0,K) removes the weight rounding only.
 * out_kind: 0 16-bit (same type as the operands), 1 fp32, 2 split [hi | hi | lo] (3N columns), 3 split [hi | lo] (2N columns). */
int smk_gemm_split(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                   int epilogue, int out_kind, int f16, int n_terms, const int32_t* a_off, const int32_t* w_off, void* stream);
/* fp16 product with fp8 correction terms (the fp16s mode's patch-embed / proj / fc1 / fc2; vision_transformer.py:88-94,131,184-188):
 * C = A_hi·W_hi^T + (e4m3(A_hi)·e4m3(W_lo·2^15)^T + e4m3(A_lo·2^11)·e4m3(W_hi·2^4)^T)·2^-15 + bias — the accuracy of the 3-term fp16 split
 * with the two correction products at the fp8 tensor-core rate.  A [M, lda >= 2K], W [N, ldw >= 2K]: rows as written by smk_split_q8
 * ([hi fp16 (K) | 2K bytes of e4m3 operands]).  out_kind as smk_gemm_split, plus 4 = [hi fp16 | e4m3 operands] (2N fp16 columns: the A
 * operand of the next smk_gemm_q8).  K % 64 == 0, N % 128 == 0. */
int smk_gemm_q8(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                int epilogue, int out_kind, void* stream);
/* Patch-embed form of the GEMM (vision_transformer.py:184-188, :269-287): C is [n_img, hw + 1, N] token rows (row stride ldc); GEMM row m =
 * patch m % hw of image m / hw is written to token row 1 + m % hw with pos[1 + m % hw, :] (pos [hw + 1, N] fp32) and the bias added;
 * the class-token rows are not touched.  A [n_img*hw, lda], W [N, ldw]: bf16 (f16 = 0), fp16 (1) or smk_split_q8 rows (2). */
int smk_gemm_tokens(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* pos, float* C, int64_t ldc,
                    int n_img, int hw, int N, int K, int f16, void* stream);
/* x [rows, ldx] fp32 → out [rows, 2K fp16 columns]: [hi fp16 (K) | per 32 columns 32 bytes "first" + 32 bytes "second" of e4m3];
 * activations (is_weight == 0): first = e4m3(hi), second = e4m3(lo·2^11); weights: first = e4m3(lo·2^15), second = e4m3(hi·2^4). */
int smk_split_q8(const float* x, int64_t ldx, void* out, int64_t rows, int K, int is_weight, void* stream);
/* Batched form: C_b [rows_a, rows_w] fp32 = sum over terms of A_b · W_b^T for b < n_batch, where A_b = rows
 * [b*batch_a_rows + a_row0, +rows_a) of A [a_total_rows, lda] and W_b = rows [b*batch_w_rows + w_row0, +rows_w) of W [w_total_rows, ldw];
 * C [n_batch, rows_a, rows_w] contiguous, rows_w % 4 == 0.  The mask-logit contraction (maskformer.py:223 at patch resolution):
 * A = the decoder queries of image b (all layers), W = its patch tokens. */
int smk_gemm_batched(const void* A, int64_t lda, int64_t a_total_rows, int batch_a_rows, int a_row0, int rows_a, const void* W, int64_t ldw,
                     int64_t w_total_rows, int batch_w_rows, int w_row0, int rows_w, float* C, int n_batch, int K, int f16, int n_terms,
                     const int32_t* a_off, const int32_t* w_off, void* stream);
/* y = LN(x) * gamma + beta over the last dim D (fp32 statistics).  out_bf16 selects the output type. */
int smk_layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int D, float eps,
                  int out_bf16, void* stream);
/* The fp16s mode's LayerNorm (vision_transformer.py:150-163 norm1 / norm2, 1e-6): y [rows, ldy] fp16 = fp16(LN(x)) in columns [0, D);
 * lo_kind 1 adds the fp16 rounding residue in columns [D, 2D) ([hi | lo] operand rows), lo_kind 2 the e4m3 correction operands of
 * smk_split_q8 in the same 2D bytes (the A operand of smk_gemm_q8); y32 (optional) receives the fp32 normalised rows. */
int smk_layernorm_f16(const float* x, const float* gamma, const float* beta, void* y, int64_t ldy, float* y32, int64_t rows, int D,
                      float eps, int lo_kind, void* stream);
/* Patch im2col of the fp16s mode (vision_transformer.py:184-188 as a GEMM): x [B, 3, H, W] fp32, or raw uint8 pixels normalised on the
 * fly with mean_std (host, 6 floats: ((u/255) - mean) / std in IEEE fp32, datasets/base_dataset.py:250) → cols [B*hp*wp, 2*3*P*P fp16
 * columns]: [hi | lo] fp16, or [hi | e4m3 operands] when q8 != 0.  Images are zero-padded to multiples of P (:260-267). */
int smk_im2col_f16(const void* x, int is_u8, void* cols, int B, int H, int W, int P, const float* mean_std, int q8, void* stream);
/* softmax(scale * Q K^T) V for `batch` problems × `heads`, fp32 math.
 * Q row (b,i) at q + b*q_bstride + i*ldq, head h in columns [h*dh, h*dh+dh); same for k, v, o. */
int smk_attention(const void* q, const void* k, const void* v, void* o, int batch, int heads, int dh, int Lq, int Lk,
                  int64_t q_bstride, int64_t ldq, int64_t k_bstride, int64_t ldk, int64_t v_bstride, int64_t ldv,
                  int64_t o_bstride, int64_t ldo, float scale, int is_bf16, void* stream);
/* tcgen05 fused self-attention on the fused-QKV layout: qkv [B*N, 3*heads*64] bf16 (q|k|v), out [B*N, heads*64] bf16;
 * N <= 256 tokens per image (one key tile). */
int smk_attention_tc(const void* qkv, void* out, int B, int N, int heads, float scale, void* stream);
/* fp16 form of smk_attention_tc (fp16s mode): qkv / out fp16; out_mode 0 → out [B*N, ldo >= heads*64],
 * 3 → [hi | lo] split rows (ldo >= 2*heads*64): the A operand of a 3-term smk_gemm_split */
int smk_attention_tc_f16(const void* qkv, void* out, int64_t ldo, int out_mode, int B, int N, int heads, float scale, void* stream);
/* Multi-key-tile form for sequences longer than one tile (384 x 384: 577 tokens; ViT-S/8: 785; vision_transformer.py:110-130):
 * 176-key tiles with an online rescale of the TMEM accumulator.  q / k / v: 16-bit matrices (bf16, or fp16 when f16 != 0), head h at
 * columns [h*64, h*64+64) from each pointer; image b: queries at rows b*q_rows .. +Lq, keys / values at rows b*kv_rows + kv_row0 .. +Lk.
 * Lk >= 176.  out [.., ldo]: out_mode 0 16-bit (operand type), 1 fp32, 2 [hi | hi | lo], 3 [hi | lo], 4 [hi fp16 | e4m3 operands] (the
 * A operand of smk_gemm_q8; f16 only); out_mode | 8 with f16: the parts of modes 2 / 3 are bf16 (consumer: a bf16 split GEMM). */
int smk_attention_tc_multi(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t q_total_rows,
                           int64_t kv_total_rows, int q_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq,
                           int Lk, int heads, float scale, int f16, void* stream);
/* general form: q [B*Lq, ldq], k / v [kv_total_rows, ld] bf16 (head h at columns [h*64, h*64+64) of each pointer); image b's
 * queries start at row b*Lq, its keys/values at row b*kv_rows + kv_row0; out [B*Lq, ldo]: out_f32 0 = bf16, 1 = fp32,
 * 2 = bf16x3 split ([hi | hi | lo], 3*heads*64 columns). */
int smk_attention_tc_general(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t kv_total_rows,
                             int kv_rows, int kv_row0, void* out, int64_t ldo, int out_f32, int B, int Lq, int Lk, int heads,
                             float scale, void* stream);
/* tuning aid: phase trace of CTA 0 of the tcgen05 attention kernel ([16 items][10 warps][8 events] clock64 stamps in device
 * memory); NULL switches it off. */
int smk_debug_attn_trace(long long* buf);
int smk_debug_xattn_trace(long long* buf);  /* tuning aid: phase trace of CTA 0 of the cross-attention kernel ([8 images][2 roles][8 events]) */
int smk_debug_gemm_trace(long long* buf);   /* tuning aid: per-CTA wait-cycle counters of the tcgen05 GEMM (16 per CTA); NULL = off */
/* Few-query attention (decoder self / cross attention, transformer_decoder.py:271-291): Lq <= 32 queries and Lk <= 256 keys per
 * (image, head), head dim 64, bf16 inputs.  Image b: queries at rows b*Lq.., keys / values at rows b*kv_rows + kv_row0 ..
 * out_mode: 0 bf16 [B*Lq, heads*64], 1 fp32, 2 bf16x3 split [hi | hi | lo] (3*heads*64 columns). */
int smk_attention_small(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int kv_rows,
                        int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq, int Lk, int heads, float scale, void* stream);

/* the same with fp16 k / v; q fp16, or fp32 rows (ldq in floats) when q_f32 != 0 — rounded to fp16 as it is staged.  Outputs as above. */
int smk_attention_small_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int kv_rows,
                            int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq, int Lk, int heads, float scale,
                            int q_f32, void* stream);
/* Decoder self-attention in fp32 on the CUDA cores (nq <= 32 queries = keys per image): qk [B*nq, ldqk] fp32 with q in columns
 * [0, heads*64) and k in [heads*64, 2*heads*64), v [B*nq, ldv] fp32 → out3 [B*nq, 3*heads*64] bf16 split [hi | hi | lo].
 * v_sub: NULL, or [nq, heads*64] fp32 subtracted from every image's V rows (v then comes from ONE q|k|v projection of tgt + query_pos and
 * v_sub = query_pos · Wv^T removes the positional part: value = tgt, transformer_decoder.py:277). */
int smk_dec_self_attention(const float* qk, int64_t ldqk, const float* v, int64_t ldv, const float* v_sub, void* out3, int B, int nq,
                           int heads, float scale, void* stream);

/* Decoder cross-attention against the memory, restructured (smk_xattn_tc.cu; transformer_decoder.py:283-291): per image the nq*heads
 * rows of qp [n_img*nq, heads*D] fp16 (row (b, q), column (h, c): the query times the folded Wq_h^T Wk_h / 8) attend the image's hw
 * final-LN patch tokens tok [n_img*t_rows_per_img, D] fp16 (rows b*t_rows_per_img + t_row0 ..) as keys AND values, on tcgen05:
 * out [n_img*nq, 2*heads*D] fp16 split [hi | lo] = softmax(qp_h tok^T) tok per head h.  nq*heads <= 128, hw <= 208, D <= 384. */
int smk_xattn_tc(const void* qp, const void* tok, int t_rows_per_img, int t_row0, void* out, int n_img, int nq, int heads, int D, int hw,
                 void* stream);
/* the weight folding that goes with it, from one decoder layer's multihead_attn in_proj_weight [3D, D] / in_proj_bias [3D] /
 * out_proj.weight [D, D] / out_proj.bias [D]:  wg [heads*D, D] fp16 = (Wq_h^T Wk_h / 8) as GEMM weight rows (h, c), g [heads*D] its
 * bias (bq_h Wk_h / 8), mcat [D, heads*D] fp32 = Wo_h Wv_h per head block, bo2 [D] = Wo bv + bo */
int smk_xattn_fold_weights(const float* in_proj_w, const float* in_proj_b, const float* out_w, const float* out_b, void* wg, float* g,
                           float* mcat, float* bo2, int D, int heads, void* stream);

/* Online-softmax attention (64 queries x 64-key blocks per step, mma.sync) for any sequence length: 384x384 images (577 tokens),
 * ViT-S/8 (785 tokens) and, with q_lo / k_lo / v_lo non-NULL, the bf16x3 split mode (3-term products, ~fp32 accuracy).
 * Image b: queries at rows b*q_rows .. +Lq, keys / values at rows b*kv_rows + kv_row0 .. +Lk; out_mode as smk_attention_small. */
int smk_attention_fa(const void* q, const void* q_lo, int64_t ldq, const void* k, const void* k_lo, int64_t ldk, const void* v,
                     const void* v_lo, int64_t ldv, int q_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_mode,
                     int B, int Lq, int Lk, int heads, float scale, void* stream);

/* Residual GEMM with the following LayerNorm fused into the epilogue (tcgen05, N must be 384):
 * X[M,N] (fp32, in place) += A[M,K] (bf16, lda) · W[N,K]^T (bf16) + bias;  Xn[M,N] (bf16) = LayerNorm(X; gamma, beta, eps).
 * Serves attn.proj + norm2 and mlp.fc2 + the next block's norm1 (vision_transformer.py:131,165-170). */
int smk_gemm_ln(const void* A, int64_t lda, const void* W, const float* bias, float* X, const float* gamma, const float* beta,
                void* Xn, int M, int N, int K, float eps, void* stream);

/* 3-term bf16 split along K (bf16x3): x fp32 [rows,K] → out bf16 [rows,3K]; activations [hi|hi|lo], weights [hi|lo|hi];
 * gemm_bf16(split_act(A), split_weight(W)) with K' = 3K ≈ fp32 GEMM */
int smk_split3(const float* x, int64_t rows, int K, void* out, int is_weight, void* stream);
/* fp32 → bf16 (round to nearest even) */
int smk_cast_bf16(const float* in, void* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SELFMASK_B200_H_ */
