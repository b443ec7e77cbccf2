// Internal (non-ABI) launcher declarations shared between the .cu files.
#pragma once
#include "smk_common.cuh"

namespace smk {

// Traversal direction of the next row-streaming launch (LayerNorm rows, GEMM tiles, attention items): 0 = ascending,
// 1 = descending.  The encoder flips it after every launch, so each kernel starts on the rows its producer wrote LAST — the
// ones still in the 126 MB L2 — instead of the rows that were evicted first (an LRU stream larger than the cache never hits
// when producer and consumer walk in the same direction).  Host-side, read at launch time.
extern thread_local int g_traverse_rev;
extern thread_local int g_traverse_alt;   // 1: every honouring launch flips g_traverse_rev after reading it
inline int traverse_dir() {
  const int r = g_traverse_rev;
  if (g_traverse_alt) g_traverse_rev ^= 1;
  return r;
}


int layernorm_f32(const float* x, const float* res, const float* gamma, const float* beta, float* y, float* sum_out, int64_t rows,
                  int D, float eps, cudaStream_t s);
int layernorm_split3(const float* x, const float* gamma, const float* beta, __nv_bfloat16* out3, int64_t rows, int D, float eps, cudaStream_t s);
int layernorm_bf16(const float* x, const float* res, const float* gamma, const float* beta, __nv_bfloat16* y, float* y32,
                   float* sum_out, int64_t rows, int D, float eps, cudaStream_t s, __nv_bfloat16* y_lo = nullptr, int64_t ldy = 0);
int layernorm_f16(const float* x, const float* gamma, const float* beta, __half* y, __half* y_lo, int64_t ldy, float* y32,
                  __nv_bfloat16* alt_hi, __nv_bfloat16* alt_lo, int64_t rows, int D, float eps, cudaStream_t s, int64_t ld_alt = 0, int q8 = 0);
// y2s_period / y2s_stride: row r of the split final-norm output goes to row (r / period)·stride + r % period of y2s (image-major
// [B][L][nq] layout of the mask / objectness head operand: period = nq, stride = L·nq, y2s pre-offset by layer·nq rows); 0 = row r
int dec_layernorm(float* x, const float* res, const float* gamma, const float* beta, float eps, const float* pos, int period,
                  __nv_bfloat16* a3a, __nv_bfloat16* a3b, const float* gamma2, const float* beta2, float* y2, __nv_bfloat16* y2s,
                  int64_t rows, int D, cudaStream_t s, int y2s_period = 0, int y2s_stride = 0, __half* xh = nullptr);
// restructured decoder cross-attention on tcgen05 (smk_xattn_tc.cu)
bool xattn_supported(int nq, int heads, int D, int hw);
int xattn_fold_weights(const float* in_proj_w, const float* in_proj_b, const float* out_w, const float* out_b, __half* wg, float* g, float* mcat,
                       float* bo2, int D, int heads, cudaStream_t s);
int xattn_tc(const __half* qp, const __half* tok, int t_rows_per_img, int t_row0, __half* out, int n_img, int nq, int heads, int D, int hw,
             cudaStream_t s);
int gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc, int M, int N, int K,
             int epi, cudaStream_t s);
int gemm_bf16_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias, void* C, int64_t ldc,
                 int M, int N, int K, int epi, int out_f32, int tok_hw, const float* tok_pos, cudaStream_t s, int credit_k = 0);
// split-operand tcgen05 GEMM (smk_gemm_tc.cu): C = Σ_t A[:, a_off[t] : +K] · W[:, w_off[t] : +K]^T, 16-bit operands (fp16 when f16),
// out_f32: 0 16-bit, 1 fp32, 2 [hi | hi | lo] split, 3 [hi | lo] split
struct GemmTerms { int n; int a_off[3]; int w_off[3]; int q8 = 0; };
inline GemmTerms terms_plain() { return GemmTerms{1, {0, 0, 0}, {0, 0, 0}}; }
inline GemmTerms terms_wsplit(int K) { return GemmTerms{2, {0, 0, 0}, {0, K, 0}}; }        // A_hi·(W_hi + W_lo)
inline GemmTerms terms_full(int K) { return GemmTerms{3, {0, 0, K}, {0, K, 0}}; }          // hi·hi + hi·lo + lo·hi
// fp16 hi·hi + the two correction products on e4m3 tiles stored in place of the fp16 `lo` half (TcGemmParams::q8, split_q8 below)
inline GemmTerms terms_q8(int K) { return GemmTerms{2, {K, 0, 0}, {K, 0, 0}, 1}; }
// the same three products on the round-1 layouts: activations [hi | hi | lo] (3K columns), weights [hi | lo | hi]
inline GemmTerms terms_legacy3(int K) { return GemmTerms{3, {0, 0, 2 * K}, {0, K, 0}}; }
int gemm_tc_batched(const void* A, int64_t lda, int64_t a_total_rows, int batch_a_rows, int a_row0, int rows_a, const void* W, int64_t ldw,
                    int64_t w_total_rows, int batch_w_rows, int w_row0, int rows_w, float* C, int n_batch, int K, int f16, const GemmTerms& terms,
                    cudaStream_t s);
int gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K, int epi,
            int out_f32, int tok_hw, const float* tok_pos, int f16, const GemmTerms& terms, int credit_k, cudaStream_t s);
template <typename T, typename TK>
int attention(const T* q, const TK* k, const TK* v, T* o, int batch, int heads, int dh, int Lq, int Lk, int64_t q_bs, int64_t ldq,
              int64_t k_bs, int64_t ldk, int64_t v_bs, int64_t ldv, int64_t o_bs, int64_t ldo, float scale, cudaStream_t s);
// residual GEMM + the following LayerNorm in one kernel (smk_gemm_ln.cu): X += A·W^T + bias (fp32), Xn = LN(X) (bf16); N == 384
int gemm_ln_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, const float* bias, float* X, const float* gamma,
               const float* beta, __nv_bfloat16* Xn, int M, int N, int K, float eps, cudaStream_t s);
int attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int heads, float scale, cudaStream_t s);
int attention_tc_f16(const __half* qkv, __half* out, int64_t ldo, int out_mode, int B, int N, int heads, float scale, cudaStream_t s);
// multi-key-tile tcgen05 attention for Lk >= 176 (577 / 785-token encoders; smk_attn_tc_multi.cu); 16-bit operands (fp16 when f16),
// out_mode 0 16-bit, 1 fp32, 2 [hi | hi | lo], 3 [hi | lo]
int attention_tc_multi(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t q_total_rows,
                       int64_t kv_total_rows, int q_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq, int Lk,
                       int heads, float scale, int f16, cudaStream_t s, int out_bf16 = 0);
// general form: q [B*Lq, ldq] / k / v [.., ld] bf16 matrices (head h at columns [h*64, h*64+64) from the given base pointer);
// image b's queries start at row b*Lq, its keys/values at row b*kv_rows + kv_row0.  Lq <= 128, Lk <= 256.
// out: [B*Lq, ldo], fp32 when out_f32 else bf16.
int attention_tc_general(const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v, int64_t ldv,
                         int64_t kv_total_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_f32, int B, int Lq, int Lk,
                         int heads, float scale, cudaStream_t s);
// few-query attention (decoder): Lq <= 32, Lk <= 256; out_mode 0 bf16, 1 fp32, 2 bf16x3 split (smk_attn_small.cu)
int attention_small(const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v, int64_t ldv,
                    int kv_rows, int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq, int Lk, int heads, float scale,
                    cudaStream_t s, int f16 = 0, int q_f32 = 0);
// online-softmax attention for long sequences / the bf16x3 split mode (smk_attn_fa.cu); *_lo == nullptr → plain bf16 operands
int attention_fa(const __nv_bfloat16* q, const __nv_bfloat16* q_lo, int64_t ldq, const __nv_bfloat16* k, const __nv_bfloat16* k_lo, int64_t ldk,
                 const __nv_bfloat16* v, const __nv_bfloat16* v_lo, int64_t ldv, int q_rows, int kv_rows, int kv_row0, void* out, int64_t ldo,
                 int out_mode, int B, int Lq, int Lk, int heads, float scale, cudaStream_t s, int f16 = 0);
int split3_act(const float* x, int64_t ldx, const float* pos, int period, __nv_bfloat16* out_a, __nv_bfloat16* out_b, int64_t rows,
               int K, cudaStream_t s);
int split3_weight(const float* w, __nv_bfloat16* out, int64_t rows, int K, cudaStream_t s);
template <typename TIn, typename T>
int im2col(const TIn* x, T* cols, int B, int H, int W, int P, int hp, int wp, const float* mean_std /* host, 6 floats or null */,
           cudaStream_t s);
template <typename TIn>
int im2col_split_f16(const TIn* x, __half* cols, int B, int H, int W, int P, int hp, int wp, const float* mean_std, cudaStream_t s, int q8 = 0);
int split2_f16(const float* w, int64_t ldw, __half* out, int64_t rows, int K, cudaStream_t s);
// row → [hi fp16 (K) | e4m3 correction operands (2K bytes)] for the fp8-corrected GEMMs (terms_q8); is_weight picks the weight-side scales
int split_q8(const float* x, int64_t ldx, __half* out, int64_t rows, int K, int is_weight, cudaStream_t s);
int dec_self_attention(const float* qk, int64_t ldqk, const float* v, int64_t ldv, const float* v_sub, __nv_bfloat16* out3, int B, int nq, int heads,
                       float scale, cudaStream_t s);
int assemble_tokens(const float* patch_out, const float* cls, const float* pos, float* tokens, int B, int hw, int D, bool cls_only,
                    cudaStream_t s);
int add_rows(const float* a, const float* pos, float* out, int64_t rows, int D, int period, cudaStream_t s);
int cast_bf16(const float* in, __nv_bfloat16* out, int64_t n, cudaStream_t s);
int cast_f16(const float* in, __half* out, int64_t n, cudaStream_t s);
int tile_rows2(void* d0, const void* s0, int row_bytes0, void* d1, const void* s1, int row_bytes1, int64_t rows, int period, cudaStream_t s);
// tensor-core modes, scale factor 4 (smk_mask_mma.cu): logits as a batched tcgen05 GEMM with the 3-term split, then upsample + sigmoid
int mask_head_tc(const __nv_bfloat16* q3, int q_batch_rows, const __nv_bfloat16* tok16, float* logits_lowres, float* mask_pred,
                 float* logits_out, int B, int L, int layer0, int nq, int D, int hp, int wp, cudaStream_t s);
int mask_head(const float* queries, const float* tokens, float* mask_pred, float* logits_out, int B, int L, int layer0, int nq, int D,
              int hp, int wp, int sf, cudaStream_t s, bool precise = false);
int rowdot_sigmoid(const float* h, const float* w, const float* bias, float* out, int64_t rows, int D, cudaStream_t s);
int permute_lb(const float* in, float* out, int L, int B, int n, cudaStream_t s);
int query_mean(const float* qlast, float* out, int B, int nq, int D, cudaStream_t s);
int pos_bicubic(const float* pos, float* out, int g, int hp, int wp, int D, cudaStream_t s);

}  // namespace smk
