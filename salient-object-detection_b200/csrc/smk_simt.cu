// CUDA-core kernels: LayerNorm, fp32 GEMM (validation mode + the high-precision decoder tail), generic
// softmax attention, patch im2col, token assembly, residual+LayerNorm, mask head (contraction at patch
// resolution → bilinear → sigmoid), objectness tail, casts.  Memory-bound ones are vectorised, one warp
// per row with shuffle reductions.
#include <type_traits>

#include "smk_common.cuh"
#include "smk_kernels.h"

namespace smk {

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (vision_transformer.py:165,169,299 eps 1e-6; transformer_decoder.py eps 1e-5).
// One warp per row; the row lives in registers (D <= 1024, D % 128 == 0); optional residual input and
// optional second (fp32) output so that bf16-mode callers get both copies in one pass.
// ------------------------------------------------------------------------------------------------
constexpr int LN_ROWS = 1;   // rows per warp.  Measured in the fp16s step (us per launch, event-timed): 2 rows / 3 CTAs per SM 32.6, 2 / 4 31.3,
                             // 4 / 3 35.0, 1 / 8 (32 registers, spills) 33.2, **1 / 6 (40 registers) 29.8**: latency-bound, not bandwidth-bound

template <typename TOut, int kChunks>
__global__ void __launch_bounds__(256, kChunks <= 3 ? 6 : 1)
layernorm_kernel(const float* x /* may alias y: a row is fully read before it is written */, const float* __restrict__ res,
                 const float* __restrict__ gamma, const float* __restrict__ beta, TOut* y, float* __restrict__ y32, float* sum_out,
                 TOut* __restrict__ y_lo /* bf16 only: rounding residue of y, so that y + y_lo ≈ the fp32 row */,
                 TOut* __restrict__ y_dup /* bf16 only: second copy of y; (y, y_dup, y_lo) at columns 0, D, 2D of rows of ldy = 3D
                                             elements form the bf16x3 split [hi | hi | lo] a split GEMM consumes */,
                 __nv_bfloat16* __restrict__ alt_hi, __nv_bfloat16* __restrict__ alt_lo /* optional bf16 hi / lo copies (rows of D) next to
                                             a 16-bit y of another type: the final encoder norm feeds fp16 GEMMs and the bf16 mask head */,
                 int64_t ldy, int64_t ld_alt, int64_t rows, int D, float eps, int rev,
                 int q8 /* fp16 only: y_lo receives the e4m3 correction operands (smk_common.cuh split_q8x4) instead of the fp16 residue */) {
  pdl_wait();
  pdl_trigger();
  const int64_t blk = rev ? (int64_t)gridDim.x - 1 - blockIdx.x : (int64_t)blockIdx.x;     // descending: start on the rows written last
  const int64_t row0 = (blk * 8 + (threadIdx.x >> 5)) * LN_ROWS;
  if (row0 >= rows) return;
  const int lane = threadIdx.x & 31;
  float4 v[LN_ROWS][kChunks];
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    const int64_t row = row0 + r;
    if (row < rows) {
      const float4* xr = reinterpret_cast<const float4*>(x + row * D);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) v[r][c] = xr[lane + 32 * c];
      if (res) {
        const float4* rr = reinterpret_cast<const float4*>(res + row * D);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const float4 q = rr[lane + 32 * c];
          v[r][c].x += q.x; v[r][c].y += q.y; v[r][c].z += q.z; v[r][c].w += q.w;
        }
      }
    }
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    const int64_t row = row0 + r;
    if (row >= rows) break;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) s += (v[r][c].x + v[r][c].y) + (v[r][c].z + v[r][c].w);
    if (sum_out) {   // the pre-norm residual stream (x + res), needed by post-norm callers
      float4* so = reinterpret_cast<float4*>(sum_out + row * D);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) so[lane + 32 * c] = v[r][c];
    }
    const float mean = warp_sum(s) / (float)D;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      float a = v[r][c].x - mean, b = v[r][c].y - mean, cc = v[r][c].z - mean, d = v[r][c].w - mean;
      ss += (a * a + b * b) + (cc * cc + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(ss) / (float)D + eps);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int i = lane + 32 * c;
      float4 g = g4[i], b = b4[i], o;
      o.x = (v[r][c].x - mean) * rstd * g.x + b.x;
      o.y = (v[r][c].y - mean) * rstd * g.y + b.y;
      o.z = (v[r][c].z - mean) * rstd * g.z + b.z;
      o.w = (v[r][c].w - mean) * rstd * g.w + b.w;
      if (y32) reinterpret_cast<float4*>(y32 + row * D)[i] = o;
      if (y) {
        if constexpr (sizeof(TOut) == 4) {
          reinterpret_cast<float4*>(y + row * ldy)[i] = o;
        } else {
          uint2 pk, pr;
          bool done = false;
          if constexpr (std::is_same<TOut, __half>::value) {
            if (q8 && y_lo) {
              uint32_t f8, s8;
              split_q8x4<false>(o.x, o.y, o.z, o.w, pk, f8, s8);
              uint8_t* q = reinterpret_cast<uint8_t*>(y_lo + row * ldy) + q8_byte_off(4 * i);
              *reinterpret_cast<uint32_t*>(q) = f8;
              *reinterpret_cast<uint32_t*>(q + 32) = s8;
              reinterpret_cast<uint2*>(y + row * ldy)[i] = pk;
              done = true;
            }
          }
          if (!done) {
            split16x2<TOut>(o.x, o.y, pk.x, pr.x);
            split16x2<TOut>(o.z, o.w, pk.y, pr.y);
            reinterpret_cast<uint2*>(y + row * ldy)[i] = pk;
            if (y_dup) reinterpret_cast<uint2*>(y_dup + row * ldy)[i] = pk;
            if (y_lo) reinterpret_cast<uint2*>(y_lo + row * ldy)[i] = pr;
          }
        }
      }
      if constexpr (sizeof(TOut) == 2) {
        if (alt_hi) {
          uint2 pk, pr;
          split16x2<__nv_bfloat16>(o.x, o.y, pk.x, pr.x);
          split16x2<__nv_bfloat16>(o.z, o.w, pk.y, pr.y);
          reinterpret_cast<uint2*>(alt_hi + row * ld_alt)[i] = pk;
          if (alt_lo) reinterpret_cast<uint2*>(alt_lo + row * ld_alt)[i] = pr;
        }
      }
    }
  }
}

template <typename TOut>
static int launch_layernorm(const float* x, const float* res, const float* gamma, const float* beta, TOut* y, float* y32,
                            float* sum_out, TOut* y_lo, int64_t rows, int D, float eps, cudaStream_t s, TOut* y_dup = nullptr, int64_t ldy = 0,
                            __nv_bfloat16* alt_hi = nullptr, __nv_bfloat16* alt_lo = nullptr, int64_t ld_alt = 0, int q8 = 0) {
  if (ldy == 0) ldy = D;
  if (ld_alt == 0) ld_alt = D;
  SMK_REQUIRE(D % 128 == 0 && D <= 1024, "layernorm: D=%d must be a multiple of 128 and <= 1024", D);
  if (rows == 0) return SMK_OK;
  const unsigned grid = (unsigned)((rows + 8 * LN_ROWS - 1) / (8 * LN_ROWS));
  const int rev = traverse_dir();
  ProfScope prof(PROF_LAYERNORM, (double)rows * D * (4.0 + (res ? 4.0 : 0.0) + (y ? sizeof(TOut) : 0) + (y32 ? 4.0 : 0.0) + (y_lo ? sizeof(TOut) : 0) +
                                                     (y_dup ? sizeof(TOut) : 0) + (alt_hi ? 2.0 : 0.0) + (alt_lo ? 2.0 : 0.0)), s);
  switch (D / 128) {
#define SMK_LN_CASE(c) \
  case c: SMK_CHECK_CUDA(launch_pdl(layernorm_kernel<TOut, c>, dim3(grid), dim3(256), 0, s, x, res, gamma, beta, y, y32, sum_out, y_lo, y_dup, alt_hi, alt_lo, ldy, ld_alt, rows, D, eps, rev, q8)); break;
    SMK_LN_CASE(1) SMK_LN_CASE(2) SMK_LN_CASE(3) SMK_LN_CASE(4) SMK_LN_CASE(5) SMK_LN_CASE(6) SMK_LN_CASE(7) SMK_LN_CASE(8)
#undef SMK_LN_CASE
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

int layernorm_f32(const float* x, const float* res, const float* gamma, const float* beta, float* y, float* sum_out,
                  int64_t rows, int D, float eps, cudaStream_t s) {
  return launch_layernorm<float>(x, res, gamma, beta, y, nullptr, sum_out, nullptr, rows, D, eps, s);
}
int layernorm_bf16(const float* x, const float* res, const float* gamma, const float* beta, __nv_bfloat16* y, float* y32,
                   float* sum_out, int64_t rows, int D, float eps, cudaStream_t s, __nv_bfloat16* y_lo, int64_t ldy) {
  return launch_layernorm<__nv_bfloat16>(x, res, gamma, beta, y, y32, sum_out, y_lo, rows, D, eps, s, nullptr, ldy);
}

// fp16 output for the fp16s mode: y = fp16(LN(x)) in rows of ldy elements; y_lo (optional) = fp16 rounding residue, normally at
// y + D so that a row is the [hi | lo] split operand of a 3-term GEMM; y32 / alt_hi / alt_lo: optional fp32 and bf16 hi / lo copies
int layernorm_f16(const float* x, const float* gamma, const float* beta, __half* y, __half* y_lo, int64_t ldy, float* y32,
                  __nv_bfloat16* alt_hi, __nv_bfloat16* alt_lo, int64_t rows, int D, float eps, cudaStream_t s, int64_t ld_alt, int q8) {
  return launch_layernorm<__half>(x, nullptr, gamma, beta, y, y32, nullptr, y_lo, rows, D, eps, s, nullptr, ldy, alt_hi, alt_lo, ld_alt, q8);
}

// LayerNorm whose output is the bf16x3 split [hi | hi | lo] (rows of 3D bf16) of the normalised row: the A operand of a split
// GEMM, without the fp32 round trip through a separate split kernel (bf16x3 mode encoder).
int layernorm_split3(const float* x, const float* gamma, const float* beta, __nv_bfloat16* out3, int64_t rows, int D, float eps, cudaStream_t s) {
  return launch_layernorm<__nv_bfloat16>(x, nullptr, gamma, beta, out3, nullptr, nullptr, out3 + 2 * D, rows, D, eps, s, out3 + D, 3 * (int64_t)D);
}

// ------------------------------------------------------------------------------------------------
// Decoder post-norm step (transformer_decoder.py:260-297), fused:  y = LN(x + res)  (written back over x) and, as asked,
//   a3a = bf16x3 split of y, a3b = bf16x3 split of (y + query_pos)  — the A operands of the next split GEMMs —
//   y2 = LN2(y) (the shared final decoder.norm, :138-145) in fp32 and as a bf16x3 split (objectness head input).
// One warp per row, the row stays in registers.  Split rows are [hi | hi | lo], 3·D columns.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_split4(__nv_bfloat16* row3, int D, int e, float4 v) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
  uint2 hi, lo;
  hi.x = *reinterpret_cast<uint32_t*>(&h0); hi.y = *reinterpret_cast<uint32_t*>(&h1);
  lo.x = *reinterpret_cast<uint32_t*>(&l0); lo.y = *reinterpret_cast<uint32_t*>(&l1);
  *reinterpret_cast<uint2*>(row3 + e) = hi;
  *reinterpret_cast<uint2*>(row3 + D + e) = hi;
  *reinterpret_cast<uint2*>(row3 + 2 * D + e) = lo;
}

// 128-thread CTAs (4 rows): 5120 rows = 1280 CTAs fit in ONE resident wave (9 CTAs per SM x 148); with 256-thread CTAs the 640
// CTAs needed 1.08 waves of the 592 slots, i.e. a second, almost empty wave doubled the latency of this latency-bound kernel
constexpr int DLN_THREADS = 128;
template <int kChunks>
__global__ void __launch_bounds__(DLN_THREADS)
dec_layernorm_kernel(float* x, const float* __restrict__ res, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     const float* __restrict__ pos, int period, __nv_bfloat16* __restrict__ a3a, __nv_bfloat16* __restrict__ a3b,
                     const float* __restrict__ gamma2, const float* __restrict__ beta2, float* __restrict__ y2,
                     __nv_bfloat16* __restrict__ y2s, int64_t rows, int D, int y2s_period, int y2s_stride,
                     __half* __restrict__ xh /* optional: fp16(y + pos) rows of D — the fp16s mode's cross-attention query GEMM operand */) {
  pdl_wait();
  pdl_trigger();
  const int64_t row = (int64_t)blockIdx.x * (DLN_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4* xr = reinterpret_cast<float4*>(x + row * D);
  const float4* rr = res ? reinterpret_cast<const float4*>(res + row * D) : nullptr;
  float4 v[kChunks];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    v[c] = xr[lane + 32 * c];
    if (rr) {
      const float4 r = rr[lane + 32 * c];
      v[c].x += r.x; v[c].y += r.y; v[c].z += r.z; v[c].w += r.w;
    }
    s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
  }
  auto normalise = [&](const float* g, const float* b_, float sum) {
    const float mean = warp_sum(sum) / (float)D;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const float a = v[c].x - mean, b = v[c].y - mean, cc = v[c].z - mean, d = v[c].w - mean;
      ss += (a * a + b * b) + (cc * cc + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(ss) / (float)D + eps);
    float s2 = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int i = lane + 32 * c;
      const float4 gg = reinterpret_cast<const float4*>(g)[i], bb = reinterpret_cast<const float4*>(b_)[i];
      v[c].x = (v[c].x - mean) * rstd * gg.x + bb.x;
      v[c].y = (v[c].y - mean) * rstd * gg.y + bb.y;
      v[c].z = (v[c].z - mean) * rstd * gg.z + bb.z;
      v[c].w = (v[c].w - mean) * rstd * gg.w + bb.w;
      s2 += (v[c].x + v[c].y) + (v[c].z + v[c].w);
    }
    return s2;
  };
  const float s_y = normalise(gamma, beta, s);
  const float* prow = pos ? pos + (int64_t)(row % period) * D : nullptr;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int i = lane + 32 * c;
    xr[i] = v[c];
    if (a3a) store_split4(a3a + row * 3 * D, D, 4 * i, v[c]);
    if (a3b) {
      const float4 pp = reinterpret_cast<const float4*>(prow)[i];
      store_split4(a3b + row * 3 * D, D, 4 * i, make_float4(v[c].x + pp.x, v[c].y + pp.y, v[c].z + pp.z, v[c].w + pp.w));
    }
    if (xh) {
      const float4 pp = reinterpret_cast<const float4*>(prow)[i];
      uint2 u;
      u.x = Pack16<__half>::pack(v[c].x + pp.x, v[c].y + pp.y);
      u.y = Pack16<__half>::pack(v[c].z + pp.z, v[c].w + pp.w);
      reinterpret_cast<uint2*>(xh + row * D)[i] = u;
    }
  }
  if (gamma2) {
    const int64_t y2s_row = y2s_period > 0 ? (row / y2s_period) * y2s_stride + row % y2s_period : row;
    normalise(gamma2, beta2, s_y);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int i = lane + 32 * c;
      if (y2) reinterpret_cast<float4*>(y2 + row * D)[i] = v[c];
      if (y2s) store_split4(y2s + y2s_row * 3 * D, D, 4 * i, v[c]);
    }
  }
}

int dec_layernorm(float* x, const float* res, const float* gamma, const float* beta, float eps, const float* pos, int period,
                  __nv_bfloat16* a3a, __nv_bfloat16* a3b, const float* gamma2, const float* beta2, float* y2, __nv_bfloat16* y2s,
                  int64_t rows, int D, cudaStream_t s, int y2s_period, int y2s_stride, __half* xh) {
  SMK_REQUIRE(D % 128 == 0 && D <= 512, "dec_layernorm: D=%d must be a multiple of 128 and <= 512", D);
  SMK_REQUIRE((!a3b && !xh) || (pos && period > 0), "dec_layernorm: a3b / xh need the query positions");
  if (rows == 0) return SMK_OK;
  const unsigned grid = (unsigned)((rows + DLN_THREADS / 32 - 1) / (DLN_THREADS / 32));
  ProfScope prof(PROF_LAYERNORM, (double)rows * D * (8.0 + (res ? 4.0 : 0.0) + (a3a ? 6.0 : 0.0) + (a3b ? 6.0 : 0.0) + (y2 ? 4.0 : 0.0) +
                                                     (y2s ? 6.0 : 0.0)), s);
  switch (D / 128) {
#define SMK_DLN_CASE(c) \
  case c: SMK_CHECK_CUDA(launch_pdl(dec_layernorm_kernel<c>, dim3(grid), dim3(DLN_THREADS), 0, s, x, res, gamma, beta, eps, pos, period, a3a, a3b, gamma2, beta2, y2, y2s, rows, D, y2s_period, y2s_stride, xh)); break;
    SMK_DLN_CASE(1) SMK_DLN_CASE(2) SMK_DLN_CASE(3) SMK_DLN_CASE(4)
#undef SMK_DLN_CASE
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 GEMM  C[M,N] = A[M,K] · W[N,K]^T + bias  (both operands K-contiguous = nn.Linear layout)
// 64x64x16 tiles, 256 threads, 4x4 micro-tile.
// ------------------------------------------------------------------------------------------------
constexpr int GBM = 64, GBN = 64, GBK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N, int K, int epi) {
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;     // loader: row within the tile, k offset
  const int ty = tid >> 4, tx = tid & 15;          // compute: 16x16 threads, 4x4 outputs each
  const int am = m0 + lr, wn = n0 + lr;
  const float* ap = A + (int64_t)min(am, M - 1) * lda + lk;
  const float* wp = W + (int64_t)min(wn, N - 1) * ldw + lk;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GBK) {
    float4 a = *reinterpret_cast<const float4*>(ap + k0);
    float4 w = *reinterpret_cast<const float4*>(wp + k0);
    if (am >= M) a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wn >= N) w = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
    Bs[lk + 0][lr] = w.x; Bs[lk + 1][lr] = w.y; Bs[lk + 2][lr] = w.z; Bs[lk + 3][lr] = w.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (epi & SMK_EPI_GELU) v = gelu_erf(v);
      if (epi & SMK_EPI_RELU) v = fmaxf(v, 0.f);
      float* c = C + (int64_t)m * ldc + n;
      if (epi & SMK_EPI_RESIDUAL) v += *c;
      *c = v;
    }
  }
}

int gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc, int M, int N,
             int K, int epi, cudaStream_t s) {
  SMK_REQUIRE(K % GBK == 0 && lda % 4 == 0 && ldw % 4 == 0, "gemm_f32: K=%d lda=%lld ldw=%lld must be multiples of 16/4/4", K,
              (long long)lda, (long long)ldw);
  SMK_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "gemm_f32: operands must be 16-byte aligned");
  if (M == 0 || N == 0) return SMK_OK;
  dim3 grid((N + GBN - 1) / GBN, (M + GBM - 1) / GBM);
  {
    ProfScope prof(PROF_GEMM_F32, 2.0 * M * N * K, s);
    gemm_f32_kernel<<<grid, 256, 0, s>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, epi);
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// ------------------------------------------------------------------------------------------------
// Generic multi-head softmax attention, fp32 math, online softmax over 64-key tiles.
// vision_transformer.py:122-130 (scale after QK^T) and nn.MultiheadAttention (scale on q) — identical
// here because scale = 2^-3 is exact.  grid (ceil(Lq/32), heads, batch), 128 threads, dh = 64.
// ------------------------------------------------------------------------------------------------
constexpr int ATT_ROWS = 32, ATT_KT = 64, ATT_DH = 64, ATT_RPW = ATT_ROWS / 4;

template <typename T, typename TK>
__global__ void __launch_bounds__(128)
attention_kernel(const T* __restrict__ q, const TK* __restrict__ k, const TK* __restrict__ v, T* __restrict__ o, int Lq, int Lk,
                 int64_t q_bs, int64_t ldq, int64_t k_bs, int64_t ldk, int64_t v_bs, int64_t ldv, int64_t o_bs, int64_t ldo,
                 float scale) {
  __shared__ float qs[ATT_ROWS][ATT_DH];
  __shared__ float ks[ATT_KT][ATT_DH + 1];
  __shared__ float vs[ATT_KT][ATT_DH];
  const int b = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * ATT_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* qb = q + b * q_bs + h * ATT_DH;
  const TK* kb = k + b * k_bs + h * ATT_DH;
  const TK* vb = v + b * v_bs + h * ATT_DH;
  for (int i = threadIdx.x; i < ATT_ROWS * ATT_DH; i += 128) {
    const int r = i >> 6, d = i & 63;
    qs[r][d] = (r0 + r < Lq) ? to_float(qb[(int64_t)(r0 + r) * ldq + d]) * scale : 0.f;
  }
  float m[ATT_RPW], l[ATT_RPW], a0[ATT_RPW], a1[ATT_RPW];
#pragma unroll
  for (int i = 0; i < ATT_RPW; ++i) { m[i] = -INFINITY; l[i] = 0.f; a0[i] = 0.f; a1[i] = 0.f; }
  for (int kt = 0; kt < Lk; kt += ATT_KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < ATT_KT * ATT_DH; i += 128) {
      const int r = i >> 6, d = i & 63;
      const bool ok = kt + r < Lk;
      ks[r][d] = ok ? to_float(kb[(int64_t)(kt + r) * ldk + d]) : 0.f;
      vs[r][d] = ok ? to_float(vb[(int64_t)(kt + r) * ldv + d]) : 0.f;
    }
    __syncthreads();
    const bool ok0 = kt + lane < Lk, ok1 = kt + lane + 32 < Lk;
#pragma unroll
    for (int i = 0; i < ATT_RPW; ++i) {
      const float* qr = qs[warp * ATT_RPW + i];
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
      for (int d = 0; d < ATT_DH; ++d) {
        const float qv = qr[d];
        s0 = fmaf(qv, ks[lane][d], s0);
        s1 = fmaf(qv, ks[lane + 32][d], s1);
      }
      s0 = ok0 ? s0 : -INFINITY;
      s1 = ok1 ? s1 : -INFINITY;
      const float mn = fmaxf(m[i], warp_max(fmaxf(s0, s1)));   // finite: every tile holds >= 1 valid key
      const float corr = expf(m[i] - mn);
      const float p0 = expf(s0 - mn), p1 = expf(s1 - mn);
      l[i] = l[i] * corr + (p0 + p1);
      float c0 = a0[i] * corr, c1 = a1[i] * corr;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p0, j);
        c0 = fmaf(pj, vs[j][lane], c0);
        c1 = fmaf(pj, vs[j][lane + 32], c1);
      }
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p1, j);
        c0 = fmaf(pj, vs[j + 32][lane], c0);
        c1 = fmaf(pj, vs[j + 32][lane + 32], c1);
      }
      a0[i] = c0; a1[i] = c1; m[i] = mn;
    }
  }
  T* ob = o + b * o_bs + h * ATT_DH;
#pragma unroll
  for (int i = 0; i < ATT_RPW; ++i) {
    const int r = r0 + warp * ATT_RPW + i;
    if (r >= Lq) continue;
    const float inv = 1.0f / warp_sum(l[i]);
    ob[(int64_t)r * ldo + lane] = from_float<T>(a0[i] * inv);
    ob[(int64_t)r * ldo + lane + 32] = from_float<T>(a1[i] * inv);
  }
}

template <typename T, typename TK>
int attention(const T* q, const TK* k, const TK* v, T* o, int batch, int heads, int dh, int Lq, int Lk, int64_t q_bs, int64_t ldq,
              int64_t k_bs, int64_t ldk, int64_t v_bs, int64_t ldv, int64_t o_bs, int64_t ldo, float scale, cudaStream_t s) {
  SMK_REQUIRE(dh == ATT_DH, "attention: head dim %d != 64", dh);
  SMK_REQUIRE(Lq > 0 && Lk > 0 && batch <= 65535 && heads <= 65535, "attention: bad sizes");
  if (batch == 0) return SMK_OK;
  dim3 grid((Lq + ATT_ROWS - 1) / ATT_ROWS, heads, batch);
  {
    ProfScope prof(PROF_ATTENTION, 4.0 * Lq * Lk * dh * heads * batch, s);
    attention_kernel<T, TK><<<grid, 128, 0, s>>>(q, k, v, o, Lq, Lk, q_bs, ldq, k_bs, ldk, v_bs, ldv, o_bs, ldo, scale);
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}
#define SMK_ATT_INST(T, TK)                                                                                                      \
  template int attention<T, TK>(const T*, const TK*, const TK*, T*, int, int, int, int, int, int64_t, int64_t, int64_t, int64_t, \
                                int64_t, int64_t, int64_t, int64_t, float, cudaStream_t);
SMK_ATT_INST(float, float)
SMK_ATT_INST(__nv_bfloat16, __nv_bfloat16)
SMK_ATT_INST(float, __nv_bfloat16)
#undef SMK_ATT_INST

// ------------------------------------------------------------------------------------------------
// Patch im2col (vision_transformer.py:184-188 conv k=s=P as a GEMM; :260-267 zero pad right/bottom):
// cols[b*hw + py*wp + px][c*P*P + ky*P + kx] = x[b,c,py*P+ky,px*P+kx]
// One CTA per (patch row, image): the 3 x P x W pixel strip goes through shared memory so that both the image reads
// (whole rows) and the column writes (16-byte vectors of one patch row) are coalesced.
// TIn = uint8_t: raw pixels, normalised here exactly like the reference's host-side loader
// (datasets/base_dataset.py:250, torchvision to_tensor + normalize): ((float)u / 255 - mean[c]) / std[c] with IEEE
// fp32 division and no FMA contraction → bit-identical to the float path fed with host-normalised images, at a quarter
// of the host→device bytes.
// ------------------------------------------------------------------------------------------------
struct NormConst { float mean[3], std[3]; };

// `lut` (uint8 input only): the 3 x 256 possible normalised values, computed per CTA with the exact expression above
template <typename TIn> __device__ __forceinline__ float load_pixel(const TIn* p, int c, const float* lut);
template <> __device__ __forceinline__ float load_pixel<float>(const float* p, int, const float*) { return __ldg(p); }
template <> __device__ __forceinline__ float load_pixel<uint8_t>(const uint8_t* p, int c, const float* lut) { return lut[c * 256 + __ldg(p)]; }

template <typename TIn> __device__ __forceinline__ void load_pixels4(const TIn* p, int c, const float* lut, float (&v)[4]);
template <> __device__ __forceinline__ void load_pixels4<float>(const float* p, int, const float*, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load_pixels4<uint8_t>(const uint8_t* p, int c, const float* lut, float (&v)[4]) {
  const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(p));
  const float* l = lut + c * 256;
  v[0] = l[t.x]; v[1] = l[t.y]; v[2] = l[t.z]; v[3] = l[t.w];
}
__device__ __forceinline__ void store4(float* dst, const float (&v)[4]) { *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void store4(__nv_bfloat16* dst, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(dst) = u;
}

template <typename TIn, typename T>
__global__ void __launch_bounds__(256)
im2col_kernel(const TIn* __restrict__ x, T* __restrict__ cols, int H, int W, int P, int hp, int wp, NormConst nc, bool vec_ok) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t im2col_smem[];
  __shared__ float lut[sizeof(TIn) == 1 ? 3 * 256 : 1];
  if constexpr (sizeof(TIn) == 1) {      // every value a uint8 pixel can normalise to (two IEEE divisions each, done 768 times per CTA)
    for (int i = threadIdx.x; i < 3 * 256; i += 256) {
      const int c = i >> 8;
      lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), nc.mean[c]), nc.std[c]);
    }
    __syncthreads();
  }
  T* strip = reinterpret_cast<T*>(im2col_smem);         // [3*P][ws], ws = wp*P + pad
  const int py = blockIdx.x, b = blockIdx.y;
  const int Wp = wp * P, ws = Wp + 16 / (int)sizeof(T);
  const TIn* src = x + (int64_t)b * 3 * H * W;
  // phase 1: image rows → strip, 4 pixels per thread (zero beyond the image: make_input_divisible pads right/bottom)
  const int Wq = Wp >> 2;
  for (int i = threadIdx.x; i < 3 * P * Wq; i += 256) {
    const int r = i / Wq, xx = (i - r * Wq) << 2, c = r / P, ky = r - c * P, yy = py * P + ky;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (yy < H) {
      const TIn* row = src + ((int64_t)c * H + yy) * W;
      if (vec_ok && xx + 3 < W) load_pixels4<TIn>(row + xx, c, lut, v);
      else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (xx + e < W) v[e] = load_pixel<TIn>(row + xx + e, c, lut);
      }
    }
    store4(strip + r * ws + xx, v);
  }
  __syncthreads();
  // phase 2: strip → cols, 16-byte vectors (kV elements of one patch row); K = 3*P*P, P % kV == 0
  constexpr int kV = 16 / (int)sizeof(T);
  const int K = 3 * P * P, vec_per_row = P / kV, vec_per_patch = K / kV;
  T* dst = cols + ((int64_t)b * hp + py) * wp * K;
  for (int i = threadIdx.x; i < wp * vec_per_patch; i += 256) {
    const int px = i / vec_per_patch, j = i - px * vec_per_patch;     // j = (c*P + ky) * vec_per_row + kxv
    const int r = j / vec_per_row, kxv = j - r * vec_per_row;
    const uint4 v = *reinterpret_cast<const uint4*>(strip + r * ws + px * P + kxv * kV);
    *reinterpret_cast<uint4*>(dst + (int64_t)px * K + j * kV) = v;
  }
}
// fp16s mode: cols[row] = [hi (K) | lo (K)] fp16 split of the (normalised) pixels — the A operand of the 3-term patch-embed GEMM
// (the patch embed is the most rounding-sensitive contraction of the model: scripts/precision_emulation.py).  Same two phases with
// an fp32 strip; phase 2 splits 4 pixels per thread into 8 B of hi and 8 B of lo.
template <typename TIn>
__global__ void __launch_bounds__(256)
im2col_split_kernel(const TIn* __restrict__ x, __half* __restrict__ cols, int H, int W, int P, int hp, int wp, NormConst nc, bool vec_ok, int q8) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t im2col_smem[];
  __shared__ float lut[sizeof(TIn) == 1 ? 3 * 256 : 1];
  if constexpr (sizeof(TIn) == 1) {
    for (int i = threadIdx.x; i < 3 * 256; i += 256) {
      const int c = i >> 8;
      lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), nc.mean[c]), nc.std[c]);
    }
    __syncthreads();
  }
  float* strip = reinterpret_cast<float*>(im2col_smem);         // [3*P][ws], ws = wp*P + 4
  const int py = blockIdx.x, b = blockIdx.y;
  const int Wp = wp * P, ws = Wp + 4;
  const TIn* src = x + (int64_t)b * 3 * H * W;
  const int Wq = Wp >> 2;
  // index arithmetic is incremental (one division per thread, not per element): the kernel is instruction-bound, not HBM-bound
  {
    int r = threadIdx.x / Wq, xq = threadIdx.x - r * Wq;
    const int rstep = 256 / Wq, xstep = 256 - rstep * Wq;
    while (r < 3 * P) {
      const int xx = xq << 2, c = r / P, ky = r - c * P, yy = py * P + ky;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (yy < H) {
        const TIn* row = src + ((int64_t)c * H + yy) * W;
        if (vec_ok && xx + 3 < W) load_pixels4<TIn>(row + xx, c, lut, v);
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (xx + e < W) v[e] = load_pixel<TIn>(row + xx + e, c, lut);
        }
      }
      store4(strip + r * ws + xx, v);
      xq += xstep; r += rstep;
      if (xq >= Wq) { xq -= Wq; ++r; }
    }
  }
  __syncthreads();
  const int K = 3 * P * P, vec_per_row = P / 4, vec_per_patch = K / 4;
  const bool vpow2 = (vec_per_row & (vec_per_row - 1)) == 0;
  const int vshift = __popc(vec_per_row - 1);
  __half* dst = cols + ((int64_t)b * hp + py) * wp * 2 * K;
  {
    int px = threadIdx.x / vec_per_patch, j = threadIdx.x - px * vec_per_patch;      // j = (c*P + ky) * vec_per_row + kxv
    const int pstep = 256 / vec_per_patch, jstep = 256 - pstep * vec_per_patch;
    while (px < wp) {
      const int r = vpow2 ? j >> vshift : j / vec_per_row, kxv = j - r * vec_per_row;
      const float4 v = *reinterpret_cast<const float4*>(strip + r * ws + px * P + kxv * 4);
      uint2 hi, lo;
      __half* o = dst + (int64_t)px * 2 * K;
      if (q8) {
        split_q8x4<false>(v.x, v.y, v.z, v.w, hi, lo.x, lo.y);
        uint8_t* q = reinterpret_cast<uint8_t*>(o + K) + q8_byte_off(j * 4);
        *reinterpret_cast<uint2*>(o + j * 4) = hi;
        *reinterpret_cast<uint32_t*>(q) = lo.x;
        *reinterpret_cast<uint32_t*>(q + 32) = lo.y;
      } else {
        split16x2<__half>(v.x, v.y, hi.x, lo.x);
        split16x2<__half>(v.z, v.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(o + j * 4) = hi;
        *reinterpret_cast<uint2*>(o + K + j * 4) = lo;
      }
      j += jstep; px += pstep;
      if (j >= vec_per_patch) { j -= vec_per_patch; ++px; }
    }
  }
}
template <typename TIn>
int im2col_split_f16(const TIn* x, __half* cols, int B, int H, int W, int P, int hp, int wp, const float* mean_std, cudaStream_t s, int q8) {
  if (B == 0) return SMK_OK;
  SMK_REQUIRE(B <= 65535 && P % 4 == 0, "im2col_split: bad batch / patch size");
  NormConst nc{{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  if (mean_std) for (int c = 0; c < 3; ++c) { nc.mean[c] = mean_std[c]; nc.std[c] = mean_std[3 + c]; }
  const int smem = 3 * P * (wp * P + 4) * 4;
  SMK_REQUIRE(smem <= 227 * 1024, "im2col_split: image too wide for the shared-memory strip (%d bytes)", smem);
  static int attr_max[kMaxDevices] = {};
  const int dev = current_device();
  if (smem > 48 * 1024 && smem > attr_max[dev]) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(im2col_split_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_max[dev] = smem;
  }
  {
    ProfScope prof(PROF_OTHER, (double)B * hp * wp * 3 * P * P * ((double)sizeof(TIn) + 4.0), s);
    const bool vec_ok = (W & 3) == 0 && ((uintptr_t)x % (4 * sizeof(TIn))) == 0;
    SMK_CHECK_CUDA(launch_pdl(im2col_split_kernel<TIn>, dim3(hp, B), dim3(256), (size_t)smem, s, x, cols, H, W, P, hp, wp, nc, vec_ok, q8));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}
template int im2col_split_f16<float>(const float*, __half*, int, int, int, int, int, int, const float*, cudaStream_t, int);
template int im2col_split_f16<uint8_t>(const uint8_t*, __half*, int, int, int, int, int, int, const float*, cudaStream_t, int);

template <typename TIn, typename T>
int im2col(const TIn* x, T* cols, int B, int H, int W, int P, int hp, int wp, const float* mean_std, cudaStream_t s) {
  if (B == 0) return SMK_OK;
  SMK_REQUIRE(B <= 65535, "im2col: batch too large");
  SMK_REQUIRE(P % (16 / (int)sizeof(T)) == 0, "im2col: patch size %d must be a multiple of %d", P, 16 / (int)sizeof(T));
  NormConst nc{{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  if (mean_std) for (int c = 0; c < 3; ++c) { nc.mean[c] = mean_std[c]; nc.std[c] = mean_std[3 + c]; }
  const int smem = 3 * P * (wp * P + 16 / (int)sizeof(T)) * (int)sizeof(T);
  SMK_REQUIRE(smem <= 227 * 1024, "im2col: image too wide for the shared-memory strip (%d bytes)", smem);
  static int attr_max[kMaxDevices] = {};
  const int dev = current_device();
  if (smem > 48 * 1024 && smem > attr_max[dev]) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(im2col_kernel<TIn, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_max[dev] = smem;
  }
  {
    ProfScope prof(PROF_OTHER, (double)B * hp * wp * 3 * P * P * ((double)sizeof(TIn) + sizeof(T)), s);
    const bool vec_ok = (W & 3) == 0 && ((uintptr_t)x % (4 * sizeof(TIn))) == 0;   // 4-pixel vector loads stay aligned in every row
    SMK_CHECK_CUDA(launch_pdl(im2col_kernel<TIn, T>, dim3(hp, B), dim3(256), (size_t)smem, s, x, cols, H, W, P, hp, wp, nc, vec_ok));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}
template int im2col<float, float>(const float*, float*, int, int, int, int, int, int, const float*, cudaStream_t);
template int im2col<float, __nv_bfloat16>(const float*, __nv_bfloat16*, int, int, int, int, int, int, const float*, cudaStream_t);
template int im2col<uint8_t, float>(const uint8_t*, float*, int, int, int, int, int, int, const float*, cudaStream_t);
template int im2col<uint8_t, __nv_bfloat16>(const uint8_t*, __nv_bfloat16*, int, int, int, int, int, int, const float*, cudaStream_t);

// tokens[b,0,:] = cls + pos[0];  tokens[b,1+p,:] = patch_out[b*hw+p,:] + pos[1+p]   (vision_transformer.py:276-280)
__global__ void __launch_bounds__(128)
assemble_tokens_kernel(const float* __restrict__ patch_out, const float* __restrict__ cls, const float* __restrict__ pos,
                       float* __restrict__ tokens, int hw, int D) {
  pdl_wait();
  pdl_trigger();
  const int t = blockIdx.x, b = blockIdx.y;   // t in [0, hw]; a grid of (1, B) writes the cls rows only
  const float* src = (t == 0) ? cls : patch_out + ((int64_t)b * hw + t - 1) * D;
  float* dst = tokens + ((int64_t)b * (hw + 1) + t) * D;
  for (int d = threadIdx.x; d < D; d += 128) dst[d] = src[d] + pos[(int64_t)t * D + d];
}
int assemble_tokens(const float* patch_out, const float* cls, const float* pos, float* tokens, int B, int hw, int D, bool cls_only,
                    cudaStream_t s) {
  if (B == 0) return SMK_OK;
  {
    ProfScope prof(PROF_OTHER, (double)B * (cls_only ? 1 : hw + 1) * D * 8.0, s);
    SMK_CHECK_CUDA(launch_pdl(assemble_tokens_kernel, dim3(cls_only ? 1 : hw + 1, B), dim3(128), 0, s, patch_out, cls, pos, tokens, hw, D));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// out[r, :] = a[r, :] + pos[r % period, :]   (query_pos broadcast over the batch, transformer_decoder.py:257-258)
__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ pos, float* __restrict__ out,
                                int64_t rows, int D, int period) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  const int64_t r = i / D;
  const int d = (int)(i % D);
  out[i] = (a ? a[i] : 0.f) + pos[(int64_t)(r % period) * D + d];
}
int add_rows(const float* a, const float* pos, float* out, int64_t rows, int D, int period, cudaStream_t s) {
  if (rows == 0) return SMK_OK;
  {
    ProfScope prof(PROF_OTHER, (double)rows * D * 8.0, s);
    add_rows_kernel<<<(unsigned)((rows * D + 255) / 256), 256, 0, s>>>(a, pos, out, rows, D, period);
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// dst[r] = src[r % period] for two row-periodic buffers at once (16-byte vectors): the decoder state after layer 0's
// self-attention block does not depend on the image (tgt = 0), so it is computed once and tiled over the batch.
__global__ void __launch_bounds__(256)
tile_rows2_kernel(uint4* __restrict__ d0, const uint4* __restrict__ s0, int v0, uint4* __restrict__ d1, const uint4* __restrict__ s1, int v1,
                  int64_t rows, int period) {
  pdl_wait();
  pdl_trigger();
  const int vt = v0 + v1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * vt) return;
  const int64_t r = i / vt;
  const int c = (int)(i - r * vt), pr = (int)(r % period);
  if (c < v0) d0[r * v0 + c] = __ldg(s0 + (int64_t)pr * v0 + c);
  else d1[r * v1 + (c - v0)] = __ldg(s1 + (int64_t)pr * v1 + (c - v0));
}
int tile_rows2(void* d0, const void* s0, int row_bytes0, void* d1, const void* s1, int row_bytes1, int64_t rows, int period, cudaStream_t s) {
  SMK_REQUIRE(row_bytes0 % 16 == 0 && row_bytes1 % 16 == 0 && period > 0, "tile_rows2: rows must be multiples of 16 bytes");
  if (rows == 0) return SMK_OK;
  const int v0 = row_bytes0 / 16, v1 = row_bytes1 / 16;
  {
    ProfScope prof(PROF_OTHER, (double)rows * (row_bytes0 + row_bytes1), s);
    SMK_CHECK_CUDA(launch_pdl(tile_rows2_kernel, dim3((unsigned)((rows * (v0 + v1) + 255) / 256)), dim3(256), 0, s, (uint4*)d0, (const uint4*)s0, v0,
                              (uint4*)d1, (const uint4*)s1, v1, rows, period));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + i) = pk;
  } else {
    for (int64_t j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}
int cast_bf16(const float* in, __nv_bfloat16* out, int64_t n, cudaStream_t s) {
  if (n == 0) return SMK_OK;
  SMK_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 8) == 0, "cast_bf16: misaligned");
  {
    ProfScope prof(PROF_OTHER, (double)n * 6.0, s);
    cast_bf16_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, s>>>(in, out, n);
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

__global__ void cast_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}
int cast_f16(const float* in, __half* out, int64_t n, cudaStream_t s) {
  if (n == 0) return SMK_OK;
  {
    ProfScope prof(PROF_OTHER, (double)n * 6.0, s);
    cast_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n);
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// ------------------------------------------------------------------------------------------------
// 3-term bf16 split ("bf16x3") laid out along K so that ONE tensor-core GEMM with K' = 3K computes
//   A·W^T ≈ A_hi·W_hi^T + A_hi·W_lo^T + A_lo·W_hi^T      (hi = bf16(x), lo = bf16(x − hi); error ~2^-16 relative)
// activations: row → [hi | hi | lo];  weights: row → [hi | lo | hi].  Used for the decoder tail, whose objectness
// ranking needs near-fp32 GEMMs but whose FLOPs are small.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_hi_lo(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// out_a[r] = split(x[r]);  out_b[r] = split(x[r] + pos[r % period])  (either output may be null; x may be null = 0)
__global__ void __launch_bounds__(256)
split3_act_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ pos, int period,
                  __nv_bfloat16* __restrict__ out_a, __nv_bfloat16* __restrict__ out_b, int64_t rows, int K) {
  pdl_wait();
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int64_t r = i / K;
  const int k = (int)(i % K);
  const float v = x ? x[r * ldx + k] : 0.f;
  __nv_bfloat16 hi, lo;
  if (out_a) {
    split_hi_lo(v, hi, lo);
    __nv_bfloat16* o = out_a + r * 3 * K;
    o[k] = hi; o[K + k] = hi; o[2 * K + k] = lo;
  }
  if (out_b) {
    split_hi_lo(v + pos[(int64_t)(r % period) * K + k], hi, lo);
    __nv_bfloat16* o = out_b + r * 3 * K;
    o[k] = hi; o[K + k] = hi; o[2 * K + k] = lo;
  }
}
int split3_act(const float* x, int64_t ldx, const float* pos, int period, __nv_bfloat16* out_a, __nv_bfloat16* out_b, int64_t rows,
               int K, cudaStream_t s) {
  if (rows == 0) return SMK_OK;
  {
    ProfScope prof(PROF_OTHER, (double)rows * K * (4.0 + (out_a ? 6.0 : 0.0) + (out_b ? 6.0 : 0.0)), s);
    SMK_CHECK_CUDA(launch_pdl(split3_act_kernel, dim3((unsigned)((rows * K + 255) / 256)), dim3(256), 0, s, x, ldx, pos, period > 0 ? period : 1, out_a, out_b, rows, K));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}
__global__ void __launch_bounds__(256)
split3_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int64_t rows, int K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int64_t r = i / K;
  const int k = (int)(i % K);
  __nv_bfloat16 hi, lo;
  split_hi_lo(w[i], hi, lo);
  __nv_bfloat16* o = out + r * 3 * K;
  o[k] = hi; o[K + k] = lo; o[2 * K + k] = hi;
}
// fp16s mode weights: row → [hi (K) | lo (K)] fp16
__global__ void __launch_bounds__(256)
split2_f16_kernel(const float* __restrict__ w, int64_t ldw, __half* __restrict__ out, int64_t rows, int K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int64_t r = i / K;
  const int k = (int)(i % K);
  const float v = w[r * ldw + k];
  const __half hi = __float2half_rn(v);
  __half* o = out + r * 2 * K;
  o[k] = hi;
  o[K + k] = __float2half_rn(v - __half2float(hi));
}
// fp16s mode with fp8 correction terms: row → [hi fp16 (K) | e4m3 correction operands (2K bytes)], smk_common.cuh split_q8x4
template <bool kWeight>
__global__ void __launch_bounds__(256)
split_q8_kernel(const float* __restrict__ x, int64_t ldx, __half* __restrict__ out, int64_t rows, int K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per 4 columns
  const int kq = K >> 2;
  if (i >= rows * kq) return;
  const int64_t r = i / kq;
  const int c = (int)(i - r * kq) << 2;
  const float* src = x + r * ldx + c;
  uint2 hi;
  uint32_t f8, s8;
  split_q8x4<kWeight>(src[0], src[1], src[2], src[3], hi, f8, s8);
  __half* o = out + r * 2 * K;
  *reinterpret_cast<uint2*>(o + c) = hi;
  uint8_t* q = reinterpret_cast<uint8_t*>(o + K) + q8_byte_off(c);
  *reinterpret_cast<uint32_t*>(q) = f8;
  *reinterpret_cast<uint32_t*>(q + 32) = s8;
}
int split_q8(const float* x, int64_t ldx, __half* out, int64_t rows, int K, int is_weight, cudaStream_t s) {
  SMK_REQUIRE(K % 32 == 0, "split_q8: K must be a multiple of 32");
  if (rows == 0) return SMK_OK;
  const unsigned grid = (unsigned)((rows * (K / 4) + 255) / 256);
  if (is_weight) split_q8_kernel<true><<<grid, 256, 0, s>>>(x, ldx, out, rows, K);
  else split_q8_kernel<false><<<grid, 256, 0, s>>>(x, ldx, out, rows, K);
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}
int split2_f16(const float* w, int64_t ldw, __half* out, int64_t rows, int K, cudaStream_t s) {
  if (rows == 0) return SMK_OK;
  split2_f16_kernel<<<(unsigned)((rows * K + 255) / 256), 256, 0, s>>>(w, ldw, out, rows, K);
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

int split3_weight(const float* w, __nv_bfloat16* out, int64_t rows, int K, cudaStream_t s) {
  if (rows == 0) return SMK_OK;
  split3_weight_kernel<<<(unsigned)((rows * K + 255) / 256), 256, 0, s>>>(w, out, rows, K);
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// ------------------------------------------------------------------------------------------------
// Mask head (maskformer.py:144-162, :223): logits at patch resolution = queries · memory^T, then the
// pixel decoder's bilinear xsf applied to the nq-channel logits (bilinear is linear and per-channel, so
// it commutes with the contraction — SURVEY.md K12), sigmoid, store mask_pred[b,l,q,:,:].
// grid (layer groups, B), 256 threads.  One CTA owns R = (layers per group) x nq query rows of one image:
//   stage 1  logits[R][hw] = Q[R][D] · T[hw][D]^T in fp32, register-tiled (8 rows x 4 tokens per thread and tile),
//            K streamed through shared memory in chunks of 32 (float4 reads, conflict-free row stride 36);
//   stage 2  each warp produces full output rows: x sf bilinear of the logits plane held in shared memory, sigmoid,
//            coalesced 128-byte stores (the only HBM traffic that matters: nq x (hp sf) x (wp sf) x 4 B per layer).
// ------------------------------------------------------------------------------------------------
constexpr int MH_KC = 32, MH_LD = MH_KC + 4, MH_MAXT_CAP = 3, MH_THREADS = 256;

// kDouble (fp32 validation mode): the 384-term dot products are accumulated in float64, so this side of the comparison carries no
// summation-order noise of its own — what remains against the reference is the reference's own fp32 rounding (~5e-5 on logits of
// magnitude up to 70; two fp32 sums in different orders differ by up to 1e-4, the whole north_star budget).
template <int MH_MAXT, bool kDouble>   // (8 rows x 4 tokens) register tiles per thread: 1 at 224x224 (245 tiles), up to 3 at 384x384
__global__ void __launch_bounds__(MH_THREADS, (MH_MAXT == 1 && !kDouble) ? 3 : 1)
mask_head_kernel(const float* __restrict__ queries /*[Lall,B,nq,D]*/, const float* __restrict__ tokens /*[B,N,D] final LN*/,
                 float* __restrict__ mask_pred /*[B,L,nq,hp*sf,wp*sf]*/, float* __restrict__ logits_out, int B, int L, int Lg, int nq,
                 int D, int hp, int wp, int sf, int layer0) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int hw = hp * wp, N = hw + 1;
  const int b = blockIdx.y, l0 = blockIdx.x * Lg;          // first output layer of this CTA
  const int nl = min(Lg, L - l0), R = nl * nq;             // query rows owned
  const int RO = (R + 7) / 8, R8 = RO * 8;                 // row octets
  const int NQ4 = (hw + 3) / 4;                            // token quads (token n = tq + NQ4 * i, i < 4)
  float* lg = sm;                                          // [R][hw]
  float* As = lg + (size_t)Lg * nq * hw;                   // [R8][MH_LD]
  float* Ts = As + (size_t)((Lg * nq + 7) / 8 * 8) * MH_LD;   // [NQ4 * 4][MH_LD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* mem = tokens + ((int64_t)b * N + 1) * D;    // skip the cls token (maskformer.py:104)
  const int n_tiles = RO * NQ4;

  using AccT = typename std::conditional<kDouble, double, float>::type;
  AccT acc[MH_MAXT][8][4];
#pragma unroll
  for (int t = 0; t < MH_MAXT; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][j][i] = (AccT)0;

  for (int k0 = 0; k0 < D; k0 += MH_KC) {
    __syncthreads();
    for (int i = tid; i < R8 * (MH_KC / 4); i += MH_THREADS) {
      const int r = i / (MH_KC / 4), c = i % (MH_KC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R) {
        const int l = layer0 + l0 + r / nq, q = r % nq;
        v = *reinterpret_cast<const float4*>(queries + (((int64_t)l * B + b) * nq + q) * D + k0 + 4 * c);
      }
      *reinterpret_cast<float4*>(As + r * MH_LD + 4 * c) = v;
    }
    for (int i = tid; i < NQ4 * 4 * (MH_KC / 4); i += MH_THREADS) {
      const int n = i / (MH_KC / 4), c = i % (MH_KC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < hw) v = *reinterpret_cast<const float4*>(mem + (int64_t)n * D + k0 + 4 * c);
      *reinterpret_cast<float4*>(Ts + n * MH_LD + 4 * c) = v;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < MH_MAXT; ++t) {
      const int tile = tid + t * MH_THREADS;
      if (tile < n_tiles) {
        const int ro = tile / NQ4, tq = tile % NQ4;
        const float* ap = As + ro * 8 * MH_LD;
        const float* tp = Ts + tq * MH_LD;
#pragma unroll 2
        for (int k = 0; k < MH_KC; k += 4) {
          float4 tv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) tv[i] = *reinterpret_cast<const float4*>(tp + (size_t)i * NQ4 * MH_LD + k);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 av = *reinterpret_cast<const float4*>(ap + j * MH_LD + k);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              AccT a = acc[t][j][i];
              if constexpr (kDouble) {
                a = fma((double)av.x, (double)tv[i].x, a);
                a = fma((double)av.y, (double)tv[i].y, a);
                a = fma((double)av.z, (double)tv[i].z, a);
                a = fma((double)av.w, (double)tv[i].w, a);
              } else {
                a = fmaf(av.x, tv[i].x, a);
                a = fmaf(av.y, tv[i].y, a);
                a = fmaf(av.z, tv[i].z, a);
                a = fmaf(av.w, tv[i].w, a);
              }
              acc[t][j][i] = a;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < MH_MAXT; ++t) {
    const int tile = tid + t * MH_THREADS;
    if (tile < n_tiles) {
      const int ro = tile / NQ4, tq = tile % NQ4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = ro * 8 + j;
        if (r < R) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = tq + NQ4 * i;
            if (n < hw) lg[r * hw + n] = (float)acc[t][j][i];
          }
        }
      }
    }
  }
  __syncthreads();
  const int Ho = hp * sf, Wo = wp * sf;
  const float rscale = 1.0f / (float)sf;
  float* out = mask_pred + (((int64_t)b * L + l0) * nq) * Ho * Wo;
  float* lout = logits_out ? logits_out + (((int64_t)b * L + l0) * nq) * Ho * Wo : nullptr;
  if (sf == 4 && wp <= MH_THREADS) {
    // x4 fast path: a thread owns one source column i (4 output pixels per row, one 16-byte store) and walks output rows
    const int i = tid % wp, rs = tid / wp, RS = MH_THREADS / wp;
    if (rs < RS) {
      const QuadX qx = make_quadx(i, wp);
      for (int row = rs; row < R * Ho; row += RS) {
        const int r = row / Ho, y = row % Ho;
        const float* pl = lg + r * hw;
        const Tap ty = make_tap(y, rscale, hp);
        float z[4];
        quad4(pl + ty.i0 * wp, pl + ty.i1 * wp, qx, ty.l0, ty.l1, z);
        __stcs(reinterpret_cast<float4*>(out + (int64_t)row * Wo + 4 * i),
               make_float4(sigmoid_fast(z[0]), sigmoid_fast(z[1]), sigmoid_fast(z[2]), sigmoid_fast(z[3])));
        if (lout) *reinterpret_cast<float4*>(lout + (int64_t)row * Wo + 4 * i) = make_float4(z[0], z[1], z[2], z[3]);
      }
    }
    return;
  }
  for (int row = warp; row < R * Ho; row += MH_THREADS / 32) {
    const int r = row / Ho, y = row % Ho;
    const float* pl = lg + r * hw;
    Tap ty = make_tap(y, rscale, hp);
    const float* r0 = pl + ty.i0 * wp;
    const float* r1 = pl + ty.i1 * wp;
    for (int x = lane; x < Wo; x += 32) {
      Tap tx = make_tap(x, rscale, wp);
      const float z = bilerp(r0[tx.i0], r0[tx.i1], r1[tx.i0], r1[tx.i1], tx.l0, tx.l1, ty.l0, ty.l1);
      __stcs(out + (int64_t)row * Wo + x, sigmoid_fast(z));
      if (lout) lout[(int64_t)row * Wo + x] = z;
    }
  }
}
int mask_head(const float* queries, const float* tokens, float* mask_pred, float* logits_out, int B, int L, int layer0, int nq,
              int D, int hp, int wp, int sf, cudaStream_t s, bool precise) {
  if (B == 0) return SMK_OK;
  SMK_REQUIRE(D % MH_KC == 0 && B <= 65535, "mask_head: D=%d unsupported", D);
  const int hw = hp * wp, NQ4 = (hw + 3) / 4;
  // layers per CTA: as many as keep the row count <= 40 and the tile count within MH_MAXT per thread
  int Lg = 40 / nq;
  if (Lg < 1) Lg = 1;
  if (Lg > L) Lg = L;
  while (Lg > 1 && ((Lg * nq + 7) / 8) * NQ4 > MH_MAXT_CAP * MH_THREADS) --Lg;
  const int n_tiles = ((Lg * nq + 7) / 8) * NQ4;
  SMK_REQUIRE(n_tiles <= MH_MAXT_CAP * MH_THREADS, "mask_head: %d queries x %d patches exceed the register tiling", nq, hw);
  const size_t smem = ((size_t)Lg * nq * hw + (size_t)((Lg * nq + 7) / 8 * 8) * MH_LD + (size_t)NQ4 * 4 * MH_LD) * sizeof(float);
  SMK_REQUIRE(smem <= 220 * 1024, "mask_head: %zu bytes of shared memory needed", smem);
  auto kern = precise ? (n_tiles <= MH_THREADS ? mask_head_kernel<1, true> : mask_head_kernel<MH_MAXT_CAP, true>)
                      : (n_tiles <= MH_THREADS ? mask_head_kernel<1, false> : mask_head_kernel<MH_MAXT_CAP, false>);
  if (smem > 48 * 1024) SMK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    // algorithmic bytes: read tokens + queries, write the [nq, hp*sf, wp*sf] probability planes
    ProfScope prof(PROF_MASK_HEAD, (double)B * L * ((double)(hp * wp + nq) * D + (double)nq * hp * sf * wp * sf) * 4.0, s);
    SMK_CHECK_CUDA(launch_pdl(kern, dim3((L + Lg - 1) / Lg, B), dim3(MH_THREADS), smem, s, queries, tokens, mask_pred, logits_out, B, L, Lg, nq,
                              D, hp, wp, sf, layer0));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// objectness tail: out[r] = sigmoid(dot(h[r,:], w) + b)   (last Linear(384→1) of maskformer.py:254-268 + :239)
__global__ void __launch_bounds__(256)
rowdot_sigmoid_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, int64_t rows, int D) {
  pdl_wait();
  pdl_trigger();
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(h[row * D + d], w[d], s);
  s = warp_sum(s);
  if (lane == 0) out[row] = sigmoidf_(s + bias[0]);
}
int rowdot_sigmoid(const float* h, const float* w, const float* bias, float* out, int64_t rows, int D, cudaStream_t s) {
  if (rows == 0) return SMK_OK;
  {
    ProfScope prof(PROF_OTHER, (double)rows * D * 4.0, s);
    SMK_CHECK_CUDA(launch_pdl(rowdot_sigmoid_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, s, h, w, bias, out, rows, D));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// objectness comes out as [L,B,nq]; the interface wants [B,L,nq] (maskformer.py:238 permute)
__global__ void permute_lb_kernel(const float* __restrict__ in, float* __restrict__ out, int L, int B, int n) {
  pdl_wait();
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)L * B * n) return;
  const int e = (int)(i % n);
  const int64_t lb = i / n;
  const int b = (int)(lb % B), l = (int)(lb / B);
  out[((int64_t)b * L + l) * n + e] = in[i];
}
int permute_lb(const float* in, float* out, int L, int B, int n, cudaStream_t s) {
  const int64_t tot = (int64_t)L * B * n;
  if (tot == 0) return SMK_OK;
  {
    ProfScope prof(PROF_OTHER, (double)tot * 8.0, s);
    SMK_CHECK_CUDA(launch_pdl(permute_lb_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, s, in, out, L, B, n));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// features[b,:] = mean_q queries_last[b,q,:]   (maskformer.py:203)
__global__ void query_mean_kernel(const float* __restrict__ qlast, float* __restrict__ out, int nq, int D) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int qi = 0; qi < nq; ++qi) s += qlast[((int64_t)b * nq + qi) * D + d];
    out[(int64_t)b * D + d] = s / (float)nq;
  }
}
int query_mean(const float* qlast, float* out, int B, int nq, int D, cudaStream_t s) {
  if (B == 0) return SMK_OK;
  {
    ProfScope prof(PROF_OTHER, (double)B * (nq + 1) * D * 4.0, s);
    SMK_CHECK_CUDA(launch_pdl(query_mean_kernel, dim3(B), dim3(128), 0, s, qlast, out, nq, D));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// Bicubic resample of the learned position grid (vision_transformer.py:377-401; ATen upsample_bicubic2d,
// align_corners=False, A = -0.75), run once per image geometry at model creation.
__device__ __forceinline__ float cubic1(float x) { return ((-0.75f + 2.f) * x - (-0.75f + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x) { return ((-0.75f * x - 5.f * -0.75f) * x + 8.f * -0.75f) * x - 4.f * -0.75f; }
__global__ void pos_bicubic_kernel(const float* __restrict__ pos /*[1+g*g, D]*/, float* __restrict__ out /*[1+hp*wp, D]*/,
                                   int g, int hp, int wp, int D) {
  const int t = blockIdx.x;   // output token
  if (t == 0) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) out[d] = pos[d];
    return;
  }
  const int oy = (t - 1) / wp, ox = (t - 1) % wp;
  const float sy = (float)g / (float)hp, sx = (float)g / (float)wp;
  const float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
  const int iy = (int)floorf(fy), ix = (int)floorf(fx);
  const float ty = fy - (float)iy, tx = fx - (float)ix;
  float wy[4] = {cubic2(ty + 1.f), cubic1(ty), cubic1(1.f - ty), cubic2(2.f - ty)};
  float wx[4] = {cubic2(tx + 1.f), cubic1(tx), cubic1(1.f - tx), cubic2(2.f - tx)};
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < 4; ++i) {
      const int yy = min(max(iy - 1 + i, 0), g - 1);
      float r = 0.f;
      for (int j = 0; j < 4; ++j) {
        const int xx = min(max(ix - 1 + j, 0), g - 1);
        r += wx[j] * pos[(int64_t)(1 + yy * g + xx) * D + d];
      }
      acc += wy[i] * r;
    }
    out[(int64_t)t * D + d] = acc;
  }
}
int pos_bicubic(const float* pos, float* out, int g, int hp, int wp, int D, cudaStream_t s) {
  pos_bicubic_kernel<<<1 + hp * wp, 128, 0, s>>>(pos, out, g, hp, wp, D);
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace smk

using namespace smk;

extern "C" int smk_split_q8(const float* x, int64_t ldx, void* out, int64_t rows, int K, int is_weight, void* stream) {
  SMK_REQUIRE(x && out && rows >= 0 && K > 0 && ldx >= K, "smk_split_q8: bad arguments");
  return split_q8(x, ldx, (__half*)out, rows, K, is_weight, (cudaStream_t)stream);
}

extern "C" int smk_gemm_f32(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc, int M, int N,
                            int K, int epilogue, void* stream) {
  SMK_REQUIRE(A && W && C && M >= 0 && N > 0 && K > 0, "smk_gemm_f32: bad arguments");
  return gemm_f32(A, lda, W, K, bias, C, ldc, M, N, K, epilogue, (cudaStream_t)stream);
}

extern "C" int smk_layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int D, float eps,
                             int out_bf16, void* stream) {
  SMK_REQUIRE(x && gamma && beta && y && rows >= 0, "smk_layernorm: bad arguments");
  if (out_bf16) return layernorm_bf16(x, nullptr, gamma, beta, (__nv_bfloat16*)y, nullptr, nullptr, rows, D, eps, (cudaStream_t)stream, nullptr);
  return layernorm_f32(x, nullptr, gamma, beta, (float*)y, nullptr, rows, D, eps, (cudaStream_t)stream);
}

extern "C" int smk_layernorm_f16(const float* x, const float* gamma, const float* beta, void* y, int64_t ldy, float* y32, int64_t rows, int D,
                                 float eps, int lo_kind, void* stream) {
  SMK_REQUIRE(x && gamma && beta && y && rows >= 0 && lo_kind >= 0 && lo_kind <= 2 && ldy >= (lo_kind ? 2 : 1) * (int64_t)D, "smk_layernorm_f16: bad arguments");
  return layernorm_f16(x, gamma, beta, (__half*)y, lo_kind ? (__half*)y + D : nullptr, ldy, y32, nullptr, nullptr, rows, D, eps, (cudaStream_t)stream, 0,
                       lo_kind == 2);
}

extern "C" int smk_im2col_f16(const void* x, int is_u8, void* cols, int B, int H, int W, int P, const float* mean_std, int q8, void* stream) {
  SMK_REQUIRE(x && cols && B >= 0 && H > 0 && W > 0 && P > 0, "smk_im2col_f16: bad arguments");
  const int hp = (H + P - 1) / P, wp = (W + P - 1) / P;
  if (is_u8) return im2col_split_f16<uint8_t>((const uint8_t*)x, (__half*)cols, B, H, W, P, hp, wp, mean_std, (cudaStream_t)stream, q8);
  return im2col_split_f16<float>((const float*)x, (__half*)cols, B, H, W, P, hp, wp, nullptr, (cudaStream_t)stream, q8);
}

extern "C" int smk_attention(const void* q, const void* k, const void* v, void* o, int batch, int heads, int dh, int Lq, int Lk,
                             int64_t q_bstride, int64_t ldq, int64_t k_bstride, int64_t ldk, int64_t v_bstride, int64_t ldv,
                             int64_t o_bstride, int64_t ldo, float scale, int is_bf16, void* stream) {
  SMK_REQUIRE(q && k && v && o, "smk_attention: null pointer");
  if (is_bf16)
    return attention<__nv_bfloat16, __nv_bfloat16>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)o, batch,
                                    heads, dh, Lq, Lk, q_bstride, ldq, k_bstride, ldk, v_bstride, ldv, o_bstride, ldo, scale,
                                    (cudaStream_t)stream);
  return attention<float, float>((const float*)q, (const float*)k, (const float*)v, (float*)o, batch, heads, dh, Lq, Lk, q_bstride, ldq,
                          k_bstride, ldk, v_bstride, ldv, o_bstride, ldo, scale, (cudaStream_t)stream);
}

extern "C" int smk_cast_bf16(const float* in, void* out, int64_t n, void* stream) {
  SMK_REQUIRE(in && out && n >= 0, "smk_cast_bf16: bad arguments");
  return cast_bf16(in, (__nv_bfloat16*)out, n, (cudaStream_t)stream);
}

extern "C" int smk_split3(const float* x, int64_t rows, int K, void* out, int is_weight, void* stream) {
  SMK_REQUIRE(x && out && rows >= 0 && K > 0, "smk_split3: bad arguments");
  if (is_weight) return split3_weight(x, (__nv_bfloat16*)out, rows, K, (cudaStream_t)stream);
  return split3_act(x, K, nullptr, 0, (__nv_bfloat16*)out, nullptr, rows, K, (cudaStream_t)stream);
}
