// Multi-head softmax attention with an online softmax over 64-key blocks (head dim 64), for the cases the single-tile
// tcgen05 kernel (smk_attn_tc.cu, <= 256 keys, bf16 operands) does not cover:
//   * long sequences: 384x384 images (577 tokens), ViT-S/8 (785 tokens), vision_transformer.py:110-130;
//   * the bf16x3 parity mode (kSplit): Q, K, V arrive as bf16 hi + lo parts (the split epilogue of the projection GEMM) and
//     every product is the 3-term hi·hi + hi·lo + lo·hi with fp32 accumulate; P is split in registers the same way, so the
//     attention is ~fp32-accurate while still running on tensor cores.
// One CTA = 64 query rows of one (image, head), 4 warps x 16 rows; K / V blocks are double-buffered in shared memory with
// 16-byte cp.async; S, the running max / sum and O stay in registers (mma.sync m16n8k16 bf16 → fp32).
#include <type_traits>

#include "smk_mma.cuh"

namespace smk {

namespace {

using namespace mma;

constexpr int FA_DH = 64, FA_LD = 72, FA_BM = 64, FA_BN = 64, FA_THREADS = 128;
constexpr int FA_TILE = FA_BN * FA_LD;   // elements of one staged 64 x 64 tile

struct AttnFaParams {
  const __nv_bfloat16 *q, *q_lo, *k, *k_lo, *v, *v_lo;   // *_lo only with kSplit
  void* out;
  int64_t ldq, ldk, ldv, ldo;
  int Lq, Lk, q_rows, kv_rows, kv_row0, heads, out_mode;   // out_mode: 0 bf16, 1 fp32, 2 bf16x3 split [hi | hi | lo]
  float scale_log2e;
};

// rows [row0, row0 + 64) x 64 columns of a bf16 matrix → smem tile [64][FA_LD]; rows >= n_valid become zero
__device__ __forceinline__ void stage_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int row0, int n_valid, int tid) {
#pragma unroll
  for (int i = tid; i < FA_BN * 8; i += FA_THREADS) {
    const int r = i >> 3, c = i & 7;
    __nv_bfloat16* d = dst + r * FA_LD + c * 8;
    if (row0 + r < n_valid) cp_async16((uint32_t)__cvta_generic_to_shared(d), src + (int64_t)(row0 + r) * ld + c * 8);
    else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
  }
}

// kF16 (without kSplit): q / k / v / P are fp16 instead of bf16 (fp16s mode); output modes 0 / 3 then write fp16
template <bool kSplit, bool kF16>
__global__ void __launch_bounds__(FA_THREADS)
attn_fa_kernel(const AttnFaParams p) {
  using T16 = typename std::conditional<kF16, __half, __nv_bfloat16>::type;
  extern __shared__ __align__(16) uint8_t fa_smem[];
  constexpr int kParts = kSplit ? 2 : 1;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(fa_smem);    // [kParts][64][LD]
  __nv_bfloat16* sK = sQ + kParts * FA_TILE;                         // [2 stages][kParts][64][LD]
  __nv_bfloat16* sV = sK + 2 * kParts * FA_TILE;                     // [2 stages][kParts][64][LD]
  pdl_wait();
  pdl_trigger();
  const int qt = blockIdx.x, item = blockIdx.y, b = item / p.heads, h = item % p.heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qt * FA_BM;
  const int64_t q_row = (int64_t)b * p.q_rows, kv_row = (int64_t)b * p.kv_rows + p.kv_row0;
  const int n_blocks = (p.Lk + FA_BN - 1) / FA_BN;

  auto load_kv = [&](int blk, int stage) {
    const int key0 = blk * FA_BN;
    stage_tile(sK + (stage * kParts) * FA_TILE, p.k + kv_row * p.ldk + h * FA_DH, p.ldk, key0, p.Lk, tid);
    stage_tile(sV + (stage * kParts) * FA_TILE, p.v + kv_row * p.ldv + h * FA_DH, p.ldv, key0, p.Lk, tid);
    if constexpr (kSplit) {
      stage_tile(sK + (stage * kParts + 1) * FA_TILE, p.k_lo + kv_row * p.ldk + h * FA_DH, p.ldk, key0, p.Lk, tid);
      stage_tile(sV + (stage * kParts + 1) * FA_TILE, p.v_lo + kv_row * p.ldv + h * FA_DH, p.ldv, key0, p.Lk, tid);
    }
  };
  stage_tile(sQ, p.q + q_row * p.ldq + h * FA_DH, p.ldq, q0, p.Lq, tid);
  if constexpr (kSplit) stage_tile(sQ + FA_TILE, p.q_lo + q_row * p.ldq + h * FA_DH, p.ldq, q0, p.Lq, tid);
  load_kv(0, 0);
  cp_async_commit();

  const int r0 = warp * 16;
  uint32_t qa[kParts][4][4];
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // running max (scaled, log2 domain) and sum of rows g, g + 8
  const float sc = p.scale_log2e;

  for (int blk = 0; blk < n_blocks; ++blk) {
    const int stage = blk & 1;
    if (blk + 1 < n_blocks) {                 // prefetch the next block into the other stage (freed at the end of iteration blk-1)
      load_kv(blk + 1, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (blk == 0) {
#pragma unroll
      for (int part = 0; part < kParts; ++part) {
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(sQ + part * FA_TILE + (r0 + (lane & 15)) * FA_LD + (lane >> 4) * 8);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(base + kk * 32, qa[part][kk][0], qa[part][kk][1], qa[part][kk][2], qa[part][kk][3]);
      }
    }
    const __nv_bfloat16* kh = sK + (stage * kParts) * FA_TILE;
    const __nv_bfloat16* vh = sV + (stage * kParts) * FA_TILE;

    // ---- S = Q·K^T for this block: 8 tiles of 8 keys ----
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      const int off = (8 * n + (lane & 7)) * FA_LD + (lane >> 3) * 8;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(kh + off) + kp * 64, b0, b1, b2, b3);
        mma_16<kF16>(s[n], qa[0][2 * kp][0], qa[0][2 * kp][1], qa[0][2 * kp][2], qa[0][2 * kp][3], b0, b1);
        mma_16<kF16>(s[n], qa[0][2 * kp + 1][0], qa[0][2 * kp + 1][1], qa[0][2 * kp + 1][2], qa[0][2 * kp + 1][3], b2, b3);
        if constexpr (kSplit) {
          // q_lo · k_hi (same B fragments), then q_hi · k_lo
          mma_16<kF16>(s[n], qa[1][2 * kp][0], qa[1][2 * kp][1], qa[1][2 * kp][2], qa[1][2 * kp][3], b0, b1);
          mma_16<kF16>(s[n], qa[1][2 * kp + 1][0], qa[1][2 * kp + 1][1], qa[1][2 * kp + 1][2], qa[1][2 * kp + 1][3], b2, b3);
          ldmatrix_x4((uint32_t)__cvta_generic_to_shared(kh + FA_TILE + off) + kp * 64, b0, b1, b2, b3);
          mma_16<kF16>(s[n], qa[0][2 * kp][0], qa[0][2 * kp][1], qa[0][2 * kp][2], qa[0][2 * kp][3], b0, b1);
          mma_16<kF16>(s[n], qa[0][2 * kp + 1][0], qa[0][2 * kp + 1][1], qa[0][2 * kp + 1][2], qa[0][2 * kp + 1][3], b2, b3);
        }
      }
    }

    // ---- online softmax (log2 domain): keys beyond Lk are masked in the last block ----
    const int key_base = blk * FA_BN + 2 * t;
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int key = key_base + 8 * n;
      if (key >= p.Lk) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
      if (key + 1 >= p.Lk) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
      bm0 = fmaxf(bm0, fmaxf(s[n][0], s[n][1]));
      bm1 = fmaxf(bm1, fmaxf(s[n][2], s[n][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float n0 = fmaxf(m0, bm0 * sc), n1 = fmaxf(m1, bm1 * sc);     // a block always holds >= 1 valid key: finite
    const float a0 = ex2f(m0 - n0), a1 = ex2f(m1 - n1);                 // exp2(-inf) = 0 on the first block
    m0 = n0; m1 = n1;
    float bl0 = 0.f, bl1 = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[n][0] = ex2f(fmaf(s[n][0], sc, -n0));
      s[n][1] = ex2f(fmaf(s[n][1], sc, -n0));
      s[n][2] = ex2f(fmaf(s[n][2], sc, -n1));
      s[n][3] = ex2f(fmaf(s[n][3], sc, -n1));
      bl0 += s[n][0] + s[n][1];
      bl1 += s[n][2] + s[n][3];
    }
    l0 = fmaf(l0, a0, bl0);       // per-lane partial sums; reduced over the quad at the end
    l1 = fmaf(l1, a1, bl1);
#pragma unroll
    for (int d = 0; d < 8; ++d) { o[d][0] *= a0; o[d][1] *= a0; o[d][2] *= a1; o[d][3] *= a1; }

    // ---- O += P·V ----
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t ph[4], pl[4];
      if constexpr (kSplit) {
        pack2_split(s[2 * j][0], s[2 * j][1], ph[0], pl[0]);
        pack2_split(s[2 * j][2], s[2 * j][3], ph[1], pl[1]);
        pack2_split(s[2 * j + 1][0], s[2 * j + 1][1], ph[2], pl[2]);
        pack2_split(s[2 * j + 1][2], s[2 * j + 1][3], ph[3], pl[3]);
      } else {
        ph[0] = Pack16<T16>::pack(s[2 * j][0], s[2 * j][1]);
        ph[1] = Pack16<T16>::pack(s[2 * j][2], s[2 * j][3]);
        ph[2] = Pack16<T16>::pack(s[2 * j + 1][0], s[2 * j + 1][1]);
        ph[3] = Pack16<T16>::pack(s[2 * j + 1][2], s[2 * j + 1][3]);
      }
      const int off = (16 * j + (lane & 15)) * FA_LD + (lane >> 4) * 8;
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(vh + off) + dp * 32, b0, b1, b2, b3);
        mma_16<kF16>(o[2 * dp], ph[0], ph[1], ph[2], ph[3], b0, b1);
        mma_16<kF16>(o[2 * dp + 1], ph[0], ph[1], ph[2], ph[3], b2, b3);
        if constexpr (kSplit) {
          mma_16<kF16>(o[2 * dp], pl[0], pl[1], pl[2], pl[3], b0, b1);
          mma_16<kF16>(o[2 * dp + 1], pl[0], pl[1], pl[2], pl[3], b2, b3);
          ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(vh + FA_TILE + off) + dp * 32, b0, b1, b2, b3);
          mma_16<kF16>(o[2 * dp], ph[0], ph[1], ph[2], ph[3], b0, b1);
          mma_16<kF16>(o[2 * dp + 1], ph[0], ph[1], ph[2], ph[3], b2, b3);
        }
      }
    }
    __syncthreads();              // every warp is done with this stage before the next iteration's prefetch overwrites it
  }

  // ---- O / rowsum → out ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int D = p.heads * FA_DH;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = q0 + r0 + g + 8 * half;
    if (row >= p.Lq) continue;
    const float inv = half ? i1 : i0;
    const int64_t grow = q_row + row;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const float x0 = o[d][2 * half] * inv, x1 = o[d][2 * half + 1] * inv;
      const int col = h * FA_DH + 8 * d + 2 * t;
      if (p.out_mode == 1) {
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + grow * p.ldo + col) = make_float2(x0, x1);
      } else if (p.out_mode == 2) {     // bf16 split [hi | hi | lo] whatever the operand type (decoder out-projection input)
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + grow * p.ldo + col;
        uint32_t hi, lo;
        split16x2<__nv_bfloat16>(x0, x1, hi, lo);
        *reinterpret_cast<uint32_t*>(orow) = hi;
        *reinterpret_cast<uint32_t*>(orow + D) = hi;
        *reinterpret_cast<uint32_t*>(orow + 2 * D) = lo;
      } else {                          // 0: 16-bit (operand type), 3: [hi | lo] split in the operand type
        T16* orow = reinterpret_cast<T16*>(p.out) + grow * p.ldo + col;
        uint32_t hi, lo;
        split16x2<T16>(x0, x1, hi, lo);
        *reinterpret_cast<uint32_t*>(orow) = hi;
        if (p.out_mode == 3) *reinterpret_cast<uint32_t*>(orow + D) = lo;
      }
    }
  }
}

template <bool kSplit, bool kF16>
int launch_fa(const AttnFaParams& p, int B, cudaStream_t s) {
  const size_t smem = (size_t)(kSplit ? 2 : 1) * 5 * FA_TILE * sizeof(__nv_bfloat16);
  static DeviceOnce attr_set;
  if (smem > 48 * 1024 && attr_set.first()) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(attn_fa_kernel<kSplit, kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  {
    // tensor work actually issued: 3 MMAs per product in split mode; credited as the algorithmic 4·Lq·Lk·64 FLOP
    ProfScope prof(PROF_ATTENTION_TC, 4.0 * p.Lq * p.Lk * FA_DH * p.heads * B, s);
    SMK_CHECK_CUDA(launch_pdl(attn_fa_kernel<kSplit, kF16>, dim3((unsigned)((p.Lq + FA_BM - 1) / FA_BM), (unsigned)(B * p.heads)), dim3(FA_THREADS), smem, s, p));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace

// q / k / v: bf16 [rows, ld]; head h in columns [h*64, h*64+64).  Image b: queries at rows b*q_rows .. +Lq, keys / values at rows
// b*kv_rows + kv_row0 .. +Lk.  q_lo / k_lo / v_lo non-null → bf16x3 split mode (same leading dimensions as the hi parts).
int attention_fa(const __nv_bfloat16* q, const __nv_bfloat16* q_lo, int64_t ldq, const __nv_bfloat16* k, const __nv_bfloat16* k_lo, int64_t ldk,
                 const __nv_bfloat16* v, const __nv_bfloat16* v_lo, int64_t ldv, int q_rows, int kv_rows, int kv_row0, void* out, int64_t ldo,
                 int out_mode, int B, int Lq, int Lk, int heads, float scale, cudaStream_t s, int f16) {
  const bool split = q_lo != nullptr;
  SMK_REQUIRE(!(split && f16), "attention_fa: the split form takes bf16 parts");
  SMK_REQUIRE(Lq >= 1 && Lk >= 1 && B >= 1 && heads >= 1 && (int64_t)B * heads <= 65535, "attention_fa: bad sizes (B*heads <= 65535)");
  SMK_REQUIRE(split == (k_lo != nullptr) && split == (v_lo != nullptr), "attention_fa: give all three lo parts or none");
  SMK_REQUIRE(out_mode >= 0 && out_mode <= 3 && (out_mode != 2 || ldo >= 3 * (int64_t)heads * FA_DH) && (out_mode != 3 || ldo >= 2 * (int64_t)heads * FA_DH),
              "attention_fa: bad output mode / ldo");
  SMK_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, "attention_fa: leading dimensions must be multiples of 8");
  for (const void* ptr_ : {(const void*)q, (const void*)q_lo, (const void*)k, (const void*)k_lo, (const void*)v, (const void*)v_lo})
    SMK_REQUIRE(((uintptr_t)ptr_ % 16) == 0, "attention_fa: operands must be 16-byte aligned");
  SMK_REQUIRE(ldo % 2 == 0 && ((uintptr_t)out % 8) == 0, "attention_fa: output must be 8-byte aligned with an even row stride");
  AttnFaParams p{q, q_lo, k, k_lo, v, v_lo, out, ldq, ldk, ldv, ldo, Lq, Lk, q_rows, kv_rows, kv_row0, heads, out_mode,
                 scale * 1.4426950408889634f};
  if (f16) return launch_fa<false, true>(p, B, s);
  return split ? launch_fa<true, false>(p, B, s) : launch_fa<false, false>(p, B, s);
}

}  // namespace smk

extern "C" int smk_attention_fa(const void* q, const void* q_lo, int64_t ldq, const void* k, const void* k_lo, int64_t ldk, const void* v,
                                const void* v_lo, int64_t ldv, int q_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_mode,
                                int B, int Lq, int Lk, int heads, float scale, void* stream) {
  SMK_REQUIRE(q && k && v && out, "smk_attention_fa: null pointer");
  return smk::attention_fa((const __nv_bfloat16*)q, (const __nv_bfloat16*)q_lo, ldq, (const __nv_bfloat16*)k, (const __nv_bfloat16*)k_lo, ldk,
                           (const __nv_bfloat16*)v, (const __nv_bfloat16*)v_lo, ldv, q_rows, kv_rows, kv_row0, out, ldo, out_mode, B, Lq, Lk,
                           heads, scale, (cudaStream_t)stream, 0);
}
