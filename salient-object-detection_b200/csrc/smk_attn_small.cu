// Softmax attention for FEW queries per (image, head): the decoder's self-attention (nq x nq) and cross-attention
// (nq queries x hw patch keys), transformer_decoder.py:271-291 via nn.MultiheadAttention; head dim 64.
//
// With nq = 10 / 20 query rows a 128-row tcgen05 tile is 84-92 % padding and the persistent TMEM pipeline of
// smk_attn_tc.cu pays ~5 us of serialised latency per (image, head) item (40 us self- / 56 us cross-attention per decoder
// layer at batch 256, ncu in profiles/r01_ncu.md).  Here one small CTA (2 warps) owns one (image, head): K and V go to
// shared memory with 16-byte cp.async (coalesced 128-byte rows), each warp takes 16 query rows and runs
//   S = Q·K^T   mma.sync.m16n8k16 bf16 → fp32, scores for ALL keys stay in registers (<= 256 keys)
//   softmax     in registers (quad shuffles), exp2 with the scale folded in
//   O = P·V     P re-used straight from the S accumulator fragments (bf16), V fragments by ldmatrix.trans
// 1536 CTAs, 3 per SM: the op becomes HBM-bound on the K/V read (77 MB per layer).
#include <type_traits>

#include "smk_mma.cuh"

namespace smk {

namespace {

using namespace mma;

constexpr int AS_DH = 64, AS_LD = 72;          // smem row stride in bf16 elements (144 B: conflict-free ldmatrix)
constexpr int AS_THREADS = 64, AS_MAXQ = 32;

struct AttnSmallParams {
  const __nv_bfloat16 *q, *k, *v;
  void* out;
  int64_t ldq, ldk, ldv, ldo;
  int Lq, Lk, kv_rows, kv_row0, heads, out_mode;   // out_mode: 0 bf16, 1 fp32, 2 bf16x3 split [hi | hi | lo]
  float scale_log2e;
  int q_f32;                                       // q points to fp32 rows (ldq in floats); rounded to the operand type when staged
};

// NT = number of 8-key score tiles held in registers (keys padded to 8*NT, a multiple of 16)
// kF16: q / k / v (and P) are fp16 instead of bf16 (fp16s mode: 11 significant bits; the outputs stay bf16 / fp32 / bf16 split)
template <int NT, bool kF16>
__global__ void __launch_bounds__(AS_THREADS)
attn_small_kernel(const AttnSmallParams p) {
  using T16 = typename std::conditional<kF16, __half, __nv_bfloat16>::type;
  extern __shared__ __align__(16) uint8_t as_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(as_smem);   // [32][AS_LD]
  __nv_bfloat16* sK = sQ + AS_MAXQ * AS_LD;                         // [8*NT][AS_LD]
  __nv_bfloat16* sV = sK + 8 * NT * AS_LD;                          // [8*NT][AS_LD]
  pdl_wait();
  pdl_trigger();
  const int item = blockIdx.x, b = item / p.heads, h = item % p.heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int LkP = 8 * NT;

  // ---- stage Q, K, V (rows beyond the valid ones are zero: they add 0 to every dot product) ----
  {
    if (p.q_f32) {
      const float* qg = reinterpret_cast<const float*>(p.q) + (int64_t)b * p.Lq * p.ldq + h * AS_DH;
      for (int i = tid; i < AS_MAXQ * 8; i += AS_THREADS) {
        const int r = i >> 3, c = i & 7;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (r < p.Lq) {
          const float4 a = *reinterpret_cast<const float4*>(qg + (int64_t)r * p.ldq + c * 8), bq = *reinterpret_cast<const float4*>(qg + (int64_t)r * p.ldq + c * 8 + 4);
          u = make_uint4(Pack16<T16>::pack(a.x, a.y), Pack16<T16>::pack(a.z, a.w), Pack16<T16>::pack(bq.x, bq.y), Pack16<T16>::pack(bq.z, bq.w));
        }
        *reinterpret_cast<uint4*>(sQ + r * AS_LD + c * 8) = u;
      }
    } else {
      const __nv_bfloat16* qg = p.q + (int64_t)b * p.Lq * p.ldq + h * AS_DH;
      for (int i = tid; i < AS_MAXQ * 8; i += AS_THREADS) {
        const int r = i >> 3, c = i & 7;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sQ + r * AS_LD + c * 8);
        if (r < p.Lq) cp_async16(dst, qg + (int64_t)r * p.ldq + c * 8);
        else *reinterpret_cast<uint4*>(sQ + r * AS_LD + c * 8) = make_uint4(0, 0, 0, 0);
      }
    }
    const int64_t kv_row = (int64_t)b * p.kv_rows + p.kv_row0;
    const __nv_bfloat16* kg = p.k + kv_row * p.ldk + h * AS_DH;
    const __nv_bfloat16* vg = p.v + kv_row * p.ldv + h * AS_DH;
    // two cp.async groups: {Q, K} then {V} — the scores and the softmax run while V is still in flight
    for (int i = tid; i < LkP * 8; i += AS_THREADS) {
      const int r = i >> 3, c = i & 7;
      if (r < p.Lk) cp_async16((uint32_t)__cvta_generic_to_shared(sK + r * AS_LD + c * 8), kg + (int64_t)r * p.ldk + c * 8);
      else *reinterpret_cast<uint4*>(sK + r * AS_LD + c * 8) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
    for (int i = tid; i < LkP * 8; i += AS_THREADS) {
      const int r = i >> 3, c = i & 7;
      if (r < p.Lk) cp_async16((uint32_t)__cvta_generic_to_shared(sV + r * AS_LD + c * 8), vg + (int64_t)r * p.ldv + c * 8);
      else *reinterpret_cast<uint4*>(sV + r * AS_LD + c * 8) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
  }
  const int r0 = warp * 16;
  if (r0 >= p.Lq) {                             // this warp's 16 query rows are all padding: only see the V group through
    cp_async_wait<0>();
    __syncthreads();
    return;
  }
  const int g = lane >> 2, t = lane & 3;

  // ---- S = Q·K^T (fragment layouts of mma.m16n8k16: A row g / g+8, cols 2t.. ; B k 2t.., col g; C row g / g+8, cols 2t, 2t+1) ----
  uint32_t qa[4][4];
  {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sQ + (r0 + (lane & 15)) * AS_LD + (lane >> 4) * 8);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(base + kk * 32, qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3]);
  }
  float s[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    // one ldmatrix.x4 = B fragments of two k-steps: matrices (keys 8n.., d 16kk), (d 16kk+8), (d 16kk+16), (d 16kk+24)
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sK + (8 * n + (lane & 7)) * AS_LD + (lane >> 3) * 8);
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(base + kp * 64, b0, b1, b2, b3);
      mma_16<kF16>(s[n], qa[2 * kp][0], qa[2 * kp][1], qa[2 * kp][2], qa[2 * kp][3], b0, b1);
      mma_16<kF16>(s[n], qa[2 * kp + 1][0], qa[2 * kp + 1][1], qa[2 * kp + 1][2], qa[2 * kp + 1][3], b2, b3);
    }
  }

  // ---- softmax over the valid keys; rows g (c0, c1) and g + 8 (c2, c3) ----
  const float sc = p.scale_log2e;
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int key = 8 * n + 2 * t;
    if (key < p.Lk) { m0 = fmaxf(m0, s[n][0]); m1 = fmaxf(m1, s[n][2]); }
    if (key + 1 < p.Lk) { m0 = fmaxf(m0, s[n][1]); m1 = fmaxf(m1, s[n][3]); }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  const float ms0 = m0 * sc, ms1 = m1 * sc;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int key = 8 * n + 2 * t;
    s[n][0] = key < p.Lk ? ex2f(fmaf(s[n][0], sc, -ms0)) : 0.f;
    s[n][1] = key + 1 < p.Lk ? ex2f(fmaf(s[n][1], sc, -ms0)) : 0.f;
    s[n][2] = key < p.Lk ? ex2f(fmaf(s[n][2], sc, -ms1)) : 0.f;
    s[n][3] = key + 1 < p.Lk ? ex2f(fmaf(s[n][3], sc, -ms1)) : 0.f;
    l0 += s[n][0] + s[n][1];
    l1 += s[n][2] + s[n][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

  cp_async_wait<0>();
  __syncthreads();                              // V (second cp.async group) has landed for every thread's copies
  // ---- O = P·V: the C fragments of score tiles 2j, 2j+1 are exactly the A fragment of k-step j ----
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
  for (int j = 0; j < NT / 2; ++j) {
    const uint32_t a0 = Pack16<T16>::pack(s[2 * j][0], s[2 * j][1]), a1 = Pack16<T16>::pack(s[2 * j][2], s[2 * j][3]);
    const uint32_t a2 = Pack16<T16>::pack(s[2 * j + 1][0], s[2 * j + 1][1]), a3 = Pack16<T16>::pack(s[2 * j + 1][2], s[2 * j + 1][3]);
    // ldmatrix.x4.trans: matrices (keys 16j.., dims 8d), (keys 16j+8.., dims 8d), (keys 16j.., dims 8d+8), (keys 16j+8.., dims 8d+8)
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sV + (16 * j + (lane & 15)) * AS_LD + (lane >> 4) * 8);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(base + dp * 32, b0, b1, b2, b3);
      mma_16<kF16>(o[2 * dp], a0, a1, a2, a3, b0, b1);
      mma_16<kF16>(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
    }
  }

  // ---- O / rowsum → out[b*Lq + row, h*64 + 8d + 2t .. +1] ----
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int row_a = r0 + g, row_b = r0 + g + 8;
  const int D = p.heads * AS_DH;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = half ? row_b : row_a;
    if (row >= p.Lq) continue;
    const float inv = half ? i1 : i0;
    const int64_t grow = (int64_t)b * p.Lq + row;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const float x0 = o[d][2 * half] * inv, x1 = o[d][2 * half + 1] * inv;
      const int col = h * AS_DH + 8 * d + 2 * t;
      if (p.out_mode == 1) {
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + grow * p.ldo + col) = make_float2(x0, x1);
      } else {
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + grow * p.ldo + col;
        const __nv_bfloat162 hi = __floats2bfloat162_rn(x0, x1);
        *reinterpret_cast<__nv_bfloat162*>(orow) = hi;
        if (p.out_mode == 2) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(x0 - __low2float(hi), x1 - __high2float(hi));
          *reinterpret_cast<__nv_bfloat162*>(orow + D) = hi;
          *reinterpret_cast<__nv_bfloat162*>(orow + 2 * D) = lo;
        }
      }
    }
  }
}

template <int NT, bool kF16>
int launch_small(const AttnSmallParams& p, int B, cudaStream_t s) {
  const size_t smem = (size_t)(AS_MAXQ + 16 * NT) * AS_LD * sizeof(__nv_bfloat16);
  static DeviceOnce attr_set;
  if (smem > 48 * 1024 && attr_set.first()) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<NT, kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  {
    ProfScope prof(PROF_ATTENTION_TC, 4.0 * p.Lq * p.Lk * AS_DH * p.heads * B, s);
    SMK_CHECK_CUDA(launch_pdl(attn_small_kernel<NT, kF16>, dim3((unsigned)(B * p.heads)), dim3(AS_THREADS), smem, s, p));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace

// Same argument meaning as attention_tc_general (smk_attn_tc.cu); Lq <= 32, Lk <= 256.
int attention_small(const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v, int64_t ldv,
                    int kv_rows, int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq, int Lk, int heads, float scale,
                    cudaStream_t s, int f16, int q_f32) {
  SMK_REQUIRE(Lq >= 1 && Lq <= AS_MAXQ && Lk >= 1 && Lk <= 256, "attention_small: Lq=%d (1..32) / Lk=%d (1..256) not supported", Lq, Lk);
  SMK_REQUIRE(B >= 1 && heads >= 1 && (int64_t)B * heads < (1 << 30), "attention_small: bad batch/heads");
  SMK_REQUIRE(out_mode >= 0 && out_mode <= 2 && (out_mode != 2 || ldo >= 3 * (int64_t)heads * AS_DH), "attention_small: bad output mode / ldo");
  SMK_REQUIRE(!q_f32 || ldq % 4 == 0, "attention_small: fp32 q rows must be 16-byte aligned");
  SMK_REQUIRE((q_f32 || ldq % 8 == 0) && ldk % 8 == 0 && ldv % 8 == 0 && ((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0,
              "attention_small: q/k/v rows must be 16-byte aligned");
  SMK_REQUIRE(ldo % 2 == 0 && ((uintptr_t)out % 8) == 0, "attention_small: output must be 8-byte aligned with an even row stride");
  AttnSmallParams p{q, k, v, out, ldq, ldk, ldv, ldo, Lq, Lk, kv_rows, kv_row0, heads, out_mode, scale * 1.4426950408889634f, q_f32};
  const int nt = (Lk + 15) / 16 * 2;
  if (f16) {
    if (nt <= 4) return launch_small<4, true>(p, B, s);
    if (nt <= 26) return launch_small<26, true>(p, B, s);
    return launch_small<32, true>(p, B, s);
  }
  if (nt <= 4) return launch_small<4, false>(p, B, s);
  if (nt <= 26) return launch_small<26, false>(p, B, s);
  return launch_small<32, false>(p, B, s);
}

}  // namespace smk

extern "C" int smk_attention_small(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int kv_rows,
                                   int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq, int Lk, int heads, float scale,
                                   void* stream) {
  SMK_REQUIRE(q && k && v && out, "smk_attention_small: null pointer");
  return smk::attention_small((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, kv_rows, kv_row0,
                              out, ldo, out_mode, B, Lq, Lk, heads, scale, (cudaStream_t)stream, 0, 0);
}

/* the same with fp16 q / k / v (outputs unchanged: bf16 / fp32 / bf16 split) */
extern "C" int smk_attention_small_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int kv_rows,
                                       int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq, int Lk, int heads, float scale,
                                       int q_f32, void* stream) {
  SMK_REQUIRE(q && k && v && out, "smk_attention_small_f16: null pointer");
  return smk::attention_small((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, kv_rows, kv_row0,
                              out, ldo, out_mode, B, Lq, Lk, heads, scale, (cudaStream_t)stream, 1, q_f32);
}
