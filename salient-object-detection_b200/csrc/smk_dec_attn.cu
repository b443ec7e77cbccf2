// Decoder self-attention in fp32 (transformer_decoder.py:271-281 via nn.MultiheadAttention: q = k = tgt + query_pos, v = tgt).
//
// nq x nq scores per (image, head) with nq = 10 / 20: 0.6 MFLOP per image and layer, but the most precision-sensitive
// contraction of the whole path — query_embed is N(0, 1), so the scores are large and the softmax is peaked: rounding q / k / v / P
// to bf16 moves the mask logits by 7e-2, to fp16 by 8e-3 (scripts/precision_emulation.py), against a 2e-2 budget for the whole
// model.  So this one runs on the CUDA cores in fp32 with expf: one CTA per (image, head), everything in shared memory.
// Output: the bf16x3 split [hi | hi | lo] the out-projection's split GEMM consumes (same layout as attention_small's mode 2).
#include "smk_common.cuh"
#include "smk_kernels.h"

namespace smk {

namespace {

constexpr int DSA_THREADS = 128, DSA_MAXQ = 32, DSA_DH = 64;

__global__ void __launch_bounds__(DSA_THREADS)
dec_self_attn_kernel(const float* __restrict__ qk, int64_t ldqk, const float* __restrict__ v, int64_t ldv, const float* __restrict__ v_sub,
                     __nv_bfloat16* __restrict__ out3, int nq, int heads, float scale) {
  __shared__ float sq[DSA_MAXQ][DSA_DH];
  __shared__ float sk[DSA_MAXQ][DSA_DH + 1];
  __shared__ float sv[DSA_MAXQ][DSA_DH];
  __shared__ float sp[DSA_MAXQ][DSA_MAXQ + 1];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x / heads, h = blockIdx.x % heads, D = heads * DSA_DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)b * nq;
  for (int i = tid; i < nq * (DSA_DH / 4); i += DSA_THREADS) {
    const int r = i >> 4, c = (i & 15) << 2;
    const float4 a = *reinterpret_cast<const float4*>(qk + (row0 + r) * ldqk + h * DSA_DH + c);
    const float4 kk = *reinterpret_cast<const float4*>(qk + (row0 + r) * ldqk + D + h * DSA_DH + c);
    float4 vv = *reinterpret_cast<const float4*>(v + (row0 + r) * ldv + h * DSA_DH + c);
    if (v_sub) {      // v is the V slice of a merged q|k|v projection of (tgt + query_pos): take query_pos · Wv^T out again (value = tgt)
      const float4 cs = __ldg(reinterpret_cast<const float4*>(v_sub + (int64_t)r * D + h * DSA_DH + c));
      vv.x -= cs.x; vv.y -= cs.y; vv.z -= cs.z; vv.w -= cs.w;
    }
    sq[r][c] = a.x * scale; sq[r][c + 1] = a.y * scale; sq[r][c + 2] = a.z * scale; sq[r][c + 3] = a.w * scale;   // q scaled first, as torch does
    sk[r][c] = kk.x; sk[r][c + 1] = kk.y; sk[r][c + 2] = kk.z; sk[r][c + 3] = kk.w;
    *reinterpret_cast<float4*>(&sv[r][c]) = vv;
  }
  __syncthreads();
  for (int e = tid; e < nq * nq; e += DSA_THREADS) {
    const int i = e / nq, j = e - i * nq;
    float s = 0.f;
#pragma unroll 16
    for (int d = 0; d < DSA_DH; ++d) s = fmaf(sq[i][d], sk[j][d], s);
    sp[i][j] = s;
  }
  __syncthreads();
  for (int i = warp; i < nq; i += DSA_THREADS / 32) {     // one warp per row: lanes over the (<= 32) keys
    const float s = lane < nq ? sp[i][lane] : -INFINITY;
    const float m = warp_max(s);
    const float e = lane < nq ? expf(s - m) : 0.f;
    const float l = warp_sum(e);
    if (lane < nq) sp[i][lane] = e / l;
  }
  __syncthreads();
  // O = P·V: each thread owns two adjacent head dims of some rows → 4-byte bf16x2 stores, 128 B per warp and part
  const int d2 = (tid & 31) * 2;
  for (int i = tid >> 5; i < nq; i += DSA_THREADS / 32) {
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < nq; ++j) {
      const float pj = sp[i][j];
      o0 = fmaf(pj, sv[j][d2], o0);
      o1 = fmaf(pj, sv[j][d2 + 1], o1);
    }
    __nv_bfloat16* orow = out3 + (row0 + i) * 3 * D + h * DSA_DH + d2;
    const __nv_bfloat162 hi = __floats2bfloat162_rn(o0, o1);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(o0 - __low2float(hi), o1 - __high2float(hi));
    *reinterpret_cast<__nv_bfloat162*>(orow) = hi;
    *reinterpret_cast<__nv_bfloat162*>(orow + D) = hi;
    *reinterpret_cast<__nv_bfloat162*>(orow + 2 * D) = lo;
  }
}

}  // namespace

// qk: [B*nq, ldqk] fp32 (q in columns [0, D), k in [D, 2D)); v: [B*nq, ldv] fp32; v_sub: nullptr or [nq, D] fp32 subtracted from every image's
// V rows (v = the V slice of ONE q|k|v projection of tgt + query_pos, v_sub = query_pos · Wv^T: value = tgt, transformer_decoder.py:277);
// out3: [B*nq, 3D] bf16 split [hi | hi | lo].
// (A one-warp-per-unit variant — lane = query row, everything in registers, no block barrier — was measured slower: 27.6 us against
// this kernel's 24.6 us warm at B = 256: its 2560-FMA serial chains per lane leave the schedulers idle.  This kernel is bound by
// shared-memory wavefronts — ~2800 per CTA for the scalar score / P·V reads — not by the 0.6 MFLOP of math.)
int dec_self_attention(const float* qk, int64_t ldqk, const float* v, int64_t ldv, const float* v_sub, __nv_bfloat16* out3, int B, int nq, int heads,
                       float scale, cudaStream_t s) {
  SMK_REQUIRE(nq >= 1 && nq <= DSA_MAXQ, "dec_self_attention: nq=%d not supported (1..32)", nq);
  SMK_REQUIRE(ldqk % 4 == 0 && ldv % 4 == 0 && ((uintptr_t)qk % 16) == 0 && ((uintptr_t)v % 16) == 0, "dec_self_attention: misaligned q/k/v");
  SMK_REQUIRE(((uintptr_t)v_sub % 16) == 0, "dec_self_attention: misaligned v_sub");
  if (B == 0) return SMK_OK;
  {
    ProfScope prof(PROF_ATTENTION, 4.0 * nq * nq * DSA_DH * heads * B, s);
    SMK_CHECK_CUDA(launch_pdl(dec_self_attn_kernel, dim3((unsigned)(B * heads)), dim3(DSA_THREADS), 0, s, qk, ldqk, v, ldv, v_sub, out3, nq, heads, scale));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace smk

extern "C" int smk_dec_self_attention(const float* qk, int64_t ldqk, const float* v, int64_t ldv, const float* v_sub, void* out3, int B, int nq,
                                      int heads, float scale, void* stream) {
  SMK_REQUIRE(qk && v && out3, "smk_dec_self_attention: null pointer");
  return smk::dec_self_attention(qk, ldqk, v, ldv, v_sub, (__nv_bfloat16*)out3, B, nq, heads, scale, (cudaStream_t)stream);
}
