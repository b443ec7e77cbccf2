// Decoder cross-attention against the encoder memory on tcgen05 (transformer_decoder.py:283-291 via nn.MultiheadAttention with
// query = tgt + query_pos, key = value = memory, pos = None), restructured so that the memory is never projected:
//
//   s_h = (x Wq_h^T + bq_h)(m Wk_h^T + bk_h)^T / 8  =  (x G_h + g_h) m^T + const(row)        G_h = Wq_h^T Wk_h / 8   [D x D]
//   o_h = softmax(s_h) (m Wv_h^T + bv_h)            =  (softmax(s_h) m) Wv_h^T + bv_h        (rows of the softmax sum to 1)
//   out = concat_h(o_h) Wo^T + bo                    =  Σ_h (softmax(s_h) m) M_h + (Wo bv + bo)   M_h = Wv_h^T Wo_h^T   [D x D]
//
// (the per-row constant x·… + bq_h·bk_h does not change a softmax).  With nq = 20 queries and 6 heads the 120 rows (q, h) of one
// image are ONE 128-row UMMA tile, the keys AND the values are the image's 196 final-LN patch tokens [196, D] — one shared-memory
// tile — and the per-layer K/V projection of the memory (56 % of the decoder's FLOPs, a 465 MB tensor at batch 256) disappears:
//
//   Q' = x [G_0 | … | G_5] + g           plain tcgen05 GEMM (fp16), rows (b, q), columns (h, c)  →  viewed as rows (b, q, h) of D
//   this kernel, per image:  S = Q' · T^T  (UMMA 128 x nk x 16, K = D, both operands from shared memory; Q' streams through a
//                            2-slot ring, T stays resident) → softmax per row (one thread per row, single pass over TMEM as in
//                            smk_attn_tc.cu) → P (fp16) back into TMEM → U = P · T (TS form, T as an MN-major operand, N = D)
//                            → U / rowsum as the fp16 split [hi | lo] of the out-projection GEMM's A operand
//   out = U [M_0; …; M_5] + (Wo bv + bo)  3-term split tcgen05 GEMM over K = heads·D
//
// Numerics (scripts/precision_emulation.py, /tmp experiments recorded in DESIGN.md §4): single-pass fp16 for S and for P·T moves the
// mask logits by 2.6e-3 (the 64-wide projected K / V of the textbook form are far more sensitive to rounding than the LayerNorm-ed
// tokens); U itself must keep ~16 bits (it is added to the residual stream) — hence the split output and the 3-term out GEMM.
#include <type_traits>

#include "smk_tc.cuh"

namespace smk {

using namespace tc;

constexpr int XA_BM = 128, XA_KC = 64;                 // query rows per image tile; channels per k-chunk (128 B of fp16)
constexpr int XA_MAXK = 208;                           // keys per image (padded to 16): 6 resident chunks of 208 x 128 B = 156 KB
constexpr int XA_THREADS = 192;                        // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 softmax / epilogue (one row per thread)
constexpr int XA_Q_BYTES = XA_BM * 128, XA_QSLOTS = 2;
constexpr int XA_STG_BYTES = 8192;                     // per epilogue warp: 32 rows x 128 B of hi + the same of lo
constexpr int XA_U_COL = 128;                          // U accumulator columns [128, 128 + D) — over S's tail, which is consumed first

struct XAttnParams {
  int rows, Lk, nk_pad, D, n_chunks;                   // valid query rows per image (nq·heads), keys, padded keys, channels, D / 64
  int q_rows_per_img, t_rows_per_img, t_row0;          // image b: Q' rows from b·q_rows_per_img, tokens from b·t_rows_per_img + t_row0
  int heads, nq, n_img;
  __half* out;                                         // [n_img·nq, 2·heads·D] fp16 split rows [hi | lo]; row (b, q), column (h, c)
  float log2e;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// optional phase trace of CTA 0 (tuning scripts only): [image < 8][role 0 = MMA warp, 1 = softmax warp 2][event 8] clock64 stamps
__device__ long long* g_xattn_trace = nullptr;
#define XA_TRACE(role, ev)                                                                                               \
  do {                                                                                                                   \
    if (g_xattn_trace && blockIdx.x == 0 && lane == 0 && it < 8) g_xattn_trace[(it * 2 + (role)) * 8 + (ev)] = clock64(); \
  } while (0)

namespace {

// kMaxUnits: 16-key units per row (13 = 193..208 keys: the 196 patch tokens of a 224 x 224 image); kExact: exactly that many
template <int kMaxUnits, bool kExact>
__global__ void __launch_bounds__(XA_THREADS, 1)
xattn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmT, const XAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int t_chunk_bytes = p.nk_pad * 128;            // one resident token chunk: nk_pad keys x 64 channels
  uint8_t* sT = smem;                                   // [n_chunks][nk_pad][128 B]
  uint8_t* sQ = smem + p.n_chunks * t_chunk_bytes;      // [2][128][128 B]  (t_chunk_bytes is a multiple of 2048: 1024-aligned)
  uint8_t* sStg = sQ + XA_QSLOTS * XA_Q_BYTES;          // [4 warps][hi 4 KB | lo 4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + 4 * XA_STG_BYTES);
  uint64_t *t_full = bars, *t_empty = bars + 1, *q_full = bars + 2, *q_empty = bars + 4, *s_full = bars + 6, *p_full = bars + 7,
           *u_full = bars + 8, *u_drained = bars + 9;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (p.n_img - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmT);
    mbar_init(t_full, 1); mbar_init(t_empty, 1);
    for (int i = 0; i < XA_QSLOTS; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(u_full, 1); mbar_init(u_drained, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_ptr);
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t qn = 0;                                  // Q' chunks issued so far (slot = qn & 1, use = qn >> 1)
      for (int it = 0; it < n_my; ++it) {
        const int b = blockIdx.x + it * gridDim.x;
        mbar_wait(t_empty, (it & 1) ^ 1);               // the previous image's P·T products have retired
        mbar_arrive_expect_tx(t_full, (uint32_t)(p.n_chunks * t_chunk_bytes));
        for (int c = 0; c < p.n_chunks; ++c) tma_load_2d(sT + c * t_chunk_bytes, &tmT, t_full, c * XA_KC, b * p.t_rows_per_img + p.t_row0);
        for (int c = 0; c < p.n_chunks; ++c, ++qn) {
          const int slot = qn & 1;
          mbar_wait(&q_empty[slot], ((qn >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[slot], XA_Q_BYTES);
          tma_load_2d(sQ + slot * XA_Q_BYTES, &tmQ, &q_full[slot], c * XA_KC, b * p.q_rows_per_img);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp convergent, one elected lane issues) =====
    const uint32_t idesc_s = idesc_f16_f32(XA_BM, p.nk_pad, 0, 0);
    const int n_ksteps = p.nk_pad / 16;
    uint32_t qn = 0;
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = it & 1;
      XA_TRACE(0, 0);
      if (it > 0) mbar_wait(u_drained, par ^ 1);        // TMEM is free: the previous image's U has been read out
      XA_TRACE(0, 1);
      mbar_wait(t_full, par);
      tc_fence_after_sync();
      XA_TRACE(0, 2);
      // S = Q' · T^T over the channel chunks
      for (int c = 0; c < p.n_chunks; ++c, ++qn) {
        const int slot = qn & 1;
        mbar_wait(&q_full[slot], (qn >> 1) & 1);
        tc_fence_after_sync();
        const uint64_t qd = smem_desc_k_sw128(smem_u32(sQ + slot * XA_Q_BYTES)), td = smem_desc_k_sw128(smem_u32(sT + c * t_chunk_bytes));
#pragma unroll
        for (int k = 0; k < XA_KC / 16; ++k) umma_bf16_ss_w(tmem_base, qd + (uint64_t)(2 * k), td + (uint64_t)(2 * k), idesc_s, (c | k) != 0);
        tc_commit_w(&q_empty[slot]);
      }
      tc_commit_w(s_full);
      XA_TRACE(0, 3);
      // U = P · T: P (fp16 pairs) from TMEM columns [0, nk_pad / 2), T as MN-major operand (64 channels per 128-byte row, chunk tiles
      // t_chunk_bytes apart = the descriptor's leading byte offset), N = D in pieces of <= 256 columns
      mbar_wait(p_full, par);
      tc_fence_after_sync();
      XA_TRACE(0, 4);
      for (int n0 = 0; n0 < p.D; n0 += 256) {
        const int nn = p.D - n0 < 256 ? p.D - n0 : 256;
        const uint32_t idesc_u = idesc_f16_f32(XA_BM, nn, 0, 1);
        const uint64_t vd = smem_desc_mn_sw128(smem_u32(sT + (n0 / XA_KC) * t_chunk_bytes), (uint32_t)t_chunk_bytes);
        for (int j = 0; j < n_ksteps; ++j)              // +2048 B per 16 keys → +128 in the descriptor's (addr >> 4) field
          umma_bf16_ts_w(tmem_base + (uint32_t)(XA_U_COL + n0), tmem_base + (uint32_t)(j * 8), vd + (uint64_t)(j * 128), idesc_u, j != 0);
      }
      tc_commit_w(u_full);
      tc_commit_w(t_empty);
      XA_TRACE(0, 5);
    }
  } else {
    // ===== softmax + epilogue: one thread per (query, head) row; TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                 // row within the image tile = q * heads + h
    const bool active = row < p.rows;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int nu = kExact ? kMaxUnits : (p.nk_pad >> 4);
    const float sc = p.log2e;
    const int64_t ld_out = 2 * (int64_t)p.heads * p.D, part = (int64_t)p.heads * p.D;
    for (int it = 0; it < n_my; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      const uint32_t par = it & 1;
      if (warp == 2) XA_TRACE(1, 0);
      mbar_wait(s_full, par);
      tc_fence_after_sync();
      if (warp == 2) XA_TRACE(1, 1);
      float rsum = 0.f;
      {
        // ---- single pass over S: online softmax per 16-key unit, P kept in registers (fp16 pairs); see smk_attn_tc.cu ----
        uint32_t pk[kMaxUnits][8];
        float mrec[kMaxUnits], usum[kMaxUnits];
        uint32_t va[16], vb[16];
        float M = -1.0e30f;
        auto unit = [&](uint32_t (&v)[16], uint32_t (&pu)[8], float& mr, float& us, int c0, bool last) {
          if (last) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j >= p.Lk) v[j] = 0xff800000u;
          }
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            m0 = max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            m1 = max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
          const float Mn = fmaxf(M, ceilf(fmaxf(m0, m1) * sc));
          M = Mn;
          mr = Mn;
          float e[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) e[j] = ex2_approx(fmaf(__uint_as_float(v[j]), sc, -Mn));
          us = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7])) + (((e[8] + e[9]) + (e[10] + e[11])) + ((e[12] + e[13]) + (e[14] + e[15])));
#pragma unroll
          for (int j = 0; j < 8; ++j) pu[j] = Pack16<__half>::pack(e[2 * j], e[2 * j + 1]);
        };
        tmem_ld_32x16(taddr, va);
#pragma unroll
        for (int u = 0; u < kMaxUnits; ++u) {
          if (u < nu) {
            if (u & 1) {
              tmem_ld_wait16(vb);
              if (u + 1 < nu) tmem_ld_32x16(taddr + (uint32_t)((u + 1) * 16), va);
              unit(vb, pk[u], mrec[u], usum[u], u * 16, u == nu - 1);
            } else {
              tmem_ld_wait16(va);
              if (u + 1 < nu) tmem_ld_32x16(taddr + (uint32_t)((u + 1) * 16), vb);
              unit(va, pk[u], mrec[u], usum[u], u * 16, u == nu - 1);
            }
          }
        }
        // bring every unit to the final maximum (exact power-of-two factors) and write P over S
#pragma unroll
        for (int u = 0; u < kMaxUnits; ++u) {
          if (u < nu) {
            const float fc = ex2_approx(mrec[u] - M);
            rsum = fmaf(usum[u], fc, rsum);
            const __half2 f2 = __float2half2_rn(fc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __half2 x = *reinterpret_cast<__half2*>(&pk[u][j]);
              x = __hmul2(x, f2);
              pk[u][j] = *reinterpret_cast<uint32_t*>(&x);
            }
            tmem_st_32x8(taddr + (uint32_t)(u * 8), pk[u]);
          }
        }
        tmem_st_wait();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 2) XA_TRACE(1, 2);
      // ---- epilogue: U / rowsum → fp16 split rows [hi | lo] of the out-projection operand (the kernel is bound by these writes:
      //      4 bytes per element, 47 MB per layer at batch 256) ----
      mbar_wait(u_full, par);
      tc_fence_after_sync();
      if (warp == 2) XA_TRACE(1, 3);
      const float inv = 1.0f / rsum;
      // A thread owns a row, but a row-per-lane global store touches 32 lines per instruction (32 LSU wavefronts): the 64-column
      // pieces go through a swizzled per-warp staging tile and leave with 8 lanes per row — 4 full 128-byte lines per instruction.
      const uint32_t stg = smem_u32(sStg + (warp - 2) * XA_STG_BYTES);
      const uint32_t wrow = stg + (uint32_t)lane * 128u, x7 = (uint32_t)(lane & 7);
      const int rr0 = lane >> 3, piece = lane & 7;
      for (int c0 = 0; c0 < p.D; c0 += 64) {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(taddr + (uint32_t)(XA_U_COL + c0), ra);
        tmem_ld_32x32(taddr + (uint32_t)(XA_U_COL + c0 + 32), rb);
        tmem_ld_wait32(ra);
        tmem_ld_wait32(rb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ci = 8 * j + 2 * e;
            const float a = __uint_as_float(ci < 32 ? ra[ci] : rb[ci - 32]) * inv, bb = __uint_as_float(ci + 1 < 32 ? ra[ci + 1] : rb[ci + 1 - 32]) * inv;
            split16x2<__half>(a, bb, hh[e], ll[e]);
          }
          const uint32_t off = ((uint32_t)j ^ x7) << 4;
          st_shared_v4(wrow + off, hh[0], hh[1], hh[2], hh[3]);
          st_shared_v4(wrow + 4096u + off, ll[0], ll[1], ll[2], ll[3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + rr0, r_img = quarter * 32 + rl;          // staged row → (query, head) of the image
          if (r_img < p.rows) {
            const int qq = r_img / p.heads, hh_ = r_img - qq * p.heads;
            const uint32_t so = stg + (uint32_t)rl * 128u + (((uint32_t)piece ^ (uint32_t)(rl & 7)) << 4);
            uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
            ld_shared_v4(so, h0, h1, h2, h3);
            ld_shared_v4(so + 4096u, l0, l1, l2, l3);
            __half* o = p.out + ((int64_t)b * p.nq + qq) * ld_out + (int64_t)hh_ * p.D + c0 + piece * 8;
            *reinterpret_cast<uint4*>(o) = make_uint4(h0, h1, h2, h3);
            *reinterpret_cast<uint4*>(o + part) = make_uint4(l0, l1, l2, l3);
          }
        }
        __syncwarp();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(u_drained);
      if (warp == 2) XA_TRACE(1, 4);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// weight folding (once per model): Wg[(h, c), c'] = scale · Σ_d Wk[h·64 + d, c] · Wq[h·64 + d, c'],  g[(h, c)] = scale · Σ_d bq[h·64 + d] · Wk[h·64 + d, c]
__global__ void xattn_fold_qk_kernel(const float* __restrict__ wq, const float* __restrict__ wk, const float* __restrict__ bq, __half* __restrict__ wg,
                                     float* __restrict__ g, int D, int heads, float scale) {
  const int hc = blockIdx.x, h = hc / D, c = hc - h * D;
  for (int cp = threadIdx.x; cp <= D; cp += blockDim.x) {
    double acc = 0.0;
    for (int d = 0; d < 64; ++d) {
      const double k = (double)wk[(int64_t)(h * 64 + d) * D + c];
      acc += k * (cp < D ? (double)wq[(int64_t)(h * 64 + d) * D + cp] : (double)bq[h * 64 + d]);
    }
    if (cp < D) wg[(int64_t)hc * D + cp] = __float2half_rn((float)(acc * scale));
    else g[hc] = (float)(acc * scale);
  }
}
// Mcat[n, (h, c)] = Σ_d Wo[n, h·64 + d] · Wv[h·64 + d, c]  (fp32, split afterwards);  bo2[n] = bo[n] + Σ_j Wo[n, j] · bv[j]
__global__ void xattn_fold_vo_kernel(const float* __restrict__ wv, const float* __restrict__ wo, const float* __restrict__ bv, const float* __restrict__ bo,
                                     float* __restrict__ mcat, float* __restrict__ bo2, int D, int heads) {
  const int n = blockIdx.x;
  for (int hc = threadIdx.x; hc <= heads * D; hc += blockDim.x) {
    double acc = 0.0;
    if (hc < heads * D) {
      const int h = hc / D, c = hc - h * D;
      for (int d = 0; d < 64; ++d) acc += (double)wo[(int64_t)n * D + h * 64 + d] * (double)wv[(int64_t)(h * 64 + d) * D + c];
      mcat[(int64_t)n * heads * D + hc] = (float)acc;
    } else {
      for (int j = 0; j < D; ++j) acc += (double)wo[(int64_t)n * D + j] * (double)bv[j];
      bo2[n] = (float)(acc + (double)bo[n]);
    }
  }
}

}  // namespace

bool xattn_supported(int nq, int heads, int D, int hw) {
  return nq * heads <= XA_BM && hw <= XA_MAXK && D % 64 == 0 && D <= 384 && XA_U_COL + D <= 512;
}

// in_proj_weight [3D, D] / in_proj_bias [3D] / out_proj.weight [D, D] / out_proj.bias [D] of one decoder layer's multihead_attn →
// wg [heads·D, D] fp16 (+ g [heads·D]), mcat [D, heads·D] fp32 (caller splits it), bo2 [D]
int xattn_fold_weights(const float* in_proj_w, const float* in_proj_b, const float* out_w, const float* out_b, __half* wg, float* g, float* mcat,
                       float* bo2, int D, int heads, cudaStream_t s) {
  SMK_REQUIRE(D == heads * 64, "xattn_fold_weights: head dim must be 64");
  xattn_fold_qk_kernel<<<heads * D, 128, 0, s>>>(in_proj_w, in_proj_w + (int64_t)D * D, in_proj_b, wg, g, D, heads, 0.125f);
  SMK_CHECK_LAUNCH();
  xattn_fold_vo_kernel<<<D, 256, 0, s>>>(in_proj_w + (int64_t)2 * D * D, out_w, in_proj_b + 2 * D, out_b, mcat, bo2, D, heads);
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// qp: [n_img·nq, heads·D] fp16 (row (b, q), column (h, c)); tok: [n_img·t_rows_per_img, D] fp16 final-LN tokens, image b's keys at rows
// b·t_rows_per_img + t_row0 .. + hw; out: [n_img·nq, 2·heads·D] fp16 split [hi | lo]
int xattn_tc(const __half* qp, const __half* tok, int t_rows_per_img, int t_row0, __half* out, int n_img, int nq, int heads, int D, int hw,
             cudaStream_t s) {
  SMK_REQUIRE(xattn_supported(nq, heads, D, hw), "xattn_tc: nq=%d heads=%d D=%d hw=%d not supported", nq, heads, D, hw);
  SMK_REQUIRE(n_img >= 0 && ((uintptr_t)out % 16) == 0, "xattn_tc: bad arguments");
  if (n_img == 0) return SMK_OK;
  const int rows = nq * heads, nk_pad = (hw + 15) / 16 * 16, n_chunks = D / XA_KC;
  CUtensorMap tq, tt;
  // Q' viewed as [n_img·rows, D]: row (b, q, h)
  SMK_PROPAGATE(make_tmap_bf16_2d(&tq, qp, (uint64_t)D, (uint64_t)n_img * rows, (uint64_t)D * 2, XA_KC, XA_BM));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tt, tok, (uint64_t)D, (uint64_t)n_img * t_rows_per_img, (uint64_t)D * 2, XA_KC, (uint32_t)nk_pad));
  const int smem = n_chunks * nk_pad * 128 + XA_QSLOTS * XA_Q_BYTES + 4 * XA_STG_BYTES + 256 + 1024;
  SMK_REQUIRE(smem <= 227 * 1024, "xattn_tc: %d bytes of shared memory needed", smem);
  XAttnParams p{rows, hw, nk_pad, D, n_chunks, rows, t_rows_per_img, t_row0, heads, nq, n_img, out, 1.4426950408889634f};
  const int grid = n_img < device_sm_count() ? n_img : device_sm_count();
  const bool exact13 = nk_pad == 13 * 16;
  static DeviceOnce attr_a, attr_b;
  if (exact13) {
    if (attr_a.first()) SMK_CHECK_CUDA(cudaFuncSetAttribute(xattn_tc_kernel<13, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  } else {
    if (attr_b.first()) SMK_CHECK_CUDA(cudaFuncSetAttribute(xattn_tc_kernel<14, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  {
    // credited: the memory K/V projection this kernel makes unnecessary + the attention products (per image and layer)
    ProfScope prof(PROF_ATTENTION_TC, (double)n_img * (2.0 * hw * D * 2.0 * D + 4.0 * nq * hw * D), s,
                   (double)n_img * 2.0 * (2.0 * XA_BM * nk_pad * D));
    if (exact13) SMK_CHECK_CUDA(launch_pdl(xattn_tc_kernel<13, true>, dim3(grid), dim3(XA_THREADS), (size_t)smem, s, tq, tt, p));
    else SMK_CHECK_CUDA(launch_pdl(xattn_tc_kernel<14, false>, dim3(grid), dim3(XA_THREADS), (size_t)smem, s, tq, tt, p));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace smk

extern "C" int smk_xattn_tc(const void* qp, const void* tok, int t_rows_per_img, int t_row0, void* out, int n_img, int nq, int heads, int D,
                            int hw, void* stream) {
  SMK_REQUIRE(qp && tok && out, "smk_xattn_tc: null pointer");
  return smk::xattn_tc((const __half*)qp, (const __half*)tok, t_rows_per_img, t_row0, (__half*)out, n_img, nq, heads, D, hw, (cudaStream_t)stream);
}
extern "C" int smk_debug_xattn_trace(long long* buf) {   // tuning aid: 2 x 8 x 8 clock64 stamps of CTA 0; nullptr switches it off
  SMK_CHECK_CUDA(cudaMemcpyToSymbol(smk::g_xattn_trace, &buf, sizeof(buf)));
  return SMK_OK;
}
extern "C" int smk_xattn_fold_weights(const float* in_proj_w, const float* in_proj_b, const float* out_w, const float* out_b, void* wg, float* g,
                                      float* mcat, float* bo2, int D, int heads, void* stream) {
  SMK_REQUIRE(in_proj_w && in_proj_b && out_w && out_b && wg && g && mcat && bo2, "smk_xattn_fold_weights: null pointer");
  return smk::xattn_fold_weights(in_proj_w, in_proj_b, out_w, out_b, (__half*)wg, g, mcat, bo2, D, heads, (cudaStream_t)stream);
}
