// Fused multi-head softmax attention on tcgen05 (encoder: vision_transformer.py:110-130; decoder self / cross
// attention: transformer_decoder.py:271-291 via nn.MultiheadAttention), head dim 64, up to 256 queries and 256 keys
// per (image, head).
//
// Persistent and warp-specialised: one CTA per SM loops over (image, head) items.
//   warp 0      TMA producer: Q (one or two 128-row tiles), K and V boxes straight out of the projection outputs
//               (strided tensor maps, no permute copy :113-118).  Q/K and V have separate full/empty barriers: Q and K are
//               released as soon as both S = Q·K^T products retire, so the next item's loads overlap this item's softmax.
//   warp 1      tcgen05.mma issuer.  S_t = Q_t·K^T (UMMA 128 x nk x 64, SS) into TMEM slot t; O_t = P_t·V with P_t read
//               from TMEM (TS form: no shared-memory round trip for P) and V consumed as an MN-major operand from its
//               TMA tile.  Issue order ping-pongs the two slots (PV0, next S0, PV1, next S1) so that one tile's MMAs and
//               epilogue run under the other tile's softmax.
//   warps 4-11  two softmax groups (one per query tile, 4 warps = the 4 TMEM lane quarters): one thread per query row.
//               The TMEM read port (64 B/clk/SM) bounds this kernel, so S is read from TMEM exactly ONCE: an online
//               softmax over 16-key units keeps every unit's P in registers as packed bf16 pairs (8 registers per unit),
//               exponentiated against the running row maximum rounded UP to an integer (in the log2 domain), so that the
//               final correction of an earlier unit is a multiplication by an exact power of two (HMUL2.BF16, no second
//               rounding).  The row sum is not accumulated by the softmax warps at all: the P·V MMA runs with N = 80, its B operand's
//               second MN block being a shared-memory tile of ones, so O[:, 64] = Σ_k P[:, k] of the very bf16 values the MMA
//               multiplies (pass 3.95 k → 2.7 k cycles per tile: the 16-way add trees were the longest dependency chains of
//               the unit).  P is then written back over S with tcgen05.st;
//               after O_t lands: O / rowsum → shared-memory staging → TMA tile store through a 3-D {column,
//               row-in-image, image} tensor map, which clips the rows beyond the image's last query.
//               (A two-pass version — row max, then exp2 — read S twice: 480 instead of 272 columns per row and item,
//               8.3 k cycles per item of which 7.7 k were the TMEM port; splitting the columns over warp pairs, 16 softmax
//               warps, did not help for the same reason: profiles/r01_attention_tmem.md.)
// TMEM slot t (256 columns): S in [0, nk), P packed in [0, nk/2) (written after the whole row is in registers), O in
// [128, 192) (S columns consumed before the first P·V MMA is issued).
#include <type_traits>

#include "smk_tc.cuh"

namespace smk {

using namespace tc;

constexpr int AT_BM = 128, AT_DH = 64, AT_TMEM_COLS = 512, AT_MAXK = 256;
// three warpgroups: {TMA producer, MMA issuer, 2 idle warps} and two softmax groups.  The kernel is compiled for 168 registers
// (three warps per scheduler); setmaxnreg then moves registers from the first warpgroup to the softmax warpgroups, whose
// single-pass softmax keeps a whole row of P in registers.
// (budget: 12 warps x 168 = 4 x 48 + 8 x 224 + slack; setmaxnreg.inc blocks for ever if the CTA pool cannot cover it)
constexpr int AT_THREADS = 384, AT_REGS_CTRL = 48, AT_REGS_SOFTMAX = 224;
constexpr int AT_Q_BYTES = 2 * AT_BM * 128;          // two query tiles
constexpr int AT_KV_BYTES = AT_MAXK * 128;           // up to 256 keys x 64 dims bf16
constexpr int AT_STG_BYTES = 8192;                   // per softmax warp: 32 rows x 256 B (64 fp32) output staging
// the "ones" tile: second 64-column MN block of the P·V B operand (only its first 16 columns are used, N = 80): column 64 of O
// becomes the row sum of the bf16 P that the MMA actually multiplies — the softmax warps no longer add the exponentials up
constexpr int AT_OFF_K = AT_Q_BYTES, AT_OFF_V = AT_OFF_K + AT_KV_BYTES, AT_OFF_ONES = AT_OFF_V + AT_KV_BYTES, AT_OFF_STG = AT_OFF_ONES + AT_KV_BYTES;
constexpr int AT_PV_N = AT_DH + 16;
constexpr int AT_OFF_BAR = AT_OFF_STG + 8 * AT_STG_BYTES;
constexpr int AT_SMEM = AT_OFF_BAR + 256 /*barriers*/ + 1024 /*align*/;
constexpr int AT_O_COL = 128;
#ifndef SMK_ATTN_POLY_MASK
#define SMK_ATTN_POLY_MASK 0x0         // bit j: element j of every 16-key unit uses ex2_poly (experiment, off: see ex2_poly)
#endif
constexpr unsigned kPolyMask = SMK_ATTN_POLY_MASK;                        // O accumulator columns within a slot
static_assert(AT_O_COL + AT_PV_N <= 256, "O accumulator must stay inside its TMEM slot");
static_assert(AT_SMEM <= 227 * 1024, "attention shared memory budget");

struct AttnTcParams {
  int Lq, Lk, nk_pad;           // valid query rows / keys per image, keys padded to a multiple of 16
  int n_qtiles;                 // 1 (Lq <= 128) or 2
  int q_rows, kv_rows, kv_row0; // image b: queries start at row b*q_rows, keys/values at row b*kv_rows + kv_row0
  int heads, n_items;           // items = images x heads
  int out_f32;
  float scale_log2e;
  int rev;                      // walk the items from the last one (smk_kernels.h g_traverse_rev)
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

// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax of 2^f
// (max relative error 7.5e-5, far below the bf16 rounding of P), exponent added with one shift-add.  The MUFU (16 ex2 per clock
// and SM) is the busiest unit of the softmax pass, but the pass has no issue slots to spare either: with a quarter of the
// exponentials on this path (SMK_ATTN_POLY_MASK=0x8888) the pass takes 4.5 k instead of 3.95 k cycles (56.3 vs 53.7 us).  Off.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.0551716648f, 0.2426111251f);   // minimax fit of 2^f on [-0.5, 0.5] (relative error 7.5e-5, tests/test_host_cpu.py)
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// optional phase trace of CTA 0 (tuning scripts only): [item < 16][warp 12][event 8] clock64 stamps
__device__ long long* g_attn_trace = nullptr;
#define AT_TRACE(ev)                                                                                                     \
  do {                                                                                                                   \
    if (g_attn_trace && blockIdx.x == 0 && lane == 0 && it < 16) g_attn_trace[(it * 12 + warp) * 8 + (ev)] = clock64(); \
  } while (0)

// kMaxUnits: upper bound of the 16-key units per row; kExact: the row has exactly kMaxUnits units (13 = 193..208 keys: the
// encoder's 197 tokens), so the unrolled softmax has no run-time guards and is one basic block that ptxas can software-pipeline
// across units (warps issue in order: with a branch per unit the max → exp2 → sum chains of successive units ran back to back).
// kF16: fp16 operands (Q, K, V, P) instead of bf16 — 11 significant bits: the fp16s mode's encoder attention (single pass; its
// contribution to the mask-logit error is 1.7e-3 of the 2e-2 budget, scripts/precision_emulation.py)
template <int kMaxUnits, bool kExact, bool kF16>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const __grid_constant__ CUtensorMap tmO, const AttnTcParams p) {
  using T16 = typename std::conditional<kF16, __half, __nv_bfloat16>::type;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_OFF_K;
  uint8_t* sV = smem + AT_OFF_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t *qk_full = bars, *qk_empty = bars + 1, *v_full = bars + 2, *v_empty = bars + 3;
  uint64_t *s_full = bars + 4, *p_full = bars + 6, *o_full = bars + 8, *o_drained = bars + 10;   // [2] each
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = p.n_qtiles;
  const int n_my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(qk_full, 1); mbar_init(qk_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 4); mbar_init(&o_full[t], 1); mbar_init(&o_drained[t], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, AT_TMEM_COLS);
  constexpr uint32_t kOnes = kF16 ? 0x3C003C00u : 0x3F803F80u;          // 1.0 everywhere (uniform, so the swizzle does not matter)
  for (int i = threadIdx.x; i < AT_KV_BYTES / 16; i += AT_THREADS)
    *reinterpret_cast<uint4*>(smem + AT_OFF_ONES + i * 16) = make_uint4(kOnes, kOnes, kOnes, kOnes);
  fence_proxy_async();               // generic-proxy writes → visible to the tensor core's async-proxy reads
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // through a warp reduction (REDUX → uniform register) so that ptxas can keep the MMA operands in uniform registers
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_ptr);
  pdl_wait();      // PDL: the set-up above overlaps the previous kernel's tail
  pdl_trigger();

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AT_REGS_CTRL));
   if (warp == 0) {
    // ===== TMA producer (whole warp, one elected lane issues) =====
    const uint32_t q_bytes = (uint32_t)nt * AT_BM * 128u, kv_bytes = (uint32_t)p.nk_pad * 128u;
    for (int it = 0; it < n_my_items; ++it) {
      const int item_i = blockIdx.x + it * gridDim.x, item = p.rev ? p.n_items - 1 - item_i : item_i;
      const int b = item / p.heads, h = item % p.heads;
      const int kv_row = b * p.kv_rows + p.kv_row0;
      mbar_wait(qk_empty, (it & 1) ^ 1);
      AT_TRACE(0);
      mbar_arrive_expect_tx_w(qk_full, q_bytes + kv_bytes);
      tma_load_2d_w(sQ, &tmQ, qk_full, h * AT_DH, b * p.q_rows);
      tma_load_2d_w(sK, &tmK, qk_full, h * AT_DH, kv_row);
      mbar_wait(v_empty, (it & 1) ^ 1);
      AT_TRACE(1);
      mbar_arrive_expect_tx_w(v_full, kv_bytes);
      tma_load_2d_w(sV, &tmV, v_full, h * AT_DH, kv_row);
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp stays convergent; one elected lane issues each tcgen05 instruction) =====
    const uint32_t idesc_s = idesc_16_f32<kF16>(AT_BM, p.nk_pad, 0, 0);
    const uint32_t idesc_o = idesc_16_f32<kF16>(AT_BM, AT_PV_N, 0, 1);
    const uint64_t kd = smem_desc_k_sw128(smem_u32(sK));
    const uint64_t vd0 = smem_desc_mn_sw128(smem_u32(sV), AT_OFF_ONES - AT_OFF_V);   // MN block 1 (columns 64..79) = the ones tile
    const int n_ksteps = p.nk_pad / 16;
    auto issue_s = [&](int t) {               // S_t = Q_t · K^T
      const uint64_t qd = smem_desc_k_sw128(smem_u32(sQ + t * AT_BM * 128));
      const uint32_t d_tmem = tmem_base + (uint32_t)(t * 256);
#pragma unroll
      for (int k = 0; k < AT_DH / 16; ++k) umma_bf16_ss_w(d_tmem, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
      tc_commit_w(&s_full[t]);
    };
    auto issue_pv = [&](int t) {              // O_t = P_t · V, P_t from TMEM (8 packed columns per 16 keys)
      const uint32_t slot = tmem_base + (uint32_t)(t * 256);
#pragma unroll 4
      for (int j = 0; j < n_ksteps; ++j)      // V: +2048 B per 16 keys → +128 in the descriptor's (addr >> 4) field
        umma_bf16_ts_w(slot + AT_O_COL, slot + (uint32_t)(j * 8), vd0 + (uint64_t)(j * 128), idesc_o, j != 0);
      tc_commit_w(&o_full[t]);
    };
    if (n_my_items > 0) {
      const int it = 0;
      mbar_wait(qk_full, 0);
      tc_fence_after_sync();
      AT_TRACE(0);
      for (int t = 0; t < nt; ++t) issue_s(t);
      tc_commit_w(qk_empty);                  // Q and K may be overwritten once both S products retire
    }
    for (int it = 0; it < n_my_items; ++it) {
      const uint32_t par = it & 1;
      const bool has_next = it + 1 < n_my_items;
      mbar_wait(v_full, par);
      AT_TRACE(1);
      for (int t = 0; t < nt; ++t) {
        mbar_wait(&p_full[t], par);           // P_t is in TMEM and S_t has been fully read
        tc_fence_after_sync();
        AT_TRACE(2 + 2 * t);
        issue_pv(t);
        if (t == nt - 1) tc_commit_w(v_empty);
        if (has_next) {
          if (t == 0) mbar_wait(qk_full, par ^ 1);
          mbar_wait(&o_drained[t], par);      // slot t: this item's O has been read out
          tc_fence_after_sync();
          AT_TRACE(3 + 2 * t);
          issue_s(t);
          if (t == nt - 1) tc_commit_w(qk_empty);
        }
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AT_REGS_SOFTMAX));
    // ===== softmax + epilogue: group t = (warp - 4) / 4 owns query tile t; TMEM lane quarter = warp % 4 =====
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    if (t < nt) {
      const int row0 = t * AT_BM + quarter * 32;     // first query row (within the image) of this warp
      const bool active = row0 < p.Lq;               // warps past the last query only keep the barrier protocol going
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 256);
      const uint32_t x7s = (uint32_t)(lane & 7) << 4;   // 128-byte swizzle of the staging tile: 16-byte chunk index ^= row & 7
      uint8_t* stg_ptr = smem + AT_OFF_STG + (warp - 4) * AT_STG_BYTES;
      const uint32_t stg = smem_u32(stg_ptr);
      const int nu = kExact ? kMaxUnits : (p.nk_pad >> 4);   // 16-key units of a row (<= kMaxUnits)
      const float sc = p.scale_log2e;
      for (int it = 0; it < n_my_items; ++it) {
        const int item_i = blockIdx.x + it * gridDim.x, item = p.rev ? p.n_items - 1 - item_i : item_i;
        const uint32_t par = it & 1;
        const int b = item / p.heads, h = item % p.heads;
        AT_TRACE(0);
        mbar_wait(&s_full[t], par);
        tc_fence_after_sync();
        AT_TRACE(1);
        if (active) {
          // ---- single pass over S: online softmax per 16-key unit, P kept in registers (bf16 pairs) ----
          uint32_t pk[kMaxUnits][8];
          float mrec[kMaxUnits];                     // the running maximum each unit was exponentiated against
          uint32_t va[16], vb[16];
          float M = -1.0e30f;                        // running integer-valued max of s*scale*log2e
          auto unit = [&](uint32_t (&v)[16], uint32_t (&pu)[8], float& mr, int c0, bool last) {
            if (last) {                               // keys beyond Lk (padding of the last unit) drop out as -inf
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
            for (int j = 0; j < 16; ++j) {
              const float xj = fmaf(__uint_as_float(v[j]), sc, -Mn);
              e[j] = (kPolyMask >> j) & 1 ? ex2_poly(xj) : ex2_approx(xj);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) pu[j] = Pack16<T16>::pack(e[2 * j], e[2 * j + 1]);
          };
          tmem_ld_32x16(taddr, va);
#pragma unroll
          for (int u = 0; u < kMaxUnits; ++u) {
            if (u < nu) {
              if (u & 1) {
                tmem_ld_wait16(vb);
                if (u + 1 < nu) tmem_ld_32x16(taddr + (uint32_t)((u + 1) * 16), va);
                unit(vb, pk[u], mrec[u], u * 16, u == nu - 1);
              } else {
                tmem_ld_wait16(va);
                if (u + 1 < nu) tmem_ld_32x16(taddr + (uint32_t)((u + 1) * 16), vb);
                unit(va, pk[u], mrec[u], u * 16, u == nu - 1);
              }
            }
          }
          AT_TRACE(2);
          // ---- bring every unit to the final maximum (exact: power-of-two factors) and write P over S ----
#pragma unroll
          for (int u = 0; u < kMaxUnits; ++u) {
            if (u < nu) {
              const float fc = ex2_approx(mrec[u] - M);       // an exact power of two <= 1 (fp16: flushes towards 0 below 2^-24)
              if constexpr (kF16) {
                const __half2 f2 = __float2half2_rn(fc);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  __half2 x = *reinterpret_cast<__half2*>(&pk[u][j]);
                  x = __hmul2(x, f2);
                  pk[u][j] = *reinterpret_cast<uint32_t*>(&x);
                }
              } else {
                const __nv_bfloat162 f2 = __float2bfloat162_rn(fc);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&pk[u][j]);
                  x = __hmul2(x, f2);
                  pk[u][j] = *reinterpret_cast<uint32_t*>(&x);
                }
              }
              tmem_st_32x8(taddr + (uint32_t)(u * 8), pk[u]);
            }
          }
          tmem_st_wait();
        }
        tc_fence_before_sync();         // our TMEM reads of S / writes of P are ordered before the PV MMAs
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        AT_TRACE(3);
        // epilogue: O / rowsum → staging → TMA store to out[b, row0 .. row0+32, h*64 .. h*64+64]
        mbar_wait(&o_full[t], par);
        tc_fence_after_sync();
        AT_TRACE(4);
        uint32_t oa[32], ob[32], osum = 0x3F800000u;
        if (active) {
          tmem_ld_32x32(taddr + AT_O_COL, oa);
          tmem_ld_32x32(taddr + AT_O_COL + 32u, ob);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(osum) : "r"(taddr + AT_O_COL + 64u) : "memory");   // row sum of P
          tmem_ld_wait32(oa);
          tmem_ld_wait32(ob);
          asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(osum)::"memory");
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_drained[t]);       // slot t may receive the next item's S
        AT_TRACE(5);
        if (active) {
          if (lane == 0) bulk_wait_read<0>();            // the previous item's output tile has left the staging buffer
          __syncwarp();
          const float inv = 1.0f / __uint_as_float(osum);
          auto f = [&](uint32_t u) { return __uint_as_float(u) * inv; };
          const uint32_t srow = stg + lane * 128;
          if (p.out_f32 == 1) {
            // two 32-column fp32 boxes of 32 rows x 128 B each, 128-byte swizzle
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              st_shared_v4(srow + (((uint32_t)j << 4) ^ x7s), __float_as_uint(f(oa[4 * j])), __float_as_uint(f(oa[4 * j + 1])),
                           __float_as_uint(f(oa[4 * j + 2])), __float_as_uint(f(oa[4 * j + 3])));
              st_shared_v4(srow + 4096 + (((uint32_t)j << 4) ^ x7s), __float_as_uint(f(ob[4 * j])), __float_as_uint(f(ob[4 * j + 1])),
                           __float_as_uint(f(ob[4 * j + 2])), __float_as_uint(f(ob[4 * j + 3])));
            }
          } else if (kF16 && p.out_f32 == 4) {
            // [hi | e4m3 correction operands] (A operand of a terms_q8 GEMM): the second 128-byte box holds, per 32 columns, e4m3(hi) (32 B)
            // and e4m3(lo·2^11) (32 B)
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) {
              uint2 h[4];
              uint32_t f8[4], s8[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c0 = 16 * j2 + 4 * e;
                auto g = [&](int c) { return c < 32 ? f(oa[c < 32 ? c : 0]) : f(ob[c >= 32 ? c - 32 : 0]); };
                split_q8x4<false>(g(c0), g(c0 + 1), g(c0 + 2), g(c0 + 3), h[e], f8[e], s8[e]);
              }
              st_shared_v4(srow + (((uint32_t)(2 * j2) << 4) ^ x7s), h[0].x, h[0].y, h[1].x, h[1].y);
              st_shared_v4(srow + (((uint32_t)(2 * j2 + 1) << 4) ^ x7s), h[2].x, h[2].y, h[3].x, h[3].y);
              const uint32_t base = (uint32_t)((j2 >> 1) * 4 + (j2 & 1));
              st_shared_v4(srow + 4096 + ((base << 4) ^ x7s), f8[0], f8[1], f8[2], f8[3]);
              st_shared_v4(srow + 4096 + (((base + 2) << 4) ^ x7s), s8[0], s8[1], s8[2], s8[3]);
            }
          } else if (p.out_f32 >= 2) {
            // split output: hi and lo 64-column boxes (32 rows x 128 B each); mode 2: hi is stored twice ([hi | hi | lo]), mode 3: [hi | lo]
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t hh[4], ll[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c0 = 8 * j + 2 * e;
                const float a = c0 < 32 ? f(oa[c0]) : f(ob[c0 - 32]), bb = c0 + 1 < 32 ? f(oa[c0 + 1]) : f(ob[c0 + 1 - 32]);
                split16x2<T16>(a, bb, hh[e], ll[e]);
              }
              const uint32_t off = ((uint32_t)j << 4) ^ x7s;
              st_shared_v4(srow + off, hh[0], hh[1], hh[2], hh[3]);
              st_shared_v4(srow + 4096 + off, ll[0], ll[1], ll[2], ll[3]);
            }
          } else {
            // one 64-column bf16 box of 32 rows x 128 B
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              st_shared_v4(srow + (((uint32_t)j << 4) ^ x7s), Pack16<T16>::pack(f(oa[8 * j]), f(oa[8 * j + 1])), Pack16<T16>::pack(f(oa[8 * j + 2]), f(oa[8 * j + 3])),
                           Pack16<T16>::pack(f(oa[8 * j + 4]), f(oa[8 * j + 5])), Pack16<T16>::pack(f(oa[8 * j + 6]), f(oa[8 * j + 7])));
              st_shared_v4(srow + (((uint32_t)(j + 4) << 4) ^ x7s), Pack16<T16>::pack(f(ob[8 * j]), f(ob[8 * j + 1])), Pack16<T16>::pack(f(ob[8 * j + 2]), f(ob[8 * j + 3])),
                           Pack16<T16>::pack(f(ob[8 * j + 4]), f(ob[8 * j + 5])), Pack16<T16>::pack(f(ob[8 * j + 6]), f(ob[8 * j + 7])));
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmO, stg_ptr, h * AT_DH, row0, b);
            if (p.out_f32 == 1) tma_store_3d(&tmO, stg_ptr + 4096, h * AT_DH + 32, row0, b);
            if (p.out_f32 == 2) {
              const int Dm = p.heads * AT_DH;
              tma_store_3d(&tmO, stg_ptr, Dm + h * AT_DH, row0, b);
              tma_store_3d(&tmO, stg_ptr + 4096, 2 * Dm + h * AT_DH, row0, b);
            }
            if (p.out_f32 >= 3) tma_store_3d(&tmO, stg_ptr + 4096, p.heads * AT_DH + h * AT_DH, row0, b);
            bulk_commit();
          }
        }
        AT_TRACE(6);
      }
      if (active && lane == 0) bulk_wait<0>();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, AT_TMEM_COLS);
}

template <bool kF16>
static int attention_tc_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t kv_total_rows,
                               int kv_rows, int kv_row0, void* out, int64_t ldo, int out_f32, int B, int Lq, int Lk, int heads, float scale,
                               cudaStream_t s) {
  const int D = heads * AT_DH;
  SMK_REQUIRE(Lk >= 1 && Lk <= AT_MAXK && Lq >= 1 && Lq <= 2 * AT_BM, "attention_tc: Lq=%d / Lk=%d not supported (1..256)", Lq, Lk);
  SMK_REQUIRE(B >= 1 && heads >= 1 && (int64_t)B * heads < (1 << 30), "attention_tc: bad batch/heads");
  SMK_REQUIRE(out_f32 >= 0 && out_f32 <= 4 && (out_f32 != 4 || kF16) && (out_f32 != 2 || ldo >= 3 * (int64_t)D) && (out_f32 < 3 || ldo >= 2 * (int64_t)D),
              "attention_tc: bad output mode / ldo");
  const int esz = out_f32 == 1 ? 4 : 2;
  SMK_REQUIRE((ldo * esz) % 16 == 0 && ((uintptr_t)out % 16) == 0, "attention_tc: output must be 16-byte aligned");
  const int nk_pad = (Lk + 15) / 16 * 16;
  const int nt = Lq > AT_BM ? 2 : 1;
  CUtensorMap tq, tk, tv, to;
  SMK_PROPAGATE(make_tmap_bf16_2d(&tq, q, (uint64_t)D, (uint64_t)B * Lq, (uint64_t)ldq * 2, AT_DH, (uint32_t)(nt * AT_BM)));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tk, k, (uint64_t)D, (uint64_t)kv_total_rows, (uint64_t)ldk * 2, AT_DH, (uint32_t)nk_pad));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tv, v, (uint64_t)D, (uint64_t)kv_total_rows, (uint64_t)ldv * 2, AT_DH, (uint32_t)nk_pad));
  {
    // {column, query row within the image, image}: rows >= Lq of a 32-row output box are clipped by the TMA unit
    const int parts = out_f32 == 2 ? 3 : (out_f32 >= 3 ? 2 : 1);
    const uint64_t dims[3] = {(uint64_t)(parts * D), (uint64_t)Lq, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ldo * esz, (uint64_t)Lq * ldo * esz};
    const uint32_t box[3] = {out_f32 == 1 ? 32u : 64u, 32u, 1u};
    SMK_PROPAGATE(make_tmap_nd(&to, esz, out, 3, dims, strides, box, 128));
  }
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<13, true, kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    SMK_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<16, false, kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
  }
  const int n_items = B * heads;
  AttnTcParams p{Lq, Lk, nk_pad, nt, Lq, kv_rows, kv_row0, heads, n_items, out_f32, scale * 1.4426950408889634f, traverse_dir()};
  const int grid = n_items < device_sm_count() ? n_items : device_sm_count();
  {
    ProfScope prof(PROF_ATTENTION_TC, 4.0 * Lq * Lk * AT_DH * heads * B, s);
    if (nk_pad == 13 * 16) SMK_CHECK_CUDA(launch_pdl(attn_tc_kernel<13, true, kF16>, dim3(grid), dim3(AT_THREADS), (size_t)AT_SMEM, s, tq, tk, tv, to, p));
    else SMK_CHECK_CUDA(launch_pdl(attn_tc_kernel<16, false, kF16>, dim3(grid), dim3(AT_THREADS), (size_t)AT_SMEM, s, tq, tk, tv, to, p));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

int attention_tc_general(const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v, int64_t ldv,
                         int64_t kv_total_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_f32, int B, int Lq, int Lk,
                         int heads, float scale, cudaStream_t s) {
  SMK_REQUIRE(out_f32 != 3, "attention_tc: the [hi | lo] output is the fp16 form's");
  return attention_tc_launch<false>(q, ldq, k, ldk, v, ldv, kv_total_rows, kv_rows, kv_row0, out, ldo, out_f32, B, Lq, Lk, heads, scale, s);
}

// encoder form — qkv: [B*N, 3*D] bf16 (q | k | v, head h in columns [h*64, h*64+64) of each third); out: [B*N, D] bf16
int attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int heads, float scale, cudaStream_t s) {
  const int D = heads * AT_DH;
  return attention_tc_general(qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, (int64_t)B * N, N, 0, out, D, 0, B, N, N, heads, scale, s);
}

// fp16 encoder form — qkv: [B*N, 3*D] fp16; out: [B*N, ldo] fp16, out_mode 0 plain (ldo >= D) or 3 = [hi | lo] split (ldo >= 2D)
int attention_tc_f16(const __half* qkv, __half* out, int64_t ldo, int out_mode, int B, int N, int heads, float scale, cudaStream_t s) {
  const int D = heads * AT_DH;
  SMK_REQUIRE(out_mode == 0 || out_mode == 3 || out_mode == 4, "attention_tc_f16: output mode 0 (fp16), 3 ([hi | lo] fp16) or 4 ([hi | q8])");
  return attention_tc_launch<true>(qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, (int64_t)B * N, N, 0, out, ldo, out_mode, B, N, N, heads, scale, s);
}

}  // namespace smk

extern "C" int smk_debug_attn_trace(long long* buf) {   // tuning aid: nullptr switches the trace off
  SMK_CHECK_CUDA(cudaMemcpyToSymbol(smk::g_attn_trace, &buf, sizeof(buf)));
  return SMK_OK;
}

extern "C" int smk_attention_tc(const void* qkv, void* out, int B, int N, int heads, float scale, void* stream) {
  SMK_REQUIRE(qkv && out, "smk_attention_tc: null pointer");
  return smk::attention_tc((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, B, N, heads, scale, (cudaStream_t)stream);
}

extern "C" int smk_attention_tc_f16(const void* qkv, void* out, int64_t ldo, int out_mode, int B, int N, int heads, float scale, void* stream) {
  SMK_REQUIRE(qkv && out, "smk_attention_tc_f16: null pointer");
  return smk::attention_tc_f16((const __half*)qkv, (__half*)out, ldo, out_mode, B, N, heads, scale, (cudaStream_t)stream);
}

extern "C" int smk_attention_tc_general(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                        int64_t kv_total_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_f32, int B, int Lq,
                                        int Lk, int heads, float scale, void* stream) {
  SMK_REQUIRE(q && k && v && out, "smk_attention_tc_general: null pointer");
  return smk::attention_tc_general((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, kv_total_rows,
                                   kv_rows, kv_row0, out, ldo, out_f32, B, Lq, Lk, heads, scale, (cudaStream_t)stream);
}
