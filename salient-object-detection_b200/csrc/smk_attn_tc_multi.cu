// Fused multi-head softmax attention on tcgen05 for sequences LONGER than one key tile (vision_transformer.py:110-130 at 384 x 384:
// 577 tokens; ViT-S/8: 785 tokens), head dim 64 — the multi-key-tile sibling of smk_attn_tc.cu (same warp roles, same single-pass
// softmax, same P-over-S / TS-form P·V / ones-column row sums), with an online rescale across key tiles:
//
//   item  = (image, head, group of 256 query rows): two 128-row query tiles ping-pong on the two 256-column TMEM slots
//   step  = (item, key tile j): K_j / V_j (176 keys, double-buffered by TMA) → S_t = Q_t·K_j^T (UMMA 128 x 176 x 64) into slot
//           columns [0, 176) → one thread per query row: online softmax over eleven 16-key units against the running integer maximum
//           carried ACROSS key tiles → if the maximum moved since the previous tile, the row's O accumulator (slot columns
//           [176, 256): 64 values + the ones-column row sum) is scaled by the exact power of two 2^(M_prev − M_new) in TMEM
//           (tcgen05.ld / st; skipped warp-wide when no row's maximum changed, which is the common case after the first tile)
//           → P (16-bit pairs) over S → O_t (+)= P·V_j.  After the last key tile: O / rowsum → staging → TMA store.
//   The last key tile is the window [Lk − 176, Lk): every tile has exactly 11 units (one software-pipelined basic block, see
//   smk_attn_tc.cu) and the keys the previous tile already covered are masked to −inf.
// With O parked at columns [176, 256) a slot has room for 176-key S tiles only (the single-tile kernel overlays O on S's tail, which
// is impossible once O must survive the next tile's S): 577 keys = 4 tiles, 785 = 5.
#include <type_traits>

#include "smk_tc.cuh"

namespace smk {

using namespace tc;

namespace {

constexpr int AM_BM = 128, AM_DH = 64, AM_KT = 176, AM_UNITS = AM_KT / 16, AM_THREADS = 384, AM_REGS_CTRL = 48, AM_REGS_SOFTMAX = 224;
constexpr int AM_Q_BYTES = 2 * AM_BM * 128, AM_KV_BYTES = AM_KT * 128;       // 32 KB, 22 KB
constexpr int AM_OFF_K = AM_Q_BYTES, AM_OFF_V = AM_OFF_K + 2 * AM_KV_BYTES, AM_OFF_ONES = AM_OFF_V + 2 * AM_KV_BYTES;
constexpr int AM_OFF_STG = AM_OFF_ONES + AM_KV_BYTES, AM_STG_BYTES = 8192, AM_OFF_BAR = AM_OFF_STG + 8 * AM_STG_BYTES;
constexpr int AM_SMEM = AM_OFF_BAR + 256 + 1024;
constexpr int AM_O_COL = AM_KT, AM_PV_N = AM_DH + 16;
static_assert(AM_O_COL + AM_PV_N <= 256 && AM_SMEM <= 227 * 1024 && AM_KV_BYTES % 1024 == 0, "TMEM slot / shared memory budget");

struct AttnMultiParams {
  int Lq, Lk;                   // query rows / keys per image (>= AM_KT keys)
  int n_qgroups, n_ktiles;      // ceil(Lq / 256), ceil(Lk / 176)
  int q_rows, kv_rows, kv_row0; // image b: queries from row b*q_rows, keys / values from row b*kv_rows + kv_row0
  int heads, n_items, out_mode; // items = images x heads x query groups; out_mode 0 16-bit, 1 fp32, 2 [hi | hi | lo], 3 [hi | lo]
  float scale_log2e;
  int rev;
  int out_bf16;                 // fp16 operands only: the [hi | hi | lo] / [hi | lo] output parts are bf16 (consumer: a bf16 split GEMM)
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
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <bool kF16>
__global__ void __launch_bounds__(AM_THREADS, 1)
attn_tc_multi_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmO, const AttnMultiParams p) {
  using T16 = typename std::conditional<kF16, __half, __nv_bfloat16>::type;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AM_OFF_BAR);
  uint64_t *q_full = bars, *q_empty = bars + 1, *k_full = bars + 2, *k_empty = bars + 4, *v_full = bars + 6, *v_empty = bars + 8;   // k / v: [2]
  uint64_t *s_full = bars + 10, *p_full = bars + 12, *o_full = bars + 14;                                                          // [2] each
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nkt = p.n_ktiles, n_steps = n_my_items * nkt;
  auto decode = [&](int it, int& b, int& h, int& qg) {
    const int item_i = blockIdx.x + it * gridDim.x, item = p.rev ? p.n_items - 1 - item_i : item_i;
    qg = item % p.n_qgroups;
    const int bh = item / p.n_qgroups;
    b = bh / p.heads;
    h = bh - b * p.heads;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  constexpr uint32_t kOnes = kF16 ? 0x3C003C00u : 0x3F803F80u;
  for (int i = threadIdx.x; i < AM_KV_BYTES / 16; i += AM_THREADS)
    *reinterpret_cast<uint4*>(smem + AM_OFF_ONES + i * 16) = make_uint4(kOnes, kOnes, kOnes, kOnes);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_ptr);
  pdl_wait();
  pdl_trigger();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AM_REGS_CTRL));
    if (warp == 0) {
      // ===== TMA producer (whole warp, one elected lane issues) =====
      for (int n = 0; n < n_steps; ++n) {
        const int it = n / nkt, j = n - it * nkt, s = n & 1;
        const uint32_t spar = (uint32_t)((n >> 1) & 1);
        int b, h, qg;
        decode(it, b, h, qg);
        const int key0 = (j == nkt - 1) ? p.Lk - AM_KT : j * AM_KT;
        const int kv_row = b * p.kv_rows + p.kv_row0 + key0;
        if (j == 0) {
          mbar_wait(q_empty, (it & 1) ^ 1);                       // every S product of the previous item has retired
          mbar_arrive_expect_tx_w(q_full, AM_Q_BYTES);
          tma_load_2d_w(sQ, &tmQ, q_full, h * AM_DH, b * p.q_rows + qg * 2 * AM_BM);
        }
        mbar_wait(&k_empty[s], spar ^ 1);
        mbar_arrive_expect_tx_w(&k_full[s], AM_KV_BYTES);
        tma_load_2d_w(smem + AM_OFF_K + s * AM_KV_BYTES, &tmK, &k_full[s], h * AM_DH, kv_row);
        mbar_wait(&v_empty[s], spar ^ 1);
        mbar_arrive_expect_tx_w(&v_full[s], AM_KV_BYTES);
        tma_load_2d_w(smem + AM_OFF_V + s * AM_KV_BYTES, &tmV, &v_full[s], h * AM_DH, kv_row);
      }
    } else if (warp == 1) {
      // ===== MMA issuer (whole warp convergent; one elected lane issues each tcgen05 instruction) =====
      const uint32_t idesc_s = idesc_16_f32<kF16>(AM_BM, AM_KT, 0, 0);
      const uint32_t idesc_o = idesc_16_f32<kF16>(AM_BM, AM_PV_N, 0, 1);
      auto issue_s = [&](int t, int n) {          // S_t = Q_t · K_j^T into slot t, columns [0, 176)
        const uint64_t qd = smem_desc_k_sw128(smem_u32(sQ + t * AM_BM * 128));
        const uint64_t kd = smem_desc_k_sw128(smem_u32(smem + AM_OFF_K + (n & 1) * AM_KV_BYTES));
        const uint32_t d_tmem = tmem_base + (uint32_t)(t * 256);
#pragma unroll
        for (int k = 0; k < AM_DH / 16; ++k) umma_bf16_ss_w(d_tmem, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
        tc_commit_w(&s_full[t]);
      };
      auto issue_pv = [&](int t, int n, bool acc) {   // O_t (+)= P_t · V_j; MN block 1 of the B operand = the ones tile (row sums)
        const uint32_t sv = smem_u32(smem + AM_OFF_V + (n & 1) * AM_KV_BYTES);
        const uint64_t vd0 = smem_desc_mn_sw128(sv, (uint32_t)(AM_OFF_ONES - AM_OFF_V - (n & 1) * AM_KV_BYTES));
        const uint32_t slot = tmem_base + (uint32_t)(t * 256);
#pragma unroll
        for (int u = 0; u < AM_UNITS; ++u)
          umma_bf16_ts_w(slot + AM_O_COL, slot + (uint32_t)(u * 8), vd0 + (uint64_t)(u * 128), idesc_o, (acc || u != 0) ? 1u : 0u);
        tc_commit_w(&o_full[t]);
      };
      if (n_steps > 0) {
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after_sync();
        issue_s(0, 0);
        issue_s(1, 0);
        tc_commit_w(&k_empty[0]);
        if (nkt == 1) tc_commit_w(q_empty);
      }
      for (int n = 0; n < n_steps; ++n) {
        const int it = n / nkt, j = n - it * nkt;
        const bool has_next = n + 1 < n_steps, next_new_item = j + 1 == nkt;
        mbar_wait(&v_full[n & 1], (uint32_t)((n >> 1) & 1));
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], (uint32_t)(n & 1));        // P_t is in TMEM, S_t fully read, O_t rescaled
          tc_fence_after_sync();
          issue_pv(t, n, j > 0);
          if (t == 1) tc_commit_w(&v_empty[n & 1]);
          if (has_next) {
            if (t == 0) {
              if (next_new_item) mbar_wait(q_full, (uint32_t)((it + 1) & 1));
              mbar_wait(&k_full[(n + 1) & 1], (uint32_t)(((n + 1) >> 1) & 1));
              tc_fence_after_sync();
            }
            issue_s(t, n + 1);          // in order behind P·V_t of this step: P_t is consumed before S overwrites it; O is not touched
            if (t == 1) {
              tc_commit_w(&k_empty[(n + 1) & 1]);
              if ((n + 1) % nkt == nkt - 1) tc_commit_w(q_empty);
            }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AM_REGS_SOFTMAX));
    // ===== softmax + epilogue: group t = (warp - 4) / 4 owns query tile t; TMEM lane quarter = warp % 4 =====
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 256);
    const uint32_t x7s = (uint32_t)(lane & 7) << 4;
    uint8_t* stg_ptr = smem + AM_OFF_STG + (warp - 4) * AM_STG_BYTES;
    const uint32_t stg = smem_u32(stg_ptr);
    const float sc = p.scale_log2e;
    float M = -1.0e30f;                                 // running integer-valued maximum of s*scale*log2e, carried across key tiles
    for (int n = 0; n < n_steps; ++n) {
      const int it = n / nkt, j = n - it * nkt;
      int b, h, qg;
      decode(it, b, h, qg);
      const int row0 = qg * 2 * AM_BM + t * AM_BM + quarter * 32;     // first query row (within the image) of this warp
      const bool active = row0 < p.Lq;
      if (j == 0) M = -1.0e30f;
      // columns of this tile that an earlier tile already covered (the last tile is the window [Lk - 176, Lk))
      const int dup = (j == nkt - 1) ? j * AM_KT - (p.Lk - AM_KT) : 0;
      mbar_wait(&s_full[t], (uint32_t)(n & 1));
      tc_fence_after_sync();
      const float M_prev = M;
      uint32_t pk[AM_UNITS][8];
      float mrec[AM_UNITS];
      if (active) {
        uint32_t va[16], vb[16];
        auto unit = [&](uint32_t (&v)[16], uint32_t (&pu)[8], float& mr, int c0) {
          if (c0 < dup) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (c0 + jj < dup) v[jj] = 0xff800000u;
          }
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int jj = 0; jj < 16; jj += 4) {
            m0 = max3(m0, __uint_as_float(v[jj]), __uint_as_float(v[jj + 1]));
            m1 = max3(m1, __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
          }
          const float Mn = fmaxf(M, ceilf(fmaxf(m0, m1) * sc));
          M = Mn;
          mr = Mn;
          float e[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) e[jj] = ex2_approx(fmaf(__uint_as_float(v[jj]), sc, -Mn));
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) pu[jj] = Pack16<T16>::pack(e[2 * jj], e[2 * jj + 1]);
        };
        tmem_ld_32x16(taddr, va);
#pragma unroll
        for (int u = 0; u < AM_UNITS; ++u) {
          if (u & 1) {
            tmem_ld_wait16(vb);
            if (u + 1 < AM_UNITS) tmem_ld_32x16(taddr + (uint32_t)((u + 1) * 16), va);
            unit(vb, pk[u], mrec[u], u * 16);
          } else {
            tmem_ld_wait16(va);
            if (u + 1 < AM_UNITS) tmem_ld_32x16(taddr + (uint32_t)((u + 1) * 16), vb);
            unit(va, pk[u], mrec[u], u * 16);
          }
        }
      }
      if (j > 0) {
        // the previous key tile's P·V has retired: bring O (and its ones-column row sum) to the new maximum where it moved
        mbar_wait(&o_full[t], (uint32_t)((n - 1) & 1));
        tc_fence_after_sync();
        if (active && __any_sync(0xffffffffu, M != M_prev)) {
          const float fo = ex2_approx(M_prev - M);      // exact power of two (1 for rows whose maximum did not move)
          uint32_t oa[32], oc[16];
#pragma unroll
          for (int c = 0; c < 2; ++c) {                 // 32 columns at a time: the P registers of this tile stay live
            tmem_ld_32x32(taddr + AM_O_COL + (uint32_t)(c * 32), oa);
            tmem_ld_wait32(oa);
#pragma unroll
            for (int i = 0; i < 32; ++i) oa[i] = __float_as_uint(__uint_as_float(oa[i]) * fo);
            tmem_st_32x32(taddr + AM_O_COL + (uint32_t)(c * 32), oa);
          }
          tmem_ld_32x16(taddr + AM_O_COL + 64u, oc);
          tmem_ld_wait16(oc);
#pragma unroll
          for (int i = 0; i < 16; ++i) oc[i] = __float_as_uint(__uint_as_float(oc[i]) * fo);
          tmem_st_32x16(taddr + AM_O_COL + 64u, oc);
        }
      }
      if (active) {
        // bring every unit to the tile's final maximum (exact power-of-two factors) and write P over S
#pragma unroll
        for (int u = 0; u < AM_UNITS; ++u) {
          const float fc = ex2_approx(mrec[u] - M);
          if constexpr (kF16) {
            const __half2 f2 = __float2half2_rn(fc);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              __half2 x = *reinterpret_cast<__half2*>(&pk[u][jj]);
              x = __hmul2(x, f2);
              pk[u][jj] = *reinterpret_cast<uint32_t*>(&x);
            }
          } else {
            const __nv_bfloat162 f2 = __float2bfloat162_rn(fc);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&pk[u][jj]);
              x = __hmul2(x, f2);
              pk[u][jj] = *reinterpret_cast<uint32_t*>(&x);
            }
          }
          tmem_st_32x8(taddr + (uint32_t)(u * 8), pk[u]);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (j == nkt - 1) {
        // epilogue: O / rowsum → staging → TMA store to out[b, row0 .. row0+32, h*64 .. h*64+64]
        mbar_wait(&o_full[t], (uint32_t)(n & 1));
        tc_fence_after_sync();
        uint32_t oa[32], ob[32], osum = 0x3F800000u;
        if (active) {
          tmem_ld_32x32(taddr + AM_O_COL, oa);
          tmem_ld_32x32(taddr + AM_O_COL + 32u, ob);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(osum) : "r"(taddr + AM_O_COL + 64u) : "memory");   // row sum of P
          tmem_ld_wait32(oa);
          tmem_ld_wait32(ob);
          asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(osum)::"memory");
        }
        if (active) {
          if (lane == 0) bulk_wait_read<0>();            // the previous item's output tile has left the staging buffer
          __syncwarp();
          const float inv = 1.0f / __uint_as_float(osum);
          auto f = [&](uint32_t u) { return __uint_as_float(u) * inv; };
          const uint32_t srow = stg + lane * 128;
          if (p.out_mode == 1) {
            // two 32-column fp32 boxes of 32 rows x 128 B each, 128-byte swizzle
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              st_shared_v4(srow + (((uint32_t)j << 4) ^ x7s), __float_as_uint(f(oa[4 * j])), __float_as_uint(f(oa[4 * j + 1])),
                           __float_as_uint(f(oa[4 * j + 2])), __float_as_uint(f(oa[4 * j + 3])));
              st_shared_v4(srow + 4096 + (((uint32_t)j << 4) ^ x7s), __float_as_uint(f(ob[4 * j])), __float_as_uint(f(ob[4 * j + 1])),
                           __float_as_uint(f(ob[4 * j + 2])), __float_as_uint(f(ob[4 * j + 3])));
            }
          } else if (kF16 && p.out_mode == 4) {
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
          } else if (p.out_mode >= 2) {
            // split output: hi and lo 64-column boxes (32 rows x 128 B each); mode 2: hi is stored twice ([hi | hi | lo]), mode 3: [hi | lo]
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t hh[4], ll[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c0 = 8 * j + 2 * e;
                const float a = c0 < 32 ? f(oa[c0]) : f(ob[c0 - 32]), bb = c0 + 1 < 32 ? f(oa[c0 + 1]) : f(ob[c0 + 1 - 32]);
                if (kF16 && p.out_bf16) split16x2<__nv_bfloat16>(a, bb, hh[e], ll[e]);
                else split16x2<T16>(a, bb, hh[e], ll[e]);
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
            tma_store_3d(&tmO, stg_ptr, h * AM_DH, row0, b);
            if (p.out_mode == 1) tma_store_3d(&tmO, stg_ptr + 4096, h * AM_DH + 32, row0, b);
            if (p.out_mode == 2) {
              const int Dm = p.heads * AM_DH;
              tma_store_3d(&tmO, stg_ptr, Dm + h * AM_DH, row0, b);
              tma_store_3d(&tmO, stg_ptr + 4096, 2 * Dm + h * AM_DH, row0, b);
            }
            if (p.out_mode >= 3) tma_store_3d(&tmO, stg_ptr + 4096, p.heads * AM_DH + h * AM_DH, row0, b);
            bulk_commit();
          }
        }
      }
    }
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// q / k / v: 16-bit matrices (bf16, or fp16 when f16), head h at columns [h*64, h*64+64) of each pointer; image b's queries start at row
// b*q_rows, its keys / values at row b*kv_rows + kv_row0.  Lq >= 1, Lk >= 176 (shorter sequences: smk_attn_tc.cu).  out [B*Lq, ldo]:
// out_mode 0 16-bit (operand type), 1 fp32, 2 [hi | hi | lo], 3 [hi | lo] in the operand type (out_bf16: bf16 parts next to fp16 operands),
// 4 [hi fp16 | e4m3 correction operands] (fp16 only).
int attention_tc_multi(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t q_total_rows,
                       int64_t kv_total_rows, int q_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_mode, int B, int Lq, int Lk,
                       int heads, float scale, int f16, cudaStream_t s, int out_bf16) {
  const int D = heads * AM_DH;
  SMK_REQUIRE(!out_bf16 || (f16 && (out_mode == 2 || out_mode == 3)), "attention_tc_multi: bf16 output parts go with fp16 operands and a split output mode");
  SMK_REQUIRE(Lk >= AM_KT && Lq >= 1 && B >= 1 && heads >= 1, "attention_tc_multi: Lq=%d / Lk=%d not supported (Lk >= 176)", Lq, Lk);
  SMK_REQUIRE(out_mode >= 0 && out_mode <= 4 && (out_mode != 4 || f16) && (out_mode != 2 || ldo >= 3 * (int64_t)D) && (out_mode < 3 || ldo >= 2 * (int64_t)D),
              "attention_tc_multi: bad output mode / ldo");
  const int esz = out_mode == 1 ? 4 : 2;
  SMK_REQUIRE((ldo * esz) % 16 == 0 && ((uintptr_t)out % 16) == 0, "attention_tc_multi: output must be 16-byte aligned");
  const int n_qgroups = (Lq + 2 * AM_BM - 1) / (2 * AM_BM), n_ktiles = (Lk + AM_KT - 1) / AM_KT;
  SMK_REQUIRE((int64_t)B * heads * n_qgroups < (1 << 30), "attention_tc_multi: too many items");
  CUtensorMap tq, tk, tv, to;
  SMK_PROPAGATE(make_tmap_bf16_2d(&tq, q, (uint64_t)D, (uint64_t)q_total_rows, (uint64_t)ldq * 2, AM_DH, 2 * AM_BM));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tk, k, (uint64_t)D, (uint64_t)kv_total_rows, (uint64_t)ldk * 2, AM_DH, AM_KT));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tv, v, (uint64_t)D, (uint64_t)kv_total_rows, (uint64_t)ldv * 2, AM_DH, AM_KT));
  {
    const int parts = out_mode == 2 ? 3 : (out_mode >= 3 ? 2 : 1);
    const uint64_t dims[3] = {(uint64_t)(parts * D), (uint64_t)Lq, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ldo * esz, (uint64_t)q_rows * ldo * esz};
    const uint32_t box[3] = {out_mode == 1 ? 32u : 64u, 32u, 1u};
    SMK_PROPAGATE(make_tmap_nd(&to, esz, out, 3, dims, strides, box, 128));
  }
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_multi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_SMEM));
    SMK_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_multi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_SMEM));
  }
  const int n_items = B * heads * n_qgroups;
  AttnMultiParams p{Lq, Lk, n_qgroups, n_ktiles, q_rows, kv_rows, kv_row0, heads, n_items, out_mode, scale * 1.4426950408889634f, traverse_dir(), out_bf16};
  const int grid = n_items < device_sm_count() ? n_items : device_sm_count();
  {
    ProfScope prof(PROF_ATTENTION_TC, 4.0 * Lq * Lk * AM_DH * heads * B, s);
    if (f16) SMK_CHECK_CUDA(launch_pdl(attn_tc_multi_kernel<true>, dim3(grid), dim3(AM_THREADS), (size_t)AM_SMEM, s, tq, tk, tv, to, p));
    else SMK_CHECK_CUDA(launch_pdl(attn_tc_multi_kernel<false>, dim3(grid), dim3(AM_THREADS), (size_t)AM_SMEM, s, tq, tk, tv, to, p));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace smk

extern "C" int smk_attention_tc_multi(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t q_total_rows,
                                      int64_t kv_total_rows, int q_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_mode, int B,
                                      int Lq, int Lk, int heads, float scale, int f16, void* stream) {
  SMK_REQUIRE(q && k && v && out, "smk_attention_tc_multi: null pointer");
  return smk::attention_tc_multi(q, ldq, k, ldk, v, ldv, q_total_rows, kv_total_rows, q_rows, kv_rows, kv_row0, out, ldo, out_mode & 7, B, Lq, Lk, heads,
                                 scale, f16, (cudaStream_t)stream, (out_mode >> 3) & 1);
}
