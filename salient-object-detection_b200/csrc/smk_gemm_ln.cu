// Residual GEMM with the following LayerNorm fused into its epilogue (bf16 tensor-core mode, encoder blocks):
//     X[M,N] += A[M,K] · W[N,K]^T + bias          (attn.proj / mlp.fc2 + residual, vision_transformer.py:131,168-170)
//     Xn[M,N] = LayerNorm(X; gamma, beta, eps)     (the next norm2 / norm1, :165,169) as the bf16 operand of the next GEMM
// for N = 384 = one full row per tile.  A stand-alone LayerNorm launch reads the 77 MB fp32 residual stream again right after
// the GEMM wrote it (24 of them per forward pass, ~22 us each); here the row is normalised while it is still on the SM.
//
// One CTA owns 128 rows x all 384 columns: the fp32 accumulator fills 384 of the 512 TMEM columns (two UMMA 128x192x16 per
// k-step), so there is no second accumulator buffer — MMAs and epilogue of a CTA alternate, and the operand ring keeps
// prefetching the next tile's first k-blocks under the epilogue.
//   warp 0  TMA producer (A 128x64 + W 2 x 192x64 boxes per stage, 128-byte swizzle)
//   warp 1  tcgen05.mma issuer
//   warps 2-9  epilogue, TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 (192 columns = 6 chunks of 32):
//     pass 1  v = acc + bias + x_old; x_old comes as a TMA tile (32 x 32 fp32, swizzled) so that every thread reads its own row
//             conflict-free; v goes back to TMEM (tcgen05.st), to X through the same staging buffer + TMA store, and into sum(v)
//     pass 2  sum((v - mean)^2) from the TMEM copy — the same two-pass statistics as layernorm_kernel
//     pass 3  (v - mean) * rstd * gamma + beta → bf16 → staging → TMA store to Xn
//   The two warps that share a row (column halves) exchange their partial sums through shared memory and a 64-thread named
//   barrier.
#include <stdlib.h>

#include "smk_tc.cuh"
#include "smk_kernels.h"

namespace smk {

using namespace tc;

namespace {

constexpr int GL_BM = 128, GL_BK = 64, GL_N = 384, GL_NH = 192, GL_STAGES = 2, GL_EPI_WARPS = 8, GL_THREADS = 64 + 32 * GL_EPI_WARPS;
constexpr int GL_A_BYTES = GL_BM * GL_BK * 2, GL_BH_BYTES = GL_NH * GL_BK * 2, GL_STAGE_BYTES = GL_A_BYTES + 2 * GL_BH_BYTES;
constexpr int GL_STG_PER_WARP = 2 * 4096;                    // two 32 x 32 fp32 tiles (x_old in, v / bf16 out), alternating per chunk
constexpr int GL_OFF_STG = GL_STAGES * GL_STAGE_BYTES, GL_OFF_EXCH = GL_OFF_STG + GL_EPI_WARPS * GL_STG_PER_WARP;
constexpr int GL_OFF_BAR = GL_OFF_EXCH + 2 * 2 * GL_BM * 4;  // exchange: [sum | sumsq][column half][row]
constexpr int GL_SMEM = GL_OFF_BAR + 512 + 1024;
constexpr int GL_TMEM_COLS = 512;
static_assert(GL_SMEM <= 227 * 1024, "shared memory budget");

struct GemmLnParams {
  int M, K;
  const float *bias, *gamma, *beta;
  float eps;
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__global__ void __launch_bounds__(GL_THREADS, 1)
gemm_ln_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmXn, const GemmLnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + GL_OFF_STG;
  float* exch = reinterpret_cast<float*>(smem + GL_OFF_EXCH);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GL_OFF_BAR);
  uint64_t* empty_bar = full_bar + GL_STAGES;
  uint64_t* tmem_full = empty_bar + GL_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint64_t* x_bar = tmem_empty + 1;                           // [epilogue warp][buffer]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(x_bar + 2 * GL_EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (p.M + GL_BM - 1) / GL_BM, k_blocks = p.K / GL_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmXn);
    for (int i = 0; i < GL_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, GL_EPI_WARPS);
    for (int i = 0; i < 2 * GL_EPI_WARPS; ++i) mbar_init(&x_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, GL_TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_ptr);
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int m_blk = blockIdx.x; m_blk < m_blocks; m_blk += gridDim.x) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * GL_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], GL_STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GL_BK, m_blk * GL_BM);
          tma_load_2d(sa + GL_A_BYTES, &tmW, &full_bar[stage], kb * GL_BK, 0);
          tma_load_2d(sa + GL_A_BYTES + GL_BH_BYTES, &tmW, &full_bar[stage], kb * GL_BK, GL_NH);
          if (++stage == GL_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(GL_BM, GL_NH, 0, 0);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int m_blk = blockIdx.x; m_blk < m_blocks; m_blk += gridDim.x) {
        mbar_wait(tmem_empty, acc_phase ^ 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * GL_STAGE_BYTES);
          const uint64_t a_desc = smem_desc_k_sw128(sa), b0_desc = smem_desc_k_sw128(sa + GL_A_BYTES),
                         b1_desc = smem_desc_k_sw128(sa + GL_A_BYTES + GL_BH_BYTES);
#pragma unroll
          for (int k = 0; k < GL_BK / 16; ++k) {
            umma_bf16_ss(tmem_base, a_desc + (uint64_t)(2 * k), b0_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            umma_bf16_ss(tmem_base + GL_NH, a_desc + (uint64_t)(2 * k), b1_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == GL_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tmem_full);
        acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2, quarter = warp & 3, ch = ew >> 2;
    constexpr int kChunks = GL_NH / 32;
    uint8_t* stg = staging + ew * GL_STG_PER_WARP;
    const uint32_t stg_u32 = smem_u32(stg);
    uint64_t* xb = x_bar + 2 * ew;
    const int row_in_tile = quarter * 32 + lane;
    float* ex_sum = exch + ch * GL_BM + row_in_tile;                    // mine
    const float* ex_sum_o = exch + (ch ^ 1) * GL_BM + row_in_tile;     // the other column half's
    float* ex_sq = exch + 2 * GL_BM + ch * GL_BM + row_in_tile;
    const float* ex_sq_o = exch + 2 * GL_BM + (ch ^ 1) * GL_BM + row_in_tile;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ch * GL_NH);
    uint32_t acc_phase = 0, x_phase = 0;      // x_phase: bit b = parity of x_old buffer b
    auto r4 = [](float f) { return __float_as_uint(f); };
    for (int m_blk = blockIdx.x; m_blk < m_blocks; m_blk += gridDim.x) {
      const int row0 = m_blk * GL_BM + quarter * 32;
      // x_old tile of chunk 0 can be fetched under the MMAs (buffer 0 is free: its last store was drained below)
      if (lane == 0) {
        mbar_arrive_expect_tx(&xb[0], 4096);
        tma_load_2d(stg, &tmX, &xb[0], ch * GL_NH, row0);
      }
      mbar_wait(tmem_full, acc_phase);
      tc_fence_after_sync();
      // ---- pass 1: v = acc + bias + x_old → TMEM, X; row sum ----
      float s1 = 0.f;
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        const int n0 = ch * GL_NH + ci * 32, buf = ci & 1;
        uint8_t* sb = stg + buf * 4096;
        if (ci + 1 < kChunks && lane == 0) {          // prefetch the next chunk's x_old into the other buffer
          bulk_wait_read<0>();                         // its previous TMA store has finished reading the buffer
          mbar_arrive_expect_tx(&xb[buf ^ 1], 4096);
          tma_load_2d(stg + (buf ^ 1) * 4096, &tmX, &xb[buf ^ 1], n0 + 32, row0);
        }
        uint32_t r[32];
        tmem_ld_32x32(lane_taddr + (uint32_t)(ci * 32), r);
        mbar_wait(&xb[buf], (x_phase >> buf) & 1u);
        x_phase ^= 1u << buf;
        const uint32_t srow = stg_u32 + buf * 4096 + lane * 128;
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t a, b, c, d;
          ld_shared_v4(srow + ((j ^ (lane & 7)) << 4), a, b, c, d);
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
          v[4 * j] = __uint_as_float(a) + bv.x; v[4 * j + 1] = __uint_as_float(b) + bv.y;
          v[4 * j + 2] = __uint_as_float(c) + bv.z; v[4 * j + 3] = __uint_as_float(d) + bv.w;
        }
        tmem_ld_wait32(r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += __uint_as_float(r[j]);
          s1 += v[j];
          r[j] = __float_as_uint(v[j]);
        }
        {
          uint32_t lo16[16], hi16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { lo16[j] = r[j]; hi16[j] = r[16 + j]; }
          tmem_st_32x16(lane_taddr + (uint32_t)(ci * 32), lo16);
          tmem_st_32x16(lane_taddr + (uint32_t)(ci * 32 + 16), hi16);
        }
        __syncwarp();                                  // every lane has read its x_old row before the buffer is overwritten
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(srow + ((j ^ (lane & 7)) << 4), r4(v[4 * j]), r4(v[4 * j + 1]), r4(v[4 * j + 2]), r4(v[4 * j + 3]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmX, sb, n0, row0);
          bulk_commit();
        }
      }
      tmem_st_wait();
      // ---- mean: exchange the partial sums of the two column halves ----
      *ex_sum = s1;
      named_bar_sync(1 + quarter, 64);
      const float mean = (s1 + *ex_sum_o) * (1.0f / (float)GL_N);
      // ---- pass 2: sum of squared deviations from the TMEM copy ----
      float s2 = 0.f;
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        uint32_t r[32];
        tmem_ld_32x32(lane_taddr + (uint32_t)(ci * 32), r);
        tmem_ld_wait32(r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(r[j]) - mean;
          s2 = fmaf(d, d, s2);
        }
      }
      *ex_sq = s2;
      named_bar_sync(1 + quarter, 64);
      const float rstd = 1.0f / sqrtf((s2 + *ex_sq_o) * (1.0f / (float)GL_N) + p.eps);
      // ---- pass 3: normalise → bf16 → Xn ----
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        const int n0 = ch * GL_NH + ci * 32, buf = ci & 1;
        uint8_t* sb = stg + buf * 4096;
        uint32_t r[32];
        tmem_ld_32x32(lane_taddr + (uint32_t)(ci * 32), r);
        tmem_ld_wait32(r);
        if (ci == kChunks - 1) {                       // the accumulator columns are free again
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);
        }
        float o[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + n0) + j), b = __ldg(reinterpret_cast<const float4*>(p.beta + n0) + j);
          o[4 * j] = (__uint_as_float(r[4 * j]) - mean) * rstd * g.x + b.x;
          o[4 * j + 1] = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * g.y + b.y;
          o[4 * j + 2] = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * g.z + b.z;
          o[4 * j + 3] = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * g.w + b.w;
        }
        if (lane == 0) bulk_wait_read<1>();            // the store that last used this buffer has read it
        __syncwarp();
        const uint32_t srow = stg_u32 + buf * 4096 + lane * 64;      // 32 rows x 64 B, 64-byte swizzle
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(srow + ((j ^ ((lane >> 1) & 3)) << 4), pack_bf16x2(o[8 * j], o[8 * j + 1]), pack_bf16x2(o[8 * j + 2], o[8 * j + 3]),
                       pack_bf16x2(o[8 * j + 4], o[8 * j + 5]), pack_bf16x2(o[8 * j + 6], o[8 * j + 7]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmXn, sb, n0, row0);
          bulk_commit();
        }
      }
      if (lane == 0) bulk_wait_read<0>();              // both buffers drained: the next tile's x_old load may overwrite buffer 0
      __syncwarp();
      acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, GL_TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------------------------------------
// Cluster variant: two CTAs of a cluster split the 384 columns (128 x 192 tiles each).  192 accumulator columns leave room for
// two TMEM buffers, so the epilogue of tile i runs under the MMAs of tile i+1 and the HBM traffic of the epilogue is spread over
// the whole kernel instead of arriving in bursts.  LayerNorm statistics: every CTA forms (mean, M2) of its 192 columns with the
// same two-pass scheme as above and the two halves are combined exactly (Chan et al.: mean = (m_a + m_b) / 2,
// M2 = M2_a + M2_b + (m_a - m_b)^2 · 96) after ONE exchange through distributed shared memory per tile: each thread stores its
// row's (mean, M2) into the peer CTA's buffer (st.shared::cluster) and arrives on the peer's mbarrier with release.cluster.
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int G2_NH = 192, G2_STAGES = 2, G2_A_BYTES = GL_BM * GL_BK * 2, G2_B_BYTES = G2_NH * GL_BK * 2, G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_STG_PER_WARP = 4 * 4096;      // three 32 x 32 fp32 x_old tiles (refilled for the next tile right after pass 1) + one output tile
constexpr int G2_OFF_STG = G2_STAGES * G2_STAGE_BYTES, G2_OFF_EXCH = G2_OFF_STG + GL_EPI_WARPS * G2_STG_PER_WARP;
constexpr int G2_OFF_XEXCH = G2_OFF_EXCH + 2 * 2 * GL_BM * 4;            // in-CTA exchange: [sum | sumsq][column half][row]
constexpr int G2_OFF_BAR = G2_OFF_XEXCH + 2 * 2 * GL_BM * 4;             // cross-CTA exchange: [tile parity][mean | M2][row]
constexpr int G2_SMEM = G2_OFF_BAR + 512 + 1024;
static_assert(G2_SMEM <= 227 * 1024, "shared memory budget");

// remote store that completes 4 transaction bytes on an mbarrier of the destination CTA: the consumer's plain mbarrier wait
// then sees the data, without the cluster-scope fences (CCTL.IVALL + MEMBAR, ~20 % of the kernel's stall samples) that a
// st.shared::cluster + mbarrier.arrive.release.cluster / try_wait.acquire.cluster pair costs
__device__ __forceinline__ void st_async_cluster_f32(uint32_t cluster_addr, float v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr), "r"(__float_as_uint(v)),
               "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 4000000000LL) {
      printf("smk: cluster mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

__global__ void __launch_bounds__(GL_THREADS, 1)
gemm_ln2_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ CUtensorMap tmXn, const GemmLnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + G2_OFF_STG;
  float* exch = reinterpret_cast<float*>(smem + G2_OFF_EXCH);
  float* xexch = reinterpret_cast<float*>(smem + G2_OFF_XEXCH);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_OFF_BAR);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full = empty_bar + G2_STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint64_t* x_bar = tmem_empty + 2;                 // [epilogue warp][buffer]
  // [tile parity][quarter]: the peer CTA's 32 rows of (mean, M2) have landed.  Two barriers per quarter, alternating with the
  // tile parity: this CTA sends before it waits, so the peer may already be sending tile t+1 while a warp here still waits for
  // tile t — with one barrier that could complete two phases under a waiter; tile t+2 cannot be sent before tile t was consumed.
  uint64_t* cx_bar = x_bar + 3 * GL_EPI_WARPS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(cx_bar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_cl = gridDim.x >> 1, cl = blockIdx.x >> 1;
  const int m_blocks = (p.M + GL_BM - 1) / GL_BM, k_blocks = p.K / GL_BK;
  const int col0 = (int)rank * G2_NH;               // this CTA's columns [col0, col0 + 192)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmXn);
    for (int i = 0; i < G2_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], GL_EPI_WARPS); }
    for (int i = 0; i < 3 * GL_EPI_WARPS; ++i) mbar_init(&x_bar[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&cx_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, GL_TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync_all();                               // the peer's barriers are initialised before any remote arrive
  tc_fence_after_sync();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_ptr);
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int m_blk = cl; m_blk < m_blocks; m_blk += n_cl) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G2_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], G2_STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GL_BK, m_blk * GL_BM);
          tma_load_2d(sa + G2_A_BYTES, &tmW, &full_bar[stage], kb * GL_BK, col0);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(GL_BM, G2_NH, 0, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int m_blk = cl; m_blk < m_blocks; m_blk += n_cl) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * G2_NH);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint64_t a_desc = smem_desc_k_sw128(sa), b_desc = smem_desc_k_sw128(sa + G2_A_BYTES);
#pragma unroll
          for (int k = 0; k < GL_BK / 16; ++k) umma_bf16_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          tc_commit(&empty_bar[stage]);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int ew = warp - 2, quarter = warp & 3, ch = ew >> 2;
    constexpr int kCols = G2_NH / 2, kChunks = kCols / 32;               // 96 columns = 3 chunks per warp
    uint8_t* stg = staging + ew * G2_STG_PER_WARP;
    const uint32_t stg_u32 = smem_u32(stg);
    uint64_t* xb = x_bar + 3 * ew;
    uint8_t* out_buf = stg + 3 * 4096;
    const uint32_t out_u32 = stg_u32 + 3 * 4096;
    const int row_in_tile = quarter * 32 + lane;
    float* ex_sum = exch + ch * GL_BM + row_in_tile;
    const float* ex_sum_o = exch + (ch ^ 1) * GL_BM + row_in_tile;
    float* ex_sq = exch + 2 * GL_BM + ch * GL_BM + row_in_tile;
    const float* ex_sq_o = exch + 2 * GL_BM + (ch ^ 1) * GL_BM + row_in_tile;
    const uint32_t peer_xexch = mapa_shared(smem_u32(xexch), rank ^ 1u), peer_cx_bar = mapa_shared(smem_u32(&cx_bar[quarter]), rank ^ 1u);   // + 32 B for parity 1
    int acc = 0;
    uint32_t acc_phase = 0, x_phase = 0, cx_phase = 0, tile_par = 0;
    auto r4 = [](float f) { return __float_as_uint(f); };
    // x_old tiles of a whole m-block (3 chunks) are fetched as soon as the previous tile's stores have drained the buffers
    auto fetch_x = [&](int m_blk) {
      if (lane == 0) {
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) {
          mbar_arrive_expect_tx(&xb[ci], 4096);
          tma_load_2d(stg + ci * 4096, &tmX, &xb[ci], col0 + ch * kCols + ci * 32, m_blk * GL_BM + quarter * 32);
        }
      }
    };
    if (cl < m_blocks) fetch_x(cl);
    for (int m_blk = cl; m_blk < m_blocks; m_blk += n_cl) {
      const int row0 = m_blk * GL_BM + quarter * 32, ncol0 = col0 + ch * kCols;
      const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * G2_NH + ch * kCols);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      // ---- pass 1: v = acc + bias + x_old → TMEM, X; row sum ----
      float s1 = 0.f;
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        const int n0 = ncol0 + ci * 32, buf = ci;
        uint32_t r[32];
        tmem_ld_32x32(lane_taddr + (uint32_t)(ci * 32), r);
        mbar_wait(&xb[buf], (x_phase >> buf) & 1u);
        x_phase ^= 1u << buf;
        const uint32_t srow = stg_u32 + buf * 4096 + lane * 128;
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t a, b, c, d;
          ld_shared_v4(srow + ((j ^ (lane & 7)) << 4), a, b, c, d);
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
          v[4 * j] = __uint_as_float(a) + bv.x; v[4 * j + 1] = __uint_as_float(b) + bv.y;
          v[4 * j + 2] = __uint_as_float(c) + bv.z; v[4 * j + 3] = __uint_as_float(d) + bv.w;
        }
        tmem_ld_wait32(r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += __uint_as_float(r[j]);
          s1 += v[j];
          r[j] = __float_as_uint(v[j]);
        }
        {
          uint32_t lo16[16], hi16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { lo16[j] = r[j]; hi16[j] = r[16 + j]; }
          tmem_st_32x16(lane_taddr + (uint32_t)(ci * 32), lo16);
          tmem_st_32x16(lane_taddr + (uint32_t)(ci * 32 + 16), hi16);
        }
        // v goes back over the x_old tile (each lane only touches its own row); the three X stores are issued together below:
        // a TMA store takes ~1 us to read its tile, so waiting for one per chunk serialised the whole epilogue
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(srow + ((j ^ (lane & 7)) << 4), r4(v[4 * j]), r4(v[4 * j + 1]), r4(v[4 * j + 2]), r4(v[4 * j + 3]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) tma_store_2d(&tmX, stg + ci * 4096, ncol0 + ci * 32, row0);
        bulk_commit();
      }
      tmem_st_wait();
      // ---- mean of this CTA's 192 columns ----
      *ex_sum = s1;
      named_bar_sync(1 + quarter, 64);
      const float mean_l = (s1 + *ex_sum_o) * (1.0f / (float)G2_NH);
      // ---- pass 2: M2 of this CTA's 192 columns about their own mean ----
      float s2 = 0.f;
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        uint32_t r[32];
        tmem_ld_32x32(lane_taddr + (uint32_t)(ci * 32), r);
        tmem_ld_wait32(r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(r[j]) - mean_l;
          s2 = fmaf(d, d, s2);
        }
      }
      *ex_sq = s2;
      named_bar_sync(1 + quarter, 64);
      const float m2_l = s2 + *ex_sq_o;
      // ---- one exchange with the peer CTA (the other 192 columns of the same rows), then the exact combination ----
      float* mine = xexch + tile_par * 2 * GL_BM;             // the peer wrote its (mean, M2) of this tile here
      if (ch == 0) {
        if (lane == 0) mbar_arrive_expect_tx(&cx_bar[tile_par * 4 + quarter], 32 * 2 * 4);     // what the peer sends for these 32 rows
        const uint32_t dst = peer_xexch + (uint32_t)((tile_par * 2 * GL_BM + row_in_tile) * 4), bar = peer_cx_bar + tile_par * 32u;
        st_async_cluster_f32(dst, mean_l, bar);
        st_async_cluster_f32(dst + GL_BM * 4, m2_l, bar);
      }
      mbar_wait(&cx_bar[tile_par * 4 + quarter], (cx_phase >> tile_par) & 1u);
      cx_phase ^= 1u << tile_par;
      const float mean_p = mine[row_in_tile], m2_p = mine[GL_BM + row_in_tile];
      const float mean = 0.5f * (mean_l + mean_p), dm = mean_l - mean_p;
      const float rstd = 1.0f / sqrtf((m2_l + m2_p + dm * dm * (float)(G2_NH / 2)) * (1.0f / (float)GL_N) + p.eps);
      tile_par ^= 1;
      if (lane == 0) bulk_wait_read<0>();                    // the X stores have read the x_old tiles (issued two passes ago)
      __syncwarp();
      if (m_blk + n_cl < m_blocks) fetch_x(m_blk + n_cl);    // next tile's x_old lands under pass 3 and the next tile's MMAs
      // ---- pass 3: normalise → bf16 → Xn ----
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        const int n0 = ncol0 + ci * 32;
        uint32_t r[32];
        tmem_ld_32x32(lane_taddr + (uint32_t)(ci * 32), r);
        tmem_ld_wait32(r);
        if (ci == kChunks - 1) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        float o[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + n0) + j), b = __ldg(reinterpret_cast<const float4*>(p.beta + n0) + j);
          o[4 * j] = (__uint_as_float(r[4 * j]) - mean) * rstd * g.x + b.x;
          o[4 * j + 1] = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * g.y + b.y;
          o[4 * j + 2] = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * g.z + b.z;
          o[4 * j + 3] = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * g.w + b.w;
        }
        if (lane == 0) bulk_wait_read<1>();                // the store before the previous one has read this half of the output region
        __syncwarp();
        const uint32_t srow = out_u32 + (ci & 1) * 2048 + lane * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(srow + ((j ^ ((lane >> 1) & 3)) << 4), pack_bf16x2(o[8 * j], o[8 * j + 1]), pack_bf16x2(o[8 * j + 2], o[8 * j + 3]),
                       pack_bf16x2(o[8 * j + 4], o[8 * j + 5]), pack_bf16x2(o[8 * j + 6], o[8 * j + 7]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmXn, out_buf + (ci & 1) * 2048, n0, row0);
          bulk_commit();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before_sync();
  cluster_sync_all();          // the peer may still store into this CTA's exchange buffer / arrive on its barriers until here
  if (warp == 1) tmem_dealloc(tmem_base, GL_TMEM_COLS);
}

}  // namespace


// X[M,384] += A[M,K]·W[384,K]^T + bias (fp32, in place);  Xn[M,384] = LayerNorm(X) (bf16)
int gemm_ln_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, const float* bias, float* X, const float* gamma,
               const float* beta, __nv_bfloat16* Xn, int M, int N, int K, float eps, cudaStream_t s) {
  SMK_REQUIRE(N == GL_N && K % GL_BK == 0 && K >= GL_BK, "gemm_ln: needs N == 384 and K %% 64 == 0 (N=%d K=%d)", N, K);
  SMK_REQUIRE(bias && gamma && beta && ((uintptr_t)bias % 16) == 0 && ((uintptr_t)gamma % 16) == 0 && ((uintptr_t)beta % 16) == 0,
              "gemm_ln: bias / gamma / beta must be 16-byte aligned");
  if (M == 0) return SMK_OK;
  const int g_gl_sms = device_sm_count();
  static int variant = -1;      // SMK_GEMM_LN_V = 1: one CTA per 128 x 384 tile; 2 (default): cluster of two CTAs, 128 x 192 each
  if (variant < 0) {
    const char* e = getenv("SMK_GEMM_LN_V");
    variant = e ? atoi(e) : 2;
  }
  CUtensorMap ta, tw, tx, txn;
  SMK_PROPAGATE(make_tmap_bf16_2d(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GL_BK, GL_BM));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tw, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, GL_BK, GL_NH));
  SMK_PROPAGATE(make_tmap_2d(&tx, 4, X, (uint64_t)N, (uint64_t)M, (uint64_t)N * 4, 32, 32, 128));
  SMK_PROPAGATE(make_tmap_2d(&txn, 2, Xn, (uint64_t)N, (uint64_t)M, (uint64_t)N * 2, 32, 32, 64));
  static DeviceOnce attr_set;
  if (attr_set.first()) SMK_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GL_SMEM));
  const int m_blocks = (M + GL_BM - 1) / GL_BM;
  GemmLnParams p{M, K, bias, gamma, beta, eps};
  if (variant == 2) {
    static DeviceOnce attr2_set;
    if (attr2_set.first()) SMK_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    const int clusters = m_blocks < g_gl_sms / 2 ? m_blocks : g_gl_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(GL_THREADS);
    cfg.dynamicSmemBytes = G2_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    int n_attr = 1;
    if (pdl_enabled()) {
      attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
      ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    {
      ProfScope prof(PROF_GEMM_TC, 2.0 * M * N * K, s);
      SMK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln2_tc_kernel, ta, tw, tx, txn, p));
    }
    SMK_CHECK_LAUNCH();
    return SMK_OK;
  }
  {
    ProfScope prof(PROF_GEMM_TC, 2.0 * M * N * K, s);
    SMK_CHECK_CUDA(launch_pdl(gemm_ln_tc_kernel, dim3((unsigned)(m_blocks < g_gl_sms ? m_blocks : g_gl_sms)), dim3(GL_THREADS), (size_t)GL_SMEM, s, ta,
                              tw, tx, txn, p));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace smk

extern "C" int smk_gemm_ln(const void* A, int64_t lda, const void* W, const float* bias, float* X, const float* gamma, const float* beta,
                           void* Xn, int M, int N, int K, float eps, void* stream) {
  SMK_REQUIRE(A && W && X && Xn && M >= 0, "smk_gemm_ln: bad arguments");
  return smk::gemm_ln_tc((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, bias, X, gamma, beta, (__nv_bfloat16*)Xn, M, N, K, eps,
                         (cudaStream_t)stream);
}
