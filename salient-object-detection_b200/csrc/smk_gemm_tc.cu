// bf16 tensor-core GEMM for sm_100a:  C[M,N] = A[M,K] · W[N,K]^T (+bias, GELU/ReLU, +residual)
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected thread) and
// TMEM owner, warps 2-9 = epilogue (two warpgroups, each owning half of the tile's columns).  Operands are staged
// by TMA into a ring of 128-byte-swizzled shared-memory tiles (a stage = all distinct operand tiles of a k-block); fp32 accumulators live in TMEM and are
// double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue: TMEM → registers (tcgen05.ld, one accumulator row per thread) → bias / activation → swizzled
// shared-memory staging tile (conflict-free 16-byte stores) → TMA tile store.  A per-thread row store straight to
// global memory costs 32 L1 wavefronts per instruction (every lane a different 128-byte line) and made the
// epilogue 2-5x slower than the MMAs; the TMA store writes full lines.  The fp32 residual stream is updated by a
// TMA *reduce-add* (cp.reduce.async.bulk.tensor .add, performed at L2): X += A·W^T + b without the SM ever reading X.
// With K = 384 a tile's MMAs take only 12 cycles per accumulator column, so the activation must stay near
// 10 instructions per element (gelu_fast below).
// Serves: patch-embed (vision_transformer.py:184-188), qkv / proj / fc1 / fc2 (:113,131,88-94), the
// decoder's memory K/V projection (transformer_decoder.py:283-291 via nn.MultiheadAttention in_proj).
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <type_traits>

#include "smk_tc.cuh"

namespace smk {

using namespace tc;

constexpr int TC_BM = 128, TC_BK = 64;   // epilogue warps: 8 (default) or 16 (kEW template parameter); threads = 64 + 32 * kEW
constexpr int TC_STAGING_PER_WARP = 4096;   // 32 rows x 128 B (fp32 chunk) or 2 x (32 rows x 64 B) (bf16 chunks, double-buffered)

// exact-GELU (vision_transformer.py:78 nn.GELU, erf form) as  relu(x) − 0.5·|x|·(1 − erf(|x|/√2))  with
// 1 − erf(u/√2) = 2^(u·q(u)), q a degree-4 fit (monotone beyond the fit range, so no clamp is needed):
// max abs error 1.9e-5 = 0.11 bf16 ulp of the result (the bf16-mode output is rounded to bf16 right after).
// 7 FMA/ALU-pipe instructions + one MUFU.EX2; the fp32 validation mode keeps erff().
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fabsf(x);
  float q = fmaf(u, -0.00029095853f, 0.0056083198f);
  q = fmaf(u, q, -0.04784644f);
  q = fmaf(u, q, -0.46398392f);
  q = fmaf(u, q, -1.1496887f);
  const float t = ex2_approx(fmaf(u, q, -1.0f));  // 0.5·(1 − erf(|x|/√2)): the factor 0.5 rides in the exponent
  return fmaf(-u, t, fmaxf(x, 0.f));
}

constexpr int TC_BAR_BYTES = 512;

// (An "A-resident" schedule — the 128 x K block of A kept in shared memory for all n-blocks of an m-block — was measured 5-10 %
// slower than the strided order on this model's shapes and has been removed: profiles/r01_gemm_experiments.md.)
// kEW = 16: bf16-output tiles of 256 columns with four epilogue warps per scheduler instead of two.  The GELU / bias / pack
// epilogue is a ~300-instruction dependent-latency stream per 32-column chunk; two warps per scheduler issue only ~50 % of the
// cycles (ncu: long-scoreboard + fixed-latency waits), which made fc1 epilogue-bound (73 us vs 40 us of MMAs).  Each warp then
// owns 64 columns (2 chunks) and a single 2 KB staging buffer, so the operand ring keeps its depth.
template <int BN, int kCtas, int kEW = 8>
struct TcCfg {
  static_assert(kEW == 8 || (kEW == 16 && BN % 128 == 0), "16 epilogue warps need whole 32-column chunks per warp");
  static constexpr int kThreads = 64 + 32 * kEW;
  static constexpr int kStagingPerWarp = kEW == 16 ? 2048 : TC_STAGING_PER_WARP;
  static constexpr int kBNL = BN / kCtas;                          // B-tile rows loaded by one CTA
  static constexpr int kABytes = TC_BM * TC_BK * 2;
  static constexpr int kBBytes = kBNL * TC_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEW * kStagingPerWarp;
  static constexpr int kRingBudget = 227 * 1024 - kStagingBytes - 1024 - TC_BAR_BYTES;
  static constexpr int kStages = kRingBudget / kStageBytes > 8 ? 8 : kRingBudget / kStageBytes;
  static constexpr int kTmemCols = (BN == 128) ? 256 : 512;      // two accumulator buffers of BN columns, power-of-two allocation
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*alignment slack*/ + TC_BAR_BYTES;
  static_assert(kSmemBytes <= 227 * 1024 && kStages >= 3, "shared memory budget");
  static_assert((2 * kStages + 4) * 8 + 8 <= TC_BAR_BYTES, "barrier area");
};

// Tuning instrumentation (wait-cycle trace, stage-isolation switches) is compiled in only with -DSMK_GEMM_TUNE=1
// (SMK_BUILD_TUNE=1 python __graft_entry__.py build): it costs 2-5 % in registers and branches otherwise.
#ifndef SMK_GEMM_TUNE
#define SMK_GEMM_TUNE 0
#endif

// optional wait-cycle accounting (tuning scripts only): per CTA 16 counters
//   [0] producer: cycles waiting for a free ring slot   [1] producer: total loop cycles
//   [2] MMA: waiting for operands (full barriers)       [3] MMA: waiting for a free accumulator   [4] MMA: total
//   [5] epilogue warp 2: waiting for the accumulator    [6] epilogue warp 2: waiting for staging (bulk_wait_read)   [7] total
//   [8] tiles of this CTA   [9] epilogue warp 2: tcgen05.wait::ld   [10] epilogue warp 2: staging + store section
__device__ long long* g_gemm_trace = nullptr;
struct WaitClock {
  long long acc = 0, t0 = 0;
  bool on;
  __device__ explicit WaitClock(bool on_) : on(SMK_GEMM_TUNE && on_) {}
  __device__ __forceinline__ void begin() { if (SMK_GEMM_TUNE && on) t0 = clock64(); }
  __device__ __forceinline__ void end() { if (SMK_GEMM_TUNE && on) acc += clock64() - t0; }
};

// tuning-only switches in TcGemmParams::dbg (SMK_GEMM_DEBUG): results are garbage, timing isolates one pipeline stage
constexpr int TC_DBG_NOEPI = 1;    // epilogue warps hand the accumulator straight back (no TMEM read, math or store)
constexpr int TC_DBG_NOLOAD = 2;
constexpr int TC_DBG_NOSTORE = 8;  // epilogue reads TMEM and does the math but skips staging + TMA store
constexpr int TC_DBG_NOTMA = 16;   // epilogue stages to shared memory but does not issue the TMA store
constexpr int TC_DBG_NOCOMMIT = 4; // (with NOLOAD) no per-k-block tcgen05.commit on the ring's empty barriers   // operands are loaded for the first ring pass only; the MMAs re-read the same shared memory

constexpr int TC_MAX_TERMS = 3;
struct TcGemmParams {
  int M, N, K;    // K = reduction length of ONE term
  // Split-operand products: C = Σ_t A[:, ta[t] : ta[t]+K] · W[:, tw[t] : tw[t]+K]^T.  Operands stored as [hi | lo] rows (hi = T(x),
  // lo = T(x − hi)) give  A·W^T ≈ hi·hi + hi·lo + lo·hi  (terms (0,0), (0,K), (K,0): ~fp32 accuracy)  or  A_hi·(W_hi + W_lo)^T
  // (terms (0,0), (0,K): the weight rounding removed) without a second copy of hi.  One term (0,0) = the plain GEMM.
  int n_terms;
  // The terms as a product structure: n_a distinct A column offsets x n_w distinct W column offsets (<= 2 each) and the (i, j)
  // pairs to multiply.  One pipeline stage carries all distinct operand tiles of a k-block — A_hi, A_lo, W_hi, W_lo: 4 tiles for the
  // 3 products of a full split instead of the 6 a term-by-term k loop would load (these GEMMs are bound by the L2 → SM operand
  // feed: 64 B/clk per SM wanted at the MMA rate against ~43 B/clk available) — and the issuer runs every pair on them.
  int n_a, n_w, a_col[2], w_col[2], n_pairs, pair_a[TC_MAX_TERMS], pair_w[TC_MAX_TERMS];
  int n_stages;   // ring depth at this stage size (<= the compile-time maximum)
  int seq;        // 1: term-by-term k loop instead (a stage = one A tile + one W tile): only when a fused stage does not fit twice
  // fp8 correction terms (fp16 operands only): term 0's tiles hold e4m3 values — per 32 operand columns 32 bytes e4m3(hi) | 32 bytes
  // e4m3(lo·2^11) on the activation side, e4m3(lo·2^15) | e4m3(hi·2^4) on the weight side, so that one fp8 contraction over the
  // box yields (hi·lo + lo·hi)·2^15 at TWICE the 16-bit tensor rate — and are issued first; the first MMA of term 1 (fp16 hi·hi)
  // then takes the accumulator in with scale-input-d = 15.  Always term by term (seq).
  int q8;
  const float* bias;
  void* C;
  int64_t ldc;
  int epi;        // SMK_EPI_* flags
  int out_f32;    // 0 → 16-bit output (bf16 / fp16 per kF16), 1 → fp32 output, 2 → 3-part split output [hi | hi | lo] (3N columns),
                  // 3 → 2-part split output [hi | lo] (2N columns), 4 → [hi fp16 | e4m3 correction operands] (2N fp16 columns: the A
                  // operand of a terms_q8 GEMM)
  // token assembly for patch-embed (kDirect): output row = m + m / tok_hw + 1, value += tok_pos[(1 + m % tok_hw) * N + n]
  int tok_hw;
  const float* tok_pos;
  int dbg;
  int rev;        // strided schedule only: walk the tile sequence from its end (smk_kernels.h g_traverse_rev)
  // "swap-AB" form for narrow outputs (fc2: 384 output features, K = 1536): the kernel computes C^T = W · A^T, i.e. its M axis
  // runs over the output FEATURES (128-row tiles of W) and its N axis over the TOKENS (256-wide tiles of the activations), so
  // every MMA is 128 x 256 x 16 instead of 128 x 192 x 16 — an SS-mode MMA costs ~150-170 cycles whatever its N (profiles/
  // r01_gemm_experiments.md), so the wide form runs the tensor pipe at 1.7 instead of 0.9-1.3 PFLOP/s.  The epilogue then holds
  // one feature per lane and 32 tokens per chunk: bias is a per-lane scalar, the staging tile is written transposed
  // (token-major rows of 32 features) and the TMA store / reduce-add addresses C[token, feature] as usual.  fp32 output only.
  int trans;
  int credit_k;   // host-side bookkeeping only (profiler credit)
  double credit_flops;   // host-side: algorithmic FLOPs when M / N are padded (batched form); 0 = 2·M·N·credit_k
  // Batched form (mask logits: queries · tokens^T per image): problem b multiplies A rows [b·batch_a_rows + a_row0, +mb_per_batch·128)
  // with W rows [b·batch_b_rows + b_row0, +N); tiles = (b, m-block, n-block); C is a 3-D tensor map {column, row in batch, batch}
  // that clips the rows / columns beyond the valid ones.  mb_per_batch == 0: plain GEMM.
  int mb_per_batch, batch_a_rows, a_row0, batch_b_rows, b_row0;
};
struct TileCoord { int a_row, b_row, c_row, n_blk, batch; };
__device__ __forceinline__ TileCoord tile_coord(const TcGemmParams& p, int tile, int m_blocks, int n_blocks, int BN, int kCtas) {
  TileCoord t;
  if (p.mb_per_batch > 0) {
    const int per = p.mb_per_batch * n_blocks, b = tile / per, rem = tile - b * per, mb = rem / n_blocks;
    t.n_blk = rem - mb * n_blocks;
    t.batch = b;
    t.c_row = mb * TC_BM;
    t.a_row = b * p.batch_a_rows + p.a_row0 + mb * TC_BM;
    t.b_row = b * p.batch_b_rows + p.b_row0 + t.n_blk * BN;
    return t;
  }
  const int m_blk = p.trans ? tile % m_blocks : tile / n_blocks;
  t.n_blk = p.trans ? tile / m_blocks : tile % n_blocks;
  t.batch = 0;
  t.a_row = m_blk * TC_BM * kCtas;
  t.b_row = t.n_blk * BN;
  t.c_row = t.a_row;
  return t;
}

// kDirect: token assembly of the patch-embed GEMM (row re-indexing around the class-token rows + position embedding, 3-D output map)
// kCtas = 2: CTA pair (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile per pair, MMAs issued by the even CTA with
// M = 256; each CTA loads its own 128 rows of A and BN/2 rows of B, which halves the B bytes every SM pulls from L2 —
// the single-CTA kernel is bound by L2→SM bandwidth (BM·BN/(BM+BN) FLOP per operand byte: 64 at BN=128, 85 at 256).
// kQ8 (fp16 only): the fp8-corrected form (TcGemmParams::q8) and the [hi | e4m3] output kind 4 — a separate instantiation, so that the
// issue loop and the epilogue of every other GEMM stay exactly as they were (the run-time branches cost qkv 79 → 88 us)
template <int BN, bool kDirect, int kCtas, int kEW, bool kF16, bool kQ8>
__global__ void __launch_bounds__(64 + 32 * kEW, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                    const TcGemmParams p) {
  using Cfg = TcCfg<BN, kCtas, kEW>;
  using T16 = typename std::conditional<kF16, __half, __nv_bfloat16>::type;     // operand / 16-bit output type
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // operand ring
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = (p.N + BN - 1) / BN, m_blocks = (p.M + TC_BM * kCtas - 1) / (TC_BM * kCtas);   // N % BN == 0 unless trans
  const int kpt = p.K / TC_BK;                                  // k-blocks per term
  const int num_tiles = n_blocks * m_blocks, k_blocks = p.seq ? kpt * p.n_pairs : kpt;
  const int na_s = p.seq ? 1 : p.n_a, nw_s = p.seq ? 1 : p.n_w;          // operand tiles per stage
  // the term tables through registers with constant indices: a run-time index into the kernel-parameter arrays makes the compiler
  // copy them to local memory, and the single issuing thread then pays a local load per k-block on the critical path
  const int a_col0 = p.a_col[0], a_col1 = p.a_col[1], w_col0 = p.w_col[0], w_col1 = p.w_col[1];
  const int pa0 = p.pair_a[0], pa1 = p.pair_a[1], pa2 = p.pair_a[2], pw0 = p.pair_w[0], pw1 = p.pair_w[1], pw2 = p.pair_w[2];
  auto sel3 = [](int i, int x0, int x1, int x2) { return i == 0 ? x0 : (i == 1 ? x1 : x2); };
  const int stage_bytes = na_s * Cfg::kABytes + nw_s * Cfg::kBBytes, n_stages = p.n_stages;
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0u;
  // tiles are owned by clusters, strided order (neighbouring CTAs work on neighbouring tiles)
  const int n_cl = gridDim.x / kCtas, cl = blockIdx.x / kCtas;
  const int tile0 = cl, tile_end = num_tiles, tile_step = n_cl;
  long long* trace = (SMK_GEMM_TUNE && g_gemm_trace) ? g_gemm_trace + 16 * blockIdx.x : nullptr;
  const int dbg = SMK_GEMM_TUNE ? p.dbg : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEW * kCtas); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (kCtas == 2) tmem_alloc2(tmem_ptr, Cfg::kTmemCols);
    else tmem_alloc(tmem_ptr, Cfg::kTmemCols);
  }
  tc_fence_before_sync();
  if constexpr (kCtas == 2) cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive / TMA signal
  else __syncthreads();
  tc_fence_after_sync();
  // The TMEM base address is read from shared memory, i.e. into a per-thread register; routed through a warp reduction
  // (REDUX writes a uniform register) it becomes provably warp-uniform.  Otherwise ptxas wraps every single-thread
  // tcgen05.mma in an ELECT / R2UR.BROADCAST / branch "waterfall" that costs ~100 cycles per MMA — more than a
  // 128 x 128 x 16 MMA (64 cycles) takes to execute, which made the issuing thread the bottleneck of every GEMM.
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_ptr);
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      WaitClock w_slot(trace != nullptr), w_all(trace != nullptr);
      w_all.begin();
      int n_loads = 0;
      for (int tile_i = tile0; tile_i < tile_end; tile_i += tile_step) {
        const int tile = p.rev ? num_tiles - 1 - tile_i : tile_i;
        // trans: the feature blocks of one token block are neighbours in the sequence (they share the activation tile in L2)
        const TileCoord tc_ = tile_coord(p, tile, m_blocks, n_blocks, BN, kCtas);
        for (int kb = 0; kb < k_blocks; ++kb) {
          if ((dbg & TC_DBG_NOLOAD) && n_loads++ >= n_stages) continue;
          w_slot.begin();
          mbar_wait(&empty_bar[stage], phase ^ 1);
          w_slot.end();
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + na_s * Cfg::kABytes;
          // seq: k-block kb belongs to term kb / kpt; fused: to all terms at once
          const int term = p.seq ? kb / kpt : 0, kc = (p.seq ? kb - term * kpt : kb) * TC_BK;
          const int ia0 = p.seq ? sel3(term, pa0, pa1, pa2) : 0, iw0 = p.seq ? sel3(term, pw0, pw1, pw2) : 0;
          if constexpr (kCtas == 2) {
            // both CTAs' bytes complete on the even CTA's barrier (its MMA thread is the only consumer)
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            for (int i = 0; i < na_s; ++i) tma_load_2d_cg2(sa + i * Cfg::kABytes, &tmA, bar, (ia0 + i ? a_col1 : a_col0) + kc, tc_.a_row + (int)rank * TC_BM);
            for (int j = 0; j < nw_s; ++j) tma_load_2d_cg2(sb + j * Cfg::kBBytes, &tmB, bar, (iw0 + j ? w_col1 : w_col0) + kc, tc_.b_row + (int)rank * Cfg::kBNL);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
            for (int i = 0; i < na_s; ++i) tma_load_2d(sa + i * Cfg::kABytes, &tmA, &full_bar[stage], (ia0 + i ? a_col1 : a_col0) + kc, tc_.a_row);
            for (int j = 0; j < nw_s; ++j) tma_load_2d(sb + j * Cfg::kBBytes, &tmB, &full_bar[stage], (iw0 + j ? w_col1 : w_col0) + kc, tc_.b_row);
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
      }
      w_all.end();
      if (trace) { trace[0] = w_slot.acc; trace[1] = w_all.acc; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = idesc_16_f32<kF16>(TC_BM * kCtas, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      WaitClock w_ops(trace != nullptr), w_acc(trace != nullptr), w_all(trace != nullptr);
      w_all.begin();
      int n_used = 0;
      for (int tile = tile0; tile < tile_end; tile += tile_step) {
        w_acc.begin();
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        w_acc.end();
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          w_ops.begin();
          const bool skip_wait = (dbg & TC_DBG_NOLOAD) && n_used++ >= n_stages;
          if (!skip_wait) mbar_wait(&full_bar[stage], phase);
          w_ops.end();
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + (uint32_t)(na_s * Cfg::kABytes);
          const int n_pr = p.seq ? 1 : p.n_pairs;
          for (int pr = 0; pr < n_pr; ++pr) {
            const uint64_t a_desc = smem_desc_k_sw128(sa + (uint32_t)((p.seq ? 0 : sel3(pr, pa0, pa1, pa2)) * Cfg::kABytes));
            const uint64_t b_desc = smem_desc_k_sw128(sb + (uint32_t)((p.seq ? 0 : sel3(pr, pw0, pw1, pw2)) * Cfg::kBBytes));
            if (kQ8 && kb <= kpt) {
              if (kb < kpt) {           // fp8 tiles: four K = 32 products per 128-byte row
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if constexpr (kCtas == 2) umma_f8_ss_cg2(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                  else umma_f8_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                }
              } else {                  // first fp16 k-block: D = hi·hi + D·2^-15
                if constexpr (kCtas == 2) umma_16_ss_cg2_sd15(d_tmem, a_desc, b_desc, idesc);
                else umma_16_ss_sd15(d_tmem, a_desc, b_desc, idesc);
#pragma unroll
                for (int k = 1; k < TC_BK / 16; ++k) {
                  if constexpr (kCtas == 2) umma_bf16_ss_cg2(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, 1);
                  else umma_bf16_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, 1);
                }
              }
              continue;
            }
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {  // +32 B per 16-element K step → +2 in the (addr >> 4) field
              if constexpr (kCtas == 2) umma_bf16_ss_cg2(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | pr | k) != 0);
              else umma_bf16_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | pr | k) != 0);
            }
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if (!(dbg & TC_DBG_NOCOMMIT)) {
            if constexpr (kCtas == 2) tc_commit_cg2(&empty_bar[stage], 3);
            else tc_commit(&empty_bar[stage]);
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete → epilogue (of both CTAs)
        if constexpr (kCtas == 2) tc_commit_cg2(&tmem_full[acc], 3);
        else tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      w_all.end();
      if (trace) { trace[2] = w_ops.acc; trace[3] = w_acc.acc; trace[4] = w_all.acc; }
    }
  } else {
    // ===== epilogue: warps 2..; lane quarter = warp % 4 (hardware rule), column group = (warp - 2) / 4 =====
    const int quarter = warp & 3;
    const int col_half = (warp - 2) >> 2;
    constexpr int kChunks = BN / (8 * kEW);      // 32-column chunks per warp per tile
    uint8_t* stg = staging + (warp - 2) * Cfg::kStagingPerWarp;
    const uint32_t stg_u32 = smem_u32(stg);
    auto r4 = [](float f) { return __float_as_uint(f); };
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t it = 0;                             // staging-buffer parity (bf16 output)
    const uint32_t tmem_empty_remote = kCtas == 2 ? mapa_shared(smem_u32(&tmem_empty[0]), 0) : 0u;
    const bool tr_w = trace != nullptr && warp == 2 && lane == 0;
    WaitClock w_tm(tr_w), w_stg(tr_w), w_all(tr_w), w_ld(tr_w), w_st(tr_w);
    long long n_tiles_done = 0;
    w_all.begin();
    for (int tile_i = tile0; tile_i < tile_end; tile_i += tile_step) {
      const int tile = p.rev ? num_tiles - 1 - tile_i : tile_i;
      const TileCoord tc_ = tile_coord(p, tile, m_blocks, n_blocks, BN, kCtas);
      const int m_blk = tc_.c_row / TC_BM + (int)rank, n_blk = tc_.n_blk;     // c_row: first output row of the tile (row in batch when batched)
      ++n_tiles_done;
      w_tm.begin();
      mbar_wait(&tmem_full[acc], acc_phase);
      w_tm.end();
      tc_fence_after_sync();
      if (dbg & TC_DBG_NOEPI) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2 && rank != 0) mbar_arrive_cluster(tmem_empty_remote + (uint32_t)(acc * 8));
          else mbar_arrive(&tmem_empty[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      const int m = m_blk * TC_BM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        const int c = col_half * kChunks + ci;
        const int n0 = n_blk * BN + c * 32;
        uint32_t r[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), r);
        float4 bv[kEW == 16 ? 1 : 8];
        if (kEW != 16 && p.bias && !p.trans) {   // L1 hits after the first tile; in flight under the TMEM load
#pragma unroll
          for (int j = 0; j < (kEW == 16 ? 1 : 8); ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
        }
        w_ld.begin();
        tmem_ld_wait();
        w_ld.end();
        if (ci == kChunks - 1) {                 // accumulator fully read: hand the TMEM buffer back before the math
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (kCtas == 2 && rank != 0) mbar_arrive_cluster(tmem_empty_remote + (uint32_t)(acc * 8));
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias && p.trans) {                 // swap-AB: the lane is an output feature
          const float bm = m < p.M ? __ldg(p.bias + m) : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += bm;
        } else if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            // 16-warp variant (<= 112 registers): no prefetched copy, the (L1-resident, warp-uniform) bias is read at use
            const float4 b = kEW == 16 ? __ldg(reinterpret_cast<const float4*>(p.bias + n0) + (j >> 2)) : bv[kEW == 16 ? 0 : (j >> 2)];
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (p.epi & SMK_EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
        }
        if (p.epi & SMK_EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if constexpr (kDirect) {
          // token assembly (patch embed): tiles are per IMAGE (batched form: tile rows = patches of one image, the rows past its hw
          // patches compute the next image's and are clipped), so a 32-row chunk never crosses an image: patch r of image b → token row
          // 1 + r (row 0 is the class token) through a 3-D map {column, patch, image} based at token row 1; position embedding added
          // here.  (A flat tiling with one store per image touched needs NEGATIVE start coordinates for the second image: the TMA store
          // raises an illegal-instruction fault on those.)
          const int r_first = tc_.c_row + quarter * 32;
          {
            const int r = r_first + lane < p.tok_hw ? r_first + lane : p.tok_hw - 1;
            const float* pos_row = p.tok_pos + (int64_t)(1 + r) * p.N;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(pos_row + n0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
          const uint32_t srow = stg_u32 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(srow + ((j ^ (lane & 7)) << 4), r4(v[4 * j]), r4(v[4 * j + 1]), r4(v[4 * j + 2]), r4(v[4 * j + 3]));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && r_first < p.tok_hw) {
            tma_store_3d(&tmC, stg, n0, r_first, tc_.batch);
            bulk_commit();
          }
        } else {
          const int row0 = m_blk * TC_BM + quarter * 32;
          if (dbg & TC_DBG_NOSTORE) {
            float acc_sum = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc_sum += v[j];
            if (acc_sum == 1.2345e-30f) reinterpret_cast<float*>(p.C)[0] = acc_sum;   // keeps the math alive
            continue;
          }
          w_st.begin();
          if (kEW != 16 && p.out_f32 == 1 && p.trans) {
            // swap-AB: v[j] = C[token n0 + j, feature row0 + lane].  Staging tile: 32 token rows x 32 features (128 B), 128-byte
            // swizzle; for a fixed token the 32 lanes fill one 128-byte row (conflict-free 4-byte stores)
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            const uint32_t lcol = (uint32_t)(lane >> 2), lsub = (uint32_t)(lane & 3) << 2;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg_u32 + (uint32_t)j * 128u + (((lcol ^ (uint32_t)(j & 7))) << 4) + lsub), "r"(r4(v[j])) : "memory");
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (p.epi & SMK_EPI_RESIDUAL) tma_reduce_add_2d(&tmC, stg, row0, n0);
              else tma_store_2d(&tmC, stg, row0, n0);
              bulk_commit();
            }
          } else if (kEW != 16 && p.out_f32 == 1) {
            // staging tile: 32 rows x 128 B, 128-byte swizzle (16-byte chunk index ^= row & 7)
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            const uint32_t srow = stg_u32 + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(srow + ((j ^ (lane & 7)) << 4), r4(v[4 * j]), r4(v[4 * j + 1]), r4(v[4 * j + 2]), r4(v[4 * j + 3]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (p.mb_per_batch > 0) tma_store_3d(&tmC, stg, n0, row0, tc_.batch);
              else if (p.epi & SMK_EPI_RESIDUAL) tma_reduce_add_2d(&tmC, stg, n0, row0);
              else tma_store_2d(&tmC, stg, n0, row0);
              bulk_commit();
            }
          } else if (kEW != 16 && p.out_f32 >= 2) {
            // split output (two 32 x 64 B staging tiles): 3-part [hi | hi | lo]: hi → columns n0 and N + n0, lo → 2N + n0;
            // 2-part [hi | lo]: hi → n0, lo → N + n0
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            const uint32_t srow = stg_u32 + lane * 64;
            if (kQ8 && p.out_f32 == 4) {
              // [hi | e4m3 correction operands]: the second 64-byte box holds e4m3(hi) (32 B) and e4m3(lo·2^11) (32 B) of the 32 columns
              uint32_t f8[8], s8[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint2 ha, hb;
                split_q8x4<false>(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3], ha, f8[2 * j], s8[2 * j]);
                split_q8x4<false>(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7], hb, f8[2 * j + 1], s8[2 * j + 1]);
                st_shared_v4(srow + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4), ha.x, ha.y, hb.x, hb.y);
              }
              const uint32_t sw = (uint32_t)((lane >> 1) & 3);
              st_shared_v4(srow + 2048 + ((0u ^ sw) << 4), f8[0], f8[1], f8[2], f8[3]);
              st_shared_v4(srow + 2048 + ((1u ^ sw) << 4), f8[4], f8[5], f8[6], f8[7]);
              st_shared_v4(srow + 2048 + ((2u ^ sw) << 4), s8[0], s8[1], s8[2], s8[3]);
              st_shared_v4(srow + 2048 + ((3u ^ sw) << 4), s8[4], s8[5], s8[6], s8[7]);
            } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split16x2<T16>(v[8 * j + 2 * e], v[8 * j + 2 * e + 1], h[e], l[e]);
              const uint32_t off = (uint32_t)((j ^ ((lane >> 1) & 3)) << 4);
              st_shared_v4(srow + off, h[0], h[1], h[2], h[3]);
              st_shared_v4(srow + 2048 + off, l[0], l[1], l[2], l[3]);
            }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, stg, n0, row0);
              if (p.out_f32 == 2) {
                tma_store_2d(&tmC, stg, p.N + n0, row0);
                tma_store_2d(&tmC, stg + 2048, 2 * p.N + n0, row0);
              } else {
                tma_store_2d(&tmC, stg + 2048, p.N + n0, row0);
              }
              bulk_commit();
            }
          } else if (kEW != 16 && p.out_f32 == 0 && p.trans) {
            // swap-AB, bf16 output: v[j] = C[token n0 + j, feature row0 + lane].  Staging tile: 32 token rows x 32 features (64 B),
            // 64-byte swizzle, two buffers; for a fixed token the 32 lanes fill one 64-byte row with 2-byte stores
            uint8_t* buf = stg + (it & 1) * 2048;
            ++it;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            const uint32_t sb = stg_u32 + (uint32_t)(buf - stg) + ((uint32_t)(lane & 7) << 1), lchunk = (uint32_t)(lane >> 3);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const unsigned short hv = (unsigned short)(Pack16<T16>::pack(v[j], 0.f) & 0xffffu);
              asm volatile("st.shared.b16 [%0], %1;" ::"r"(sb + (uint32_t)j * 64u + ((lchunk ^ (uint32_t)((j >> 1) & 3)) << 4)), "h"(hv) : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, buf, row0, n0);
              bulk_commit();
            }
          } else {
            // staging tile: 32 rows x 64 B, 64-byte swizzle (16-byte chunk index ^= (row >> 1) & 3), two buffers
            uint8_t* buf = kEW == 16 ? stg : stg + (it & 1) * 2048;      // 16-warp variant: one buffer per warp
            ++it;
            if (lane == 0) {
              if constexpr (kEW == 16) bulk_wait_read<0>();
              else bulk_wait_read<1>();
            }
            __syncwarp();
            const uint32_t srow = stg_u32 + (uint32_t)(buf - stg) + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(srow + ((j ^ ((lane >> 1) & 3)) << 4), Pack16<T16>::pack(v[8 * j], v[8 * j + 1]), Pack16<T16>::pack(v[8 * j + 2], v[8 * j + 3]),
                           Pack16<T16>::pack(v[8 * j + 4], v[8 * j + 5]), Pack16<T16>::pack(v[8 * j + 6], v[8 * j + 7]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(dbg & TC_DBG_NOTMA)) {
              tma_store_2d(&tmC, buf, n0, row0);
              bulk_commit();
            }
          }
          w_st.end();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    w_all.end();
    if (lane == 0) bulk_wait<0>();   // all tile stores complete before the CTA retires
    if (tr_w) { trace[5] = w_tm.acc; trace[6] = w_stg.acc; trace[7] = w_all.acc; trace[8] = n_tiles_done; trace[9] = w_ld.acc; trace[10] = w_st.acc; }
  }
  tc_fence_before_sync();
  if constexpr (kCtas == 2) {
    cluster_sync_all();        // the peer may still read this CTA's operand tiles / signal its barriers until here
    if (warp == 1) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

int make_tmap_nd(CUtensorMap* out, int elem_bytes, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, int swizzle_bytes) {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      g_encode = (PFN_encodeTiled)fn;
  });
  if (!g_encode) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return SMK_ERR_CUDA; }
  SMK_REQUIRE(rank == 2 || rank == 3, "tensor map: rank %d unsupported", rank);
  SMK_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "tensor map: element size %d unsupported", elem_bytes);
  SMK_REQUIRE(((uintptr_t)base % 16) == 0, "tensor map: base must be 16-byte aligned");
  SMK_REQUIRE(swizzle_bytes == 0 || (int)box[0] * elem_bytes <= swizzle_bytes, "tensor map: box width %u exceeds the %d-byte swizzle span",
              box[0], swizzle_bytes);
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  switch (swizzle_bytes) {
    case 0: break;
    case 32: sw = CU_TENSOR_MAP_SWIZZLE_32B; break;
    case 64: sw = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 128: sw = CU_TENSOR_MAP_SWIZZLE_128B; break;
    default: set_error("tensor map: swizzle %d unsupported", swizzle_bytes); return SMK_ERR_INVALID;
  }
  cuuint64_t d[3], st[2];
  cuuint32_t bx[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    SMK_REQUIRE(box[i] >= 1 && box[i] <= 256 && dims[i] >= 1, "tensor map: bad box/dim %d", i);
    d[i] = dims[i];
    bx[i] = box[i];
    if (i + 1 < rank) {
      SMK_REQUIRE(strides_bytes[i] % 16 == 0, "tensor map: stride %d must be a multiple of 16 bytes", i);
      st[i] = strides_bytes[i];
    }
  }
  CUresult r = g_encode(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                        const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SMK_ERR_CUDA; }
  return SMK_OK;
}

int make_tmap_2d(CUtensorMap* out, int elem_bytes, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  const uint64_t dims[2] = {inner, outer}, strides[1] = {row_stride_bytes};
  const uint32_t box[2] = {box_inner, box_outer};
  return make_tmap_nd(out, elem_bytes, base, 2, dims, strides, box, swizzle_bytes);
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  SMK_REQUIRE(box_inner * 2 == 128, "tensor map: operand box must be 128 bytes wide");
  return make_tmap_2d(out, 2, base, inner, outer, row_stride_bytes, box_inner, box_outer, 128);
}

static int num_sms() { return device_sm_count(); }

// terms → distinct A / W column offsets + the (i, j) pairs; `swap`: the kernel's A operand is the caller's W (swap-AB form)
static int set_terms(TcGemmParams& p, const GemmTerms& tr, bool swap) {
  p.n_terms = tr.n;
  p.n_a = p.n_w = p.n_pairs = 0;
  for (int t = 0; t < tr.n; ++t) {
    const int ao = swap ? tr.w_off[t] : tr.a_off[t], wo = swap ? tr.a_off[t] : tr.w_off[t];
    int ia = -1, iw = -1;
    for (int i = 0; i < p.n_a; ++i) if (p.a_col[i] == ao) ia = i;
    for (int i = 0; i < p.n_w; ++i) if (p.w_col[i] == wo) iw = i;
    if (ia < 0) { SMK_REQUIRE(p.n_a < 2, "gemm_tc: more than two distinct A parts"); ia = p.n_a; p.a_col[p.n_a++] = ao; }
    if (iw < 0) { SMK_REQUIRE(p.n_w < 2, "gemm_tc: more than two distinct W parts"); iw = p.n_w; p.w_col[p.n_w++] = wo; }
    p.pair_a[p.n_pairs] = ia;
    p.pair_w[p.n_pairs++] = iw;
  }
  return SMK_OK;
}

template <int BN, bool kDirect, int kCtas, int kEW, bool kF16, bool kQ8>
static int launch_tc_q(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm, const TcGemmParams& p_in, cudaStream_t s) {
  using Cfg = TcCfg<BN, kCtas, kEW>;
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, kDirect, kCtas, kEW, kF16, kQ8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
  }
  TcGemmParams p = p_in;
  {
    // fused stages (all distinct operand tiles of a k-block, every product on them) or the term-by-term loop; SMK_GEMM_FUSE_TERMS = 0 / 1 forces
    static int force = -2;
    if (force == -2) {
      const char* e = getenv("SMK_GEMM_FUSE_TERMS");
      force = e ? atoi(e) : -1;
    }
    const int fused_stages = (Cfg::kStages * Cfg::kStageBytes) / (p.n_a * Cfg::kABytes + p.n_w * Cfg::kBBytes);
    // measured on B200 (same box, fp16s step, profiles/r02_gemm_terms.md): fused stages 8.78 ms per step against 10.1 ms term by term
    const bool fuse = force >= 0 ? (force != 0 && fused_stages >= 2) : fused_stages >= 2;
    p.seq = (p.n_pairs > 1 && (!fuse || p.q8)) ? 1 : 0;
    p.n_stages = p.seq || p.n_pairs == 1 ? Cfg::kStages : fused_stages;
  }
  const int tiles = (p.N / BN) * ((p.M + TC_BM * kCtas - 1) / (TC_BM * kCtas));
  const int slots = num_sms() / kCtas;                     // CTAs (kCtas = 1) or CTA pairs (2) that can be resident
  const int grid = (tiles < slots ? tiles : slots) * kCtas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (kCtas > 1) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = kCtas;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  if (pdl_enabled()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  {
    // credited work = ALGORITHMIC FLOPs of the contraction (one term; p.credit_k is the mathematical reduction length), whatever
    // number of split terms the tensor core is issued
    ProfScope prof(PROF_GEMM_TC, p.credit_flops > 0 ? p.credit_flops : 2.0 * p.M * p.N * p.credit_k, s, 2.0 * p.M * p.N * p.K * (p.q8 ? 3 : p.n_terms));
    SMK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_kernel<BN, kDirect, kCtas, kEW, kF16, kQ8>, ta, tb, tcm, p));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

template <int BN, bool kDirect, int kCtas, int kEW, bool kF16>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm, const TcGemmParams& p, cudaStream_t s) {
  if constexpr (kF16 && kEW == 8) {
    if (p.q8) return launch_tc_q<BN, kDirect, kCtas, kEW, kF16, true>(ta, tb, tcm, p, s);
  }
  SMK_REQUIRE(!p.q8 && p.out_f32 != 4, "gemm_tc: the fp8-corrected form / [hi | e4m3] output needs fp16 operands, q8 terms and 8 epilogue warps");
  return launch_tc_q<BN, kDirect, kCtas, kEW, kF16, false>(ta, tb, tcm, p, s);
}

// Tile width: minimise (waves over the SMs) x (BN + 64): measured, a k-block of MMAs costs a fixed part plus a part
// proportional to BN (SS-mode MMA floor and the A-tile feed do not shrink with BN; profiles/r01_gemm_experiments.md), so
// narrow tiles only pay off when they save whole waves; ties go to the wider tile (fewer operand bytes per MMA cycle).
// `slots` = resident CTAs (or CTA pairs), `bm` = rows per tile.
static int pick_bn(int M, int N, int bm, int slots) {
  static int forced = -1;                      // SMK_GEMM_BN=128|192|256 forces the tile width (tuning aid)
  if (forced < 0) {
    const char* e = getenv("SMK_GEMM_BN");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0 && N % forced == 0) return forced;
  const int64_t mb = (M + bm - 1) / bm;
  int best = 128;
  int64_t best_cost = -1;
  for (int bn : {128, 192, 256}) {
    if (N % bn) continue;
    const int64_t tiles = (N / bn) * mb, cost = (tiles + slots - 1) / slots * (bn + 64);
    if (best_cost < 0 || cost <= best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

// CTA pairs (cta_group::2, 256-row tiles, each CTA loads half of the weight tile): measured slower at K = 384 (the extra
// cross-CTA hand-shakes per k-block are not amortised) and faster from K >= 1024 on many-row problems, where the GEMM sits at
// the L2→SM operand-feed limit (fc2: 73.0 → 66.8 us).  SMK_GEMM_CTA_PAIR = 0 / 1 forces single CTAs / pairs (tuning).
// K here is the TOTAL reduction length issued (all split terms).
static bool use_cta_pair(int M, int K) {
  static int mode = -2;
  if (mode == -2) {
    const char* e = getenv("SMK_GEMM_CTA_PAIR");
    mode = e ? atoi(e) : -1;
  }
  if (mode >= 0) return mode != 0 && M > TC_BM;
  return K >= 1024 && M >= 16384;
}

// swap-AB form (TcGemmParams::trans): narrow fp32 outputs of long-K, many-row problems (fc2).  SMK_GEMM_SWAP_AB = 0 / 1 forces it
// off / on for every eligible call (tuning).  Measured on B200 (kernel_bench, same box): fc2 65.9 / 66.2 → 64.8 / 64.7 us, proj
// (K = 384, forced) 36.8 → 39.2 us: fc2 moves 310 MB at 4.8 TB/s either way — the fp32 residual read-modify-write and the
// hidden-tensor read, not the MMA width, bound it — so the wide MMAs only buy 2 %.
static bool use_swap_ab(int M, int N, int K) {
  static int mode = -2;
  if (mode == -2) {
    const char* e = getenv("SMK_GEMM_SWAP_AB");
    mode = e ? atoi(e) : -1;
  }
  if (N % TC_BM != 0) return false;
  if (mode >= 0) return mode != 0;
  return N <= 512 && K >= 1024 && M >= 16384;
}

// swap-AB for 16-bit outputs whose width is not a multiple of 256 (qkv: 1152 = 9 x 128 features → every MMA 128 x 256 over tokens
// instead of 128 x 192; 1773 tiles = 11.98 waves of 148).  Single-term bf16 (round 1): 48.4 us normal vs 49.5 us swapped — paced by its
// epilogue, and the transposed store costs one 2-byte st.shared per element: off.  With split terms the k loop streams 64 KB of
// operands per k-block either way, so the wider tile is 25 % fewer bytes per FLOP into the SM: the fp16s qkv (A_hi·(W_hi + W_lo))
// measures 87.5 → 71.0 us swapped (scripts/gemm_q8_bench.py --only qkv; CTA pairs: 83.3 us): ON for multi-term products.
// SMK_GEMM_SWAP_AB = 0 / 1 forces.  Never for GELU epilogues (fc1).
static bool use_swap_ab_bf16(int M, int N, int epi, int n_terms) {
  static int mode = -2;
  if (mode == -2) {
    const char* e = getenv("SMK_GEMM_SWAP_AB");
    mode = e ? atoi(e) : -1;
  }
  if (N % TC_BM != 0 || (epi & SMK_EPI_GELU)) return false;
  if (N % 256 == 0 || M < 16384) return false;
  return mode >= 0 ? mode == 1 : n_terms >= 2;
}

// 16 epilogue warps: 16-bit-output 256-column tiles of many-row problems (fc1, memory K/V); SMK_GEMM_EPI16 = 0 / 1 overrides (tuning)
static bool use_epi16(const TcGemmParams& p) {
  static int mode = -2;
  if (mode == -2) {
    const char* e = getenv("SMK_GEMM_EPI16");
    mode = e ? atoi(e) : -1;
  }
  if (p.out_f32 != 0) return false;   // (split outputs with 16 epilogue warps on CTA pairs: measured slower, fc1 148 vs 141 us)
  return mode >= 0 ? mode != 0 : p.M >= 16384;
}

template <bool kDirect, int kCtas, bool kF16>
static int launch_bn(int BN, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm, const TcGemmParams& p, cudaStream_t s) {
  if constexpr (!kDirect) {
    if (BN == 256 && !p.q8 && use_epi16(p)) return launch_tc<256, false, kCtas, 16, kF16>(ta, tb, tcm, p, s);
  }
  return BN == 256 ? launch_tc<256, kDirect, kCtas, 8, kF16>(ta, tb, tcm, p, s)
                   : (BN == 192 ? launch_tc<192, kDirect, kCtas, 8, kF16>(ta, tb, tcm, p, s) : launch_tc<128, kDirect, kCtas, 8, kF16>(ta, tb, tcm, p, s));
}

// General form.  A [M, >= lda] / W [N, >= ldw] hold 16-bit operands (bf16, or fp16 when f16 != 0); every term t multiplies the K
// columns of A starting at terms.a_off[t] with the K columns of W starting at terms.w_off[t] (element offsets, multiples of 64).
template <bool kF16>
static int gemm_tc_impl(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                        int epi, int out_f32, int tok_hw, const float* tok_pos, const GemmTerms& tr, int credit_k, cudaStream_t s) {
  SMK_REQUIRE(K % TC_BK == 0 && N % 128 == 0, "gemm_tc: need K %% 64 == 0 and N %% 128 == 0 (K=%d N=%d)", K, N);
  SMK_REQUIRE(tr.n >= 1 && tr.n <= TC_MAX_TERMS, "gemm_tc: 1..3 terms");
  int64_t a_cols = 0, w_cols = 0;
  for (int t = 0; t < tr.n; ++t) {
    SMK_REQUIRE(tr.a_off[t] >= 0 && tr.w_off[t] >= 0 && tr.a_off[t] % TC_BK == 0 && tr.w_off[t] % TC_BK == 0, "gemm_tc: term offsets must be multiples of 64");
    a_cols = std::max<int64_t>(a_cols, tr.a_off[t] + K);
    w_cols = std::max<int64_t>(w_cols, tr.w_off[t] + K);
  }
  SMK_REQUIRE(lda >= a_cols && ldw >= w_cols, "gemm_tc: operand rows shorter than the terms reach (lda=%lld ldw=%lld)", (long long)lda, (long long)ldw);
  SMK_REQUIRE(out_f32 >= 0 && out_f32 <= 4 && (out_f32 != 4 || kF16), "gemm_tc: out_f32 must be 0 (16-bit), 1 (fp32), 2 ([hi|hi|lo] split), 3 ([hi|lo] split) or 4 ([hi|q8], fp16)");
  SMK_REQUIRE(!(epi & SMK_EPI_RESIDUAL) || out_f32 == 1, "gemm_tc: residual epilogue needs fp32 output");
  SMK_REQUIRE(out_f32 != 2 || ldc >= 3 * (int64_t)N, "gemm_tc: split output needs ldc >= 3N");
  SMK_REQUIRE(out_f32 < 3 || ldc >= 2 * (int64_t)N, "gemm_tc: split output needs ldc >= 2N");
  SMK_REQUIRE(ldc % 8 == 0 && ((uintptr_t)C % 16) == 0, "gemm_tc: C must be 16-byte aligned with ldc %% 8 == 0");
  SMK_REQUIRE(!bias || ((uintptr_t)bias % 16) == 0, "gemm_tc: bias must be 16-byte aligned");
  SMK_REQUIRE(tok_hw == 0 || (out_f32 == 1 && !(epi & SMK_EPI_RESIDUAL) && tok_pos), "gemm_tc: token assembly needs a plain fp32 output");
  if (M == 0) return SMK_OK;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("SMK_GEMM_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  const int Ktot = K * (tr.q8 ? 3 : tr.n);   // tile-form heuristics go by the operand bytes streamed per output tile: a q8 row carries the same 4K bytes as [hi | lo]
  TcGemmParams p{};
  p.M = M; p.N = N; p.K = K;
  SMK_PROPAGATE(set_terms(p, tr, false));
  p.bias = bias; p.C = C; p.ldc = ldc; p.epi = epi; p.out_f32 = out_f32; p.tok_hw = tok_hw; p.tok_pos = tok_pos; p.dbg = dbg;
  p.rev = traverse_dir(); p.trans = 0; p.credit_k = credit_k > 0 ? credit_k : K;
  p.q8 = tr.q8;
  SMK_REQUIRE(!tr.q8 || (kF16 && tr.n == 2 && tr.a_off[0] == K && tr.w_off[0] == K && tr.a_off[1] == 0 && tr.w_off[1] == 0),
              "gemm_tc: fp8 correction terms need fp16 operands and the terms_q8 layout");
  if (tok_hw == 0 && ((out_f32 == 1 && use_swap_ab(M, N, Ktot)) || (out_f32 == 0 && use_swap_ab_bf16(M, N, epi, tr.n)))) {
    // C^T = W · A^T: kernel M axis = output features (N), kernel N axis = tokens (M); see TcGemmParams::trans
    CUtensorMap ta, tb, tcm;
    SMK_PROPAGATE(make_tmap_bf16_2d(&ta, W, (uint64_t)w_cols, (uint64_t)N, (uint64_t)ldw * 2, TC_BK, TC_BM));
    SMK_PROPAGATE(make_tmap_bf16_2d(&tb, A, (uint64_t)a_cols, (uint64_t)M, (uint64_t)lda * 2, TC_BK, 256));
    SMK_PROPAGATE(make_tmap_2d(&tcm, out_f32 ? 4 : 2, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * (out_f32 ? 4 : 2), 32, 32, out_f32 ? 128 : 64));
    p.M = N; p.N = M; p.trans = 1;
    SMK_PROPAGATE(set_terms(p, tr, true));
    return launch_tc<256, false, 1, 8, kF16>(ta, tb, tcm, p, s);
  }
  // q8: single CTAs (fc1 133 us against 141 us as CTA pairs, scripts/gemm_q8_bench.py) — the pair rule goes by the MMA time per k-block
  const bool pair = tok_hw == 0 && use_cta_pair(M, K * tr.n);
  const int kc = pair ? 2 : 1;
  int BN = pick_bn(M, N, TC_BM * kc, num_sms() / kc);
  if (tok_hw > 0) BN = pick_bn(((M / tok_hw) * ((tok_hw + TC_BM - 1) / TC_BM)) * TC_BM, N, TC_BM, num_sms());
  CUtensorMap ta, tb, tcm;
  SMK_PROPAGATE(make_tmap_bf16_2d(&ta, A, (uint64_t)a_cols, (uint64_t)M, (uint64_t)lda * 2, TC_BK, TC_BM));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tb, W, (uint64_t)w_cols, (uint64_t)N, (uint64_t)ldw * 2, TC_BK, (uint32_t)(BN / kc)));
  if (tok_hw > 0) {
    SMK_REQUIRE(M % tok_hw == 0, "gemm_tc: token assembly needs whole images (M %% hw == 0)");
    // per-image tiles: ceil(hw / 128) m-blocks per image, A rows of image b from row b·hw (batched form, weights shared)
    const int n_img = M / tok_hw, mb = (tok_hw + TC_BM - 1) / TC_BM;
    p.mb_per_batch = mb; p.batch_a_rows = tok_hw; p.a_row0 = 0; p.batch_b_rows = 0; p.b_row0 = 0;
    p.M = n_img * mb * TC_BM;
    p.credit_flops = 2.0 * M * N * p.credit_k;
    const uint64_t dims[3] = {(uint64_t)N, (uint64_t)tok_hw, (uint64_t)(M / tok_hw)};
    const uint64_t strides[2] = {(uint64_t)ldc * 4, (uint64_t)(tok_hw + 1) * ldc * 4};
    const uint32_t box[3] = {32u, 32u, 1u};
    SMK_PROPAGATE(make_tmap_nd(&tcm, 4, (float*)C + ldc, 3, dims, strides, box, 128));     // based at token row 1 of image 0
    return launch_bn<true, 1, kF16>(BN, ta, tb, tcm, p, s);
  }
  // output tiles of 32 rows x 32 columns per epilogue warp: 128 B (fp32, 128-byte swizzle) or 64 B (16-bit, 64-byte swizzle) per row
  const int esz = out_f32 == 1 ? 4 : 2;
  const int parts = out_f32 == 2 ? 3 : (out_f32 >= 3 ? 2 : 1);
  SMK_PROPAGATE(make_tmap_2d(&tcm, esz, C, (uint64_t)(parts * N), (uint64_t)M, (uint64_t)ldc * esz, 32, 32, out_f32 == 1 ? 128 : 64));
  if (pair) return launch_bn<false, 2, kF16>(BN, ta, tb, tcm, p, s);
  return launch_bn<false, 1, kF16>(BN, ta, tb, tcm, p, s);
}

// Batched C_b = A_b · W_b^T (fp32 output, no bias): problem b reads `rows_a` rows of A from row b·batch_a_rows + a_row0 and `rows_w`
// rows of W from row b·batch_w_rows + w_row0; C [n_batch, rows_a, rows_w] fp32 contiguous.  Tiles are 128 x 256 per (batch, m, n)
// block; operand rows beyond a problem's own (the next problem's, or zeros at the end of the tensor) only reach output rows /
// columns that the 3-D output map clips.  Serves the mask-logit contraction (queries x memory tokens per image).
template <bool kF16>
static int gemm_tc_batched_impl(const void* A, int64_t lda, int64_t a_total_rows, int batch_a_rows, int a_row0, int rows_a, const void* W,
                                int64_t ldw, int64_t w_total_rows, int batch_w_rows, int w_row0, int rows_w, float* C, int n_batch, int K,
                                const GemmTerms& tr, cudaStream_t s) {
  SMK_REQUIRE(K % TC_BK == 0 && tr.n >= 1 && tr.n <= TC_MAX_TERMS && n_batch >= 0 && rows_a >= 1 && rows_w >= 1, "gemm_tc_batched: bad shape");
  SMK_REQUIRE(((uintptr_t)C % 16) == 0 && rows_w % 4 == 0, "gemm_tc_batched: C rows must be 16-byte aligned (rows_w %% 4 == 0)");
  if (n_batch == 0) return SMK_OK;
  int64_t a_cols = 0, w_cols = 0;
  for (int t = 0; t < tr.n; ++t) {
    a_cols = std::max<int64_t>(a_cols, tr.a_off[t] + K);
    w_cols = std::max<int64_t>(w_cols, tr.w_off[t] + K);
  }
  SMK_REQUIRE(lda >= a_cols && ldw >= w_cols, "gemm_tc_batched: operand rows shorter than the terms reach");
  constexpr int BN = 256;
  const int mb = (rows_a + TC_BM - 1) / TC_BM, nb = (rows_w + BN - 1) / BN;
  TcGemmParams p{};
  p.M = n_batch * mb * TC_BM; p.N = nb * BN; p.K = K;
  SMK_PROPAGATE(set_terms(p, tr, false));
  p.bias = nullptr; p.C = C; p.ldc = rows_w; p.epi = SMK_EPI_NONE; p.out_f32 = 1; p.rev = 0; p.trans = 0; p.credit_k = K;
  p.credit_flops = 2.0 * n_batch * rows_a * rows_w * K;
  p.mb_per_batch = mb; p.batch_a_rows = batch_a_rows; p.a_row0 = a_row0; p.batch_b_rows = batch_w_rows; p.b_row0 = w_row0;
  CUtensorMap ta, tb, tcm;
  SMK_PROPAGATE(make_tmap_bf16_2d(&ta, A, (uint64_t)a_cols, (uint64_t)a_total_rows, (uint64_t)lda * 2, TC_BK, TC_BM));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tb, W, (uint64_t)w_cols, (uint64_t)w_total_rows, (uint64_t)ldw * 2, TC_BK, BN));
  const uint64_t dims[3] = {(uint64_t)rows_w, (uint64_t)rows_a, (uint64_t)n_batch};
  const uint64_t strides[2] = {(uint64_t)rows_w * 4, (uint64_t)rows_a * rows_w * 4};
  const uint32_t box[3] = {32u, 32u, 1u};
  SMK_PROPAGATE(make_tmap_nd(&tcm, 4, C, 3, dims, strides, box, 128));
  return launch_tc<BN, false, 1, 8, kF16>(ta, tb, tcm, p, s);
}
int gemm_tc_batched(const void* A, int64_t lda, int64_t a_total_rows, int batch_a_rows, int a_row0, int rows_a, const void* W, int64_t ldw,
                    int64_t w_total_rows, int batch_w_rows, int w_row0, int rows_w, float* C, int n_batch, int K, int f16, const GemmTerms& terms,
                    cudaStream_t s) {
  return f16 ? gemm_tc_batched_impl<true>(A, lda, a_total_rows, batch_a_rows, a_row0, rows_a, W, ldw, w_total_rows, batch_w_rows, w_row0, rows_w, C,
                                          n_batch, K, terms, s)
             : gemm_tc_batched_impl<false>(A, lda, a_total_rows, batch_a_rows, a_row0, rows_a, W, ldw, w_total_rows, batch_w_rows, w_row0, rows_w, C,
                                           n_batch, K, terms, s);
}

int gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K, int epi,
            int out_f32, int tok_hw, const float* tok_pos, int f16, const GemmTerms& terms, int credit_k, cudaStream_t s) {
  return f16 ? gemm_tc_impl<true>(A, lda, W, ldw, bias, C, ldc, M, N, K, epi, out_f32, tok_hw, tok_pos, terms, credit_k, s)
             : gemm_tc_impl<false>(A, lda, W, ldw, bias, C, ldc, M, N, K, epi, out_f32, tok_hw, tok_pos, terms, credit_k, s);
}

// plain bf16 form: A [M,K] bf16 (lda elements), W [N,K] bf16 (ldw elements).  credit_k: mathematical reduction length when K is
// the K' = 3K of a [hi | hi | lo] x [hi | lo | hi] split pair (0 = K)
int gemm_bf16_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias, void* C, int64_t ldc,
                 int M, int N, int K, int epi, int out_f32, int tok_hw, const float* tok_pos, cudaStream_t s, int credit_k) {
  SMK_REQUIRE(out_f32 >= 0 && out_f32 <= 2, "gemm_bf16: out_f32 must be 0 (bf16), 1 (fp32) or 2 (bf16x3 split)");
  // K' = 3K of a [hi | hi | lo] x [hi | lo | hi] split pair (the caller names the mathematical K): the same three products as fused
  // terms over the distinct halves — A columns {0, 2K}, W columns {0, K} — so that a k-block loads 4 operand tiles instead of 6
  if (credit_k > 0 && K == 3 * credit_k && credit_k % TC_BK == 0)
    return gemm_tc_impl<false>(A, lda, W, ldw, bias, C, ldc, M, N, credit_k, epi, out_f32, tok_hw, tok_pos, terms_legacy3(credit_k), credit_k, s);
  return gemm_tc_impl<false>(A, lda, W, ldw, bias, C, ldc, M, N, K, epi, out_f32, tok_hw, tok_pos, GemmTerms{1, {0, 0, 0}, {0, 0, 0}}, credit_k, s);
}

}  // namespace smk

extern "C" int smk_debug_gemm_trace(long long* buf) {   // tuning aid: >= 16 * grid counters; nullptr switches the trace off
  SMK_CHECK_CUDA(cudaMemcpyToSymbol(smk::g_gemm_trace, &buf, sizeof(buf)));
  return SMK_OK;
}

extern "C" int smk_gemm_bf16(const void* A, int64_t lda, const void* W, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                             int epilogue, int out_f32, void* stream) {
  SMK_REQUIRE(A && W && C && M >= 0 && N > 0 && K > 0, "smk_gemm_bf16: bad arguments");
  return smk::gemm_bf16_tc((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, K, bias, C, ldc, M, N, K, epilogue, out_f32, 0, nullptr,
                           (cudaStream_t)stream);
}

extern "C" int smk_gemm_batched(const void* A, int64_t lda, int64_t a_total_rows, int batch_a_rows, int a_row0, int rows_a, const void* W,
                                int64_t ldw, int64_t w_total_rows, int batch_w_rows, int w_row0, int rows_w, float* C, int n_batch, int K,
                                int f16, int n_terms, const int32_t* a_off, const int32_t* w_off, void* stream) {
  SMK_REQUIRE(A && W && C && a_off && w_off && n_terms >= 1 && n_terms <= 3, "smk_gemm_batched: bad arguments");
  smk::GemmTerms t{n_terms, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n_terms; ++i) { t.a_off[i] = a_off[i]; t.w_off[i] = w_off[i]; }
  return smk::gemm_tc_batched(A, lda, a_total_rows, batch_a_rows, a_row0, rows_a, W, ldw, w_total_rows, batch_w_rows, w_row0, rows_w, C, n_batch, K,
                              f16, t, (cudaStream_t)stream);
}

extern "C" int smk_gemm_split(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N,
                              int K, int epilogue, int out_kind, int f16, int n_terms, const int32_t* a_off, const int32_t* w_off,
                              void* stream) {
  SMK_REQUIRE(A && W && C && M >= 0 && N > 0 && K > 0 && a_off && w_off && n_terms >= 1 && n_terms <= 3, "smk_gemm_split: bad arguments");
  smk::GemmTerms t{n_terms, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n_terms; ++i) { t.a_off[i] = a_off[i]; t.w_off[i] = w_off[i]; }
  return smk::gemm_tc(A, lda, W, ldw, bias, C, ldc, M, N, K, epilogue, out_kind, 0, nullptr, f16, t, 0, (cudaStream_t)stream);
}

// Patch-embed form (vision_transformer.py:184-188 + prepare_tokens :269-287): C viewed as [n_img, hw + 1, N] token rows; GEMM row m = patch
// m % hw of image m / hw lands on token row 1 + m % hw with pos[(1 + m % hw), :] added (row 0, the class token, is not written).
// A [n_img*hw, lda], W [N, ldw]: bf16, fp16 (f16 = 1) or the q8 split rows (f16 = 2).
extern "C" int smk_gemm_tokens(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* pos, float* C, int64_t ldc,
                               int n_img, int hw, int N, int K, int f16, void* stream) {
  SMK_REQUIRE(A && W && C && pos && n_img >= 0 && hw > 0 && N > 0 && K > 0, "smk_gemm_tokens: bad arguments");
  return smk::gemm_tc(A, lda, W, ldw, bias, C, ldc, n_img * hw, N, K, SMK_EPI_NONE, 1, hw, pos, f16 != 0, f16 == 2 ? smk::terms_q8(K) : smk::terms_plain(), 0,
                      (cudaStream_t)stream);
}

extern "C" int smk_gemm_q8(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                           int epilogue, int out_kind, void* stream) {
  SMK_REQUIRE(A && W && C && M >= 0 && N > 0 && K > 0, "smk_gemm_q8: bad arguments");
  return smk::gemm_tc(A, lda, W, ldw, bias, C, ldc, M, N, K, epilogue, out_kind, 0, nullptr, 1, smk::terms_q8(K), 0, (cudaStream_t)stream);
}
