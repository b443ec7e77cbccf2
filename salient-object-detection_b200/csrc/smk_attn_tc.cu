// Fused encoder self-attention on tcgen05 (vision_transformer.py:110-130): per (image, head, 128-query tile)
//   S = Q·K^T  (UMMA 128 x Nk x 64, accumulators in TMEM)  →  single-pass softmax in registers (one thread per
//   query row, tcgen05.ld)  →  P (bf16) written to shared memory in the 128B-swizzled K-major layout  →
//   O = P·V   (V consumed as an MN-major operand straight from its TMA tile)  →  O / rowsum → bf16 → global.
// Q, K, V are read in place from the fused-QKV GEMM output [B*N, 3*D] by TMA (no permute copy, :113-118).
// One key tile: N <= 256 tokens (224x224 / patch 16 → 197).  Longer sequences use the CUDA-core kernel.
// 160 threads: warp 0 = TMA + MMA issue + TMEM owner, warps 1-4 = softmax / epilogue (TMEM lane quarters).
#include "smk_tc.cuh"

namespace smk {

using namespace tc;

constexpr int AT_BM = 128, AT_DH = 64, AT_THREADS = 160, AT_TMEM_COLS = 256;
constexpr int AT_REGION_A = 65536;   // Q (16 KB) + K (<= 32 KB), later reused for P (4 K-blocks x 16 KB)
constexpr int AT_REGION_V = 32768;
constexpr int AT_SMEM = AT_REGION_A + AT_REGION_V + 1024 /*align*/ + 128 /*barriers*/;

struct AttnTcParams {
  int N, nk_pad, D;
  __nv_bfloat16* out;
  int64_t ldo;
  float scale_log2e;
};

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sP = smem;                       // aliases Q/K once S has been produced
  uint8_t* sV = smem + AT_REGION_A;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_REGION_A + AT_REGION_V);
  uint64_t *bar_qk = bars, *bar_v = bars + 1, *bar_s = bars + 2, *bar_p = bars + 3, *bar_o = bars + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * p.N;   // first token row of this image in the [B*N, 3D] buffer

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, AT_TMEM_COLS);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t kv_bytes = (uint32_t)p.nk_pad * 128u;
      mbar_arrive_expect_tx(bar_qk, 16384u + kv_bytes);
      tma_load_2d(sQ, &tmQ, bar_qk, h * AT_DH, row0 + qt * AT_BM);
      tma_load_2d(sK, &tmKV, bar_qk, p.D + h * AT_DH, row0);
      mbar_arrive_expect_tx(bar_v, kv_bytes);
      tma_load_2d(sV, &tmKV, bar_v, 2 * p.D + h * AT_DH, row0);
      // S = Q · K^T
      mbar_wait(bar_qk, 0);
      tc_fence_after_sync();
      const uint32_t idesc_s = idesc_bf16_f32(AT_BM, p.nk_pad, 0, 0);
      const uint64_t qd = smem_desc_k_sw128(smem_u32(sQ)), kd = smem_desc_k_sw128(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < AT_DH / 16; ++k) umma_bf16_ss(tmem_base, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
      tc_commit(bar_s);
      // O = P · V   (accumulates into TMEM columns [0,64): S has been fully read out by then)
      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after_sync();
      const uint32_t idesc_o = idesc_bf16_f32(AT_BM, AT_DH, 0, 1);
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
      for (int j = 0; j < p.nk_pad / 16; ++j) {
        const uint64_t pd = smem_desc_k_sw128(pa + (uint32_t)((j >> 2) * 16384 + (j & 3) * 32));
        const uint64_t vd = smem_desc_mn_sw128(va + (uint32_t)(j * 2048), 1024);
        umma_bf16_ss(tmem_base, pd, vd, idesc_o, j != 0);
      }
      tc_commit(bar_o);
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;             // row within the tile = TMEM lane
    const int row = qt * AT_BM + r;                // token index within the image
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int n_chunks = (p.nk_pad + 31) / 32;
    mbar_wait(bar_s, 0);
    tc_fence_after_sync();
    // pass 1: row maximum over the valid keys
    float mx = -INFINITY;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < p.N) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxs = mx * p.scale_log2e;
    // pass 2: p = exp2(s*scale*log2e - max*scale*log2e); P → smem (bf16, K-major, 128B swizzle)
    float sum = 0.f;
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pv = exp2f(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mxs));
        e[j] = (c * 32 + j < p.N) ? pv : 0.f;
        sum += e[j];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {                // 4 x (8 keys = 16 bytes)
        const int key0 = c * 32 + g * 8;
        if (key0 < p.nk_pad) {
          __nv_bfloat162 t0 = __floats2bfloat162_rn(e[g * 8 + 0], e[g * 8 + 1]), t1 = __floats2bfloat162_rn(e[g * 8 + 2], e[g * 8 + 3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(e[g * 8 + 4], e[g * 8 + 5]), t3 = __floats2bfloat162_rn(e[g * 8 + 6], e[g * 8 + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
          const int kblk = key0 >> 6, chunk16 = (key0 & 63) >> 3;
          *reinterpret_cast<uint4*>(prow + kblk * 16384 + ((chunk16 ^ (r & 7)) << 4)) = pk;
        }
      }
    }
    fence_proxy_async();            // generic-proxy smem writes → visible to the tensor core (async proxy)
    tc_fence_before_sync();         // our TMEM reads of S are ordered before the PV MMAs that overwrite it
    mbar_arrive(bar_p);
    // epilogue: O / rowsum → bf16 → out[b*N + row, h*64 : h*64+64]
    mbar_wait(bar_o, 0);
    tc_fence_after_sync();
    const float inv = 1.0f / sum;
    __nv_bfloat16* orow = p.out + (int64_t)(row0 + row) * p.ldo + h * AT_DH;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      if (row < p.N) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
          pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(orow + c * 32 + j) = pk;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, AT_TMEM_COLS);
}

// qkv: [B*N, 3*D] bf16 (q | k | v, head h in columns [h*64, h*64+64) of each third); out: [B*N, D] bf16
int attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int heads, float scale, cudaStream_t s) {
  const int D = heads * AT_DH;
  SMK_REQUIRE(N >= 1 && N <= 256, "attention_tc: N=%d tokens not supported (1..256)", N);
  SMK_REQUIRE(B >= 1 && B <= 65535 && heads >= 1 && heads <= 65535, "attention_tc: bad batch/heads");
  const int nk_pad = (N + 15) / 16 * 16;
  CUtensorMap tq, tkv;
  SMK_PROPAGATE(make_tmap_bf16_2d(&tq, qkv, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D * 2, AT_DH, AT_BM));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tkv, qkv, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D * 2, AT_DH, (uint32_t)nk_pad));
  static bool attr_set = false;
  if (!attr_set) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr_set = true;
  }
  AttnTcParams p{N, nk_pad, D, out, (int64_t)D, scale * 1.4426950408889634f};
  dim3 grid((N + AT_BM - 1) / AT_BM, heads, B);
  {
    ProfScope prof(PROF_ATTENTION_TC, 4.0 * N * N * AT_DH * heads * B, s);
    attn_tc_kernel<<<grid, AT_THREADS, AT_SMEM, s>>>(tq, tkv, p);
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace smk

extern "C" int smk_attention_tc(const void* qkv, void* out, int B, int N, int heads, float scale, void* stream) {
  SMK_REQUIRE(qkv && out, "smk_attention_tc: null pointer");
  return smk::attention_tc((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, B, N, heads, scale, (cudaStream_t)stream);
}
