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
  int Lq, Lk, nk_pad;           // valid query rows / keys per image, keys padded to a multiple of 16
  int q_rows, kv_rows, kv_row0; // image b: queries start at row b*q_rows, keys/values at row b*kv_rows + kv_row0
  void* out;                    // [B*q_rows, ldo]; bf16 or fp32
  int64_t ldo;
  int out_f32;
  float scale_log2e;
};

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const AttnTcParams p) {
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
  const int q_row0 = b * p.q_rows, kv_row0 = b * p.kv_rows + p.kv_row0;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
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
      tma_load_2d(sQ, &tmQ, bar_qk, h * AT_DH, q_row0 + qt * AT_BM);
      tma_load_2d(sK, &tmK, bar_qk, h * AT_DH, kv_row0);
      mbar_arrive_expect_tx(bar_v, kv_bytes);
      tma_load_2d(sV, &tmV, bar_v, h * AT_DH, kv_row0);
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
    // warps whose 32 rows are all beyond Lq do no softmax work (their P rows keep stale-but-finite Q/K bits; MMA rows
    // are independent and those output rows are never stored)
    const int n_chunks = (qt * AT_BM + quarter * 32 < p.Lq) ? (p.nk_pad + 31) / 32 : 0;
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
        if (c * 32 + j < p.Lk) mx = fmaxf(mx, __uint_as_float(v[j]));
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
        e[j] = (c * 32 + j < p.Lk) ? pv : 0.f;
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
    const float inv = 1.0f / sum;   // inf for skipped rows (never stored)
    const int64_t orow = (int64_t)(q_row0 + row) * p.ldo + h * AT_DH;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      if (row < p.Lq) {
        if (p.out_f32) {
          float* o = reinterpret_cast<float*>(p.out) + orow + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv,
                                                            __uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
            pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
            pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(o + j) = pk;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, AT_TMEM_COLS);
}

int attention_tc_general(const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v, int64_t ldv,
                         int64_t kv_total_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_f32, int B, int Lq, int Lk,
                         int heads, float scale, cudaStream_t s) {
  const int D = heads * AT_DH;
  SMK_REQUIRE(Lk >= 1 && Lk <= 256 && Lq >= 1, "attention_tc: Lk=%d keys not supported (1..256)", Lk);
  SMK_REQUIRE(B >= 1 && B <= 65535 && heads >= 1 && heads <= 65535, "attention_tc: bad batch/heads");
  SMK_REQUIRE(ldo % 8 == 0 && ((uintptr_t)out % 16) == 0, "attention_tc: output must be 16-byte aligned");
  const int nk_pad = (Lk + 15) / 16 * 16;
  CUtensorMap tq, tk, tv;
  SMK_PROPAGATE(make_tmap_bf16_2d(&tq, q, (uint64_t)D, (uint64_t)B * Lq, (uint64_t)ldq * 2, AT_DH, AT_BM));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tk, k, (uint64_t)D, (uint64_t)kv_total_rows, (uint64_t)ldk * 2, AT_DH, (uint32_t)nk_pad));
  SMK_PROPAGATE(make_tmap_bf16_2d(&tv, v, (uint64_t)D, (uint64_t)kv_total_rows, (uint64_t)ldv * 2, AT_DH, (uint32_t)nk_pad));
  static bool attr_set = false;
  if (!attr_set) {
    SMK_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr_set = true;
  }
  AttnTcParams p{Lq, Lk, nk_pad, Lq, kv_rows, kv_row0, out, ldo, out_f32, scale * 1.4426950408889634f};
  dim3 grid((Lq + AT_BM - 1) / AT_BM, heads, B);
  {
    ProfScope prof(PROF_ATTENTION_TC, 4.0 * Lq * Lk * AT_DH * heads * B, s);
    attn_tc_kernel<<<grid, AT_THREADS, AT_SMEM, s>>>(tq, tk, tv, p);
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

// encoder form — qkv: [B*N, 3*D] bf16 (q | k | v, head h in columns [h*64, h*64+64) of each third); out: [B*N, D] bf16
int attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int heads, float scale, cudaStream_t s) {
  const int D = heads * AT_DH;
  return attention_tc_general(qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, (int64_t)B * N, N, 0, out, D, 0, B, N, N, heads, scale, s);
}

}  // namespace smk

extern "C" int smk_attention_tc(const void* qkv, void* out, int B, int N, int heads, float scale, void* stream) {
  SMK_REQUIRE(qkv && out, "smk_attention_tc: null pointer");
  return smk::attention_tc((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, B, N, heads, scale, (cudaStream_t)stream);
}

extern "C" int smk_attention_tc_general(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                        int64_t kv_total_rows, int kv_rows, int kv_row0, void* out, int64_t ldo, int out_f32, int B, int Lq,
                                        int Lk, int heads, float scale, void* stream) {
  SMK_REQUIRE(q && k && v && out, "smk_attention_tc_general: null pointer");
  return smk::attention_tc_general((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, kv_total_rows,
                                   kv_rows, kv_row0, out, ldo, out_f32, B, Lq, Lk, heads, scale, (cudaStream_t)stream);
}
