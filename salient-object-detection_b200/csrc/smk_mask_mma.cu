// Mask head for the bf16 tensor-core mode, in two kernels (maskformer.py:144-162, :223):
//   1. mask_logits_mma_kernel: logits at patch resolution, queries · memory^T per image, on tensor cores (mma.sync m16n8k16) with
//      the 3-term bf16 split (hi·hi + lo·hi + hi·lo, fp32 accumulate ≈ fp32 product): the query rows arrive already split from
//      the decoder's last LayerNorm, the memory tokens from the encoder's final LayerNorm (hi = the K/V GEMM operand, lo = residue).
//   2. mask_upsample_x4_kernel: the pixel decoder's bilinear x4 applied to the nq-channel logits (bilinear is linear and
//      per-channel, so it commutes with the contraction — SURVEY.md K12), sigmoid, 16-byte streaming stores of mask_pred.
// Splitting the fused CUDA-core kernel (smk_simt.cu mask_head_kernel, kept for the fp32 / bf16x3 modes and other scale factors)
// takes the contraction off the FMA pipe (4.6 GFLOP of fp32 FMAs per step at batch 256) and lets the write-bound part run as a
// flat, fully occupied map over planes.
#include "smk_mma.cuh"

namespace smk {

namespace {

using namespace mma;

constexpr int ML_BM = 64, ML_NT = 13, ML_BN = 2 * ML_NT * 8, ML_KC = 32, ML_LD = ML_KC + 8, ML_THREADS = 256;
constexpr int ML_STAGE = (2 * ML_BM + 2 * ML_BN) * ML_LD;     // bf16 elements per stage: Q hi, Q lo, T hi, T lo

struct MaskLogitsParams {
  const __nv_bfloat16* q3;      // [L_all, Rall, 3D] rows [hi | hi | lo]
  const __nv_bfloat16* tok_hi;  // [B*N, D]
  const __nv_bfloat16* tok_lo;  // [B*N, D]
  float* out;                   // [B, R, hw]
  int64_t Rall;
  int layer0, R, nq, hw, N, D, m_blocks;
};

__global__ void __launch_bounds__(ML_THREADS, 2)
mask_logits_mma_kernel(const MaskLogitsParams p) {
  extern __shared__ __align__(16) uint8_t ml_smem[];
  __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(ml_smem);
  pdl_wait();
  pdl_trigger();
  const int mb = blockIdx.x % p.m_blocks, nb = blockIdx.x / p.m_blocks, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int row0 = mb * ML_BM, n0 = nb * ML_BN;
  const int D = p.D;

  auto stage = [&](int chunk, int buf) {
    __nv_bfloat16* s = sm + buf * ML_STAGE;
    const int k0 = chunk * ML_KC;
    // query rows: hi at column k, lo at column 2D + k of the split row
    for (int i = tid; i < 2 * ML_BM * (ML_KC / 8); i += ML_THREADS) {
      const int part = i / (ML_BM * (ML_KC / 8)), j = i - part * (ML_BM * (ML_KC / 8)), r = j / (ML_KC / 8), c = j % (ML_KC / 8);
      __nv_bfloat16* d = s + (part * ML_BM + r) * ML_LD + c * 8;
      const int rr = row0 + r;
      if (rr < p.R) {
        const int l = rr / p.nq, q = rr - l * p.nq;
        const __nv_bfloat16* src = p.q3 + (((int64_t)(p.layer0 + l) * p.Rall + (int64_t)b * p.nq + q) * 3 + (part ? 2 : 0)) * D + k0 + c * 8;
        cp_async16((uint32_t)__cvta_generic_to_shared(d), src);
      } else {
        *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
      }
    }
    __nv_bfloat16* st = s + 2 * ML_BM * ML_LD;
    for (int i = tid; i < 2 * ML_BN * (ML_KC / 8); i += ML_THREADS) {
      const int part = i / (ML_BN * (ML_KC / 8)), j = i - part * (ML_BN * (ML_KC / 8)), r = j / (ML_KC / 8), c = j % (ML_KC / 8);
      __nv_bfloat16* d = st + (part * ML_BN + r) * ML_LD + c * 8;
      const int n = n0 + r;
      if (n < p.hw) {
        const __nv_bfloat16* src = (part ? p.tok_lo : p.tok_hi) + ((int64_t)b * p.N + 1 + n) * D + k0 + c * 8;   // skip the cls token
        cp_async16((uint32_t)__cvta_generic_to_shared(d), src);
      } else {
        *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
      }
    }
  };

  const int mt = warp & 3, nh = warp >> 2;         // 16-row tile, half of the CTA's token range
  float acc[ML_NT][4];
#pragma unroll
  for (int n = 0; n < ML_NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;

  const int n_chunks = D / ML_KC;
  stage(0, 0);
  cp_async_commit();
  for (int c = 0; c < n_chunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < n_chunks) {
      stage(c + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* s = sm + buf * ML_STAGE;
    uint32_t ah[2][4], al[2][4];
    {
      const uint32_t base = (uint32_t)__cvta_generic_to_shared(s + (mt * 16 + (lane & 15)) * ML_LD + (lane >> 4) * 8);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        ldmatrix_x4(base + kk * 32, ah[kk][0], ah[kk][1], ah[kk][2], ah[kk][3]);
        ldmatrix_x4(base + ML_BM * ML_LD * 2 + kk * 32, al[kk][0], al[kk][1], al[kk][2], al[kk][3]);
      }
    }
    const __nv_bfloat16* th = s + 2 * ML_BM * ML_LD + (nh * ML_NT * 8) * ML_LD;
#pragma unroll
    for (int n = 0; n < ML_NT; ++n) {
      const uint32_t off = (uint32_t)__cvta_generic_to_shared(th + (8 * n + (lane & 7)) * ML_LD + (lane >> 3) * 8);
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(off, b0, b1, b2, b3);                       // hi: k-steps 0 and 1 of this n-tile
      mma_bf16(acc[n], ah[0][0], ah[0][1], ah[0][2], ah[0][3], b0, b1);
      mma_bf16(acc[n], ah[1][0], ah[1][1], ah[1][2], ah[1][3], b2, b3);
      mma_bf16(acc[n], al[0][0], al[0][1], al[0][2], al[0][3], b0, b1);
      mma_bf16(acc[n], al[1][0], al[1][1], al[1][2], al[1][3], b2, b3);
      ldmatrix_x4(off + ML_BN * ML_LD * 2, b0, b1, b2, b3);   // lo
      mma_bf16(acc[n], ah[0][0], ah[0][1], ah[0][2], ah[0][3], b0, b1);
      mma_bf16(acc[n], ah[1][0], ah[1][1], ah[1][2], ah[1][3], b2, b3);
    }
    __syncthreads();          // every warp is done with this stage before the next prefetch overwrites it
  }

#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int rr = row0 + mt * 16 + g + 8 * half;
    if (rr >= p.R) continue;
    float* orow = p.out + ((int64_t)b * p.R + rr) * p.hw;
#pragma unroll
    for (int n = 0; n < ML_NT; ++n) {
      const int col = n0 + (nh * ML_NT + n) * 8 + 2 * t;
      if (col < p.hw) orow[col] = acc[n][2 * half];
      if (col + 1 < p.hw) orow[col + 1] = acc[n][2 * half + 1];
    }
  }
}

// planes [n_planes][hp][wp] (logits) → out [n_planes][4hp][4wp] = sigmoid(bilinear x4); one warp per plane, a lane owns a source
// column i (output pixels 4i .. 4i+3, one 16-byte store per row) and every kRG-th row of source cells: the horizontal blends are
// shared by the <= 4 output rows of a cell row (same arithmetic as quad4 / bilerp, so identical values to the fused kernel).
constexpr int MU_WARPS = 8;
__global__ void __launch_bounds__(MU_WARPS * 32)
mask_upsample_x4_kernel(const float* __restrict__ planes, float* __restrict__ out, float* __restrict__ logits_out, int64_t n_planes, int hp, int wp) {
  extern __shared__ float mu_smem[];
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pi = (int64_t)blockIdx.x * MU_WARPS + warp;
  if (pi >= n_planes) return;
  const int hw = hp * wp, Ho = 4 * hp, Wo = 4 * wp;
  float* pl = mu_smem + warp * hw;
  const float* src = planes + pi * hw;
  for (int i = lane; i < hw; i += 32) pl[i] = __ldg(src + i);
  __syncwarp();
  float* o = out + pi * (int64_t)Ho * Wo;
  float* lo = logits_out ? logits_out + pi * (int64_t)Ho * Wo : nullptr;
  const int lanes_per_row = min(wp, 32), rg = 32 / lanes_per_row;      // row groups handled side by side
  const int li = lane % lanes_per_row, lr = lane / lanes_per_row;
  if (lr >= rg) return;
  for (int i = li; i < wp; i += lanes_per_row) {
    const QuadX qx = make_quadx(i, wp);
    for (int v = lr; v <= hp; v += rg) {
      const int ya = v == 0 ? 0 : 4 * v - 2, yb = v == 0 ? 2 : min(4 * v + 2, Ho);
      const float* r0 = pl + max(v - 1, 0) * wp;
      const float* r1 = pl + min(v, hp - 1) * wp;
      const float am = r0[qx.cm], ac = r0[qx.cc], ap = r0[qx.cp], bm = r1[qx.cm], bc = r1[qx.cc], bp = r1[qx.cp];
      float top[4], bot[4];
      top[0] = __fmaf_rn(am, qx.l0[0], __fmul_rn(ac, qx.l1[0])); bot[0] = __fmaf_rn(bm, qx.l0[0], __fmul_rn(bc, qx.l1[0]));
      top[1] = __fmaf_rn(am, qx.l0[1], __fmul_rn(ac, qx.l1[1])); bot[1] = __fmaf_rn(bm, qx.l0[1], __fmul_rn(bc, qx.l1[1]));
      top[2] = __fmaf_rn(ac, qx.l0[2], __fmul_rn(ap, qx.l1[2])); bot[2] = __fmaf_rn(bc, qx.l0[2], __fmul_rn(bp, qx.l1[2]));
      top[3] = __fmaf_rn(ac, qx.l0[3], __fmul_rn(ap, qx.l1[3])); bot[3] = __fmaf_rn(bc, qx.l0[3], __fmul_rn(bp, qx.l1[3]));
      for (int y = ya; y < yb; ++y) {
        const Tap ty = make_tap(y, 0.25f, hp);
        float z[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) z[f] = __fmaf_rn(top[f], ty.l0, __fmul_rn(bot[f], ty.l1));
        __stcs(reinterpret_cast<float4*>(o + (int64_t)y * Wo + 4 * i),
               make_float4(sigmoid_fast(z[0]), sigmoid_fast(z[1]), sigmoid_fast(z[2]), sigmoid_fast(z[3])));
        if (lo) *reinterpret_cast<float4*>(lo + (int64_t)y * Wo + 4 * i) = make_float4(z[0], z[1], z[2], z[3]);
      }
    }
  }
}

}  // namespace

// queries (split rows of layers layer0 .. layer0 + L - 1) x memory tokens → logits [B, L*nq, hw] → mask_pred [B, L, nq, 4hp, 4wp]
int mask_head_mma(const __nv_bfloat16* q3, int64_t Rall, const __nv_bfloat16* tok_hi, const __nv_bfloat16* tok_lo, float* logits_lowres,
                  float* mask_pred, float* logits_out, int B, int L, int layer0, int nq, int D, int hp, int wp, cudaStream_t s) {
  if (B == 0) return SMK_OK;
  const int hw = hp * wp, R = L * nq;
  SMK_REQUIRE(D % ML_KC == 0 && B <= 65535, "mask_head_mma: D=%d / B=%d unsupported", D, B);
  const size_t smem1 = 2 * (size_t)ML_STAGE * sizeof(__nv_bfloat16);
  static DeviceOnce attr1;
  if (attr1.first()) SMK_CHECK_CUDA(cudaFuncSetAttribute(mask_logits_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  const int m_blocks = (R + ML_BM - 1) / ML_BM, n_blocks = (hw + ML_BN - 1) / ML_BN;
  MaskLogitsParams p{q3, tok_hi, tok_lo, logits_lowres, Rall, layer0, R, nq, hw, hw + 1, D, m_blocks};
  {
    ProfScope prof(PROF_MASK_HEAD, (double)B * ((double)(hw + R) * D + (double)R * hw) * 4.0, s);
    SMK_CHECK_CUDA(launch_pdl(mask_logits_mma_kernel, dim3((unsigned)(m_blocks * n_blocks), (unsigned)B), dim3(ML_THREADS), smem1, s, p));
  }
  SMK_CHECK_LAUNCH();
  const int64_t n_planes = (int64_t)B * R;
  const size_t smem2 = (size_t)MU_WARPS * hw * sizeof(float);
  SMK_REQUIRE(smem2 <= 200 * 1024, "mask_head_mma: %d x %d planes do not fit shared memory", hp, wp);
  if (smem2 > 48 * 1024) SMK_CHECK_CUDA(cudaFuncSetAttribute(mask_upsample_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  {
    ProfScope prof(PROF_MASK_HEAD, (double)n_planes * ((double)hw + 16.0 * hw) * 4.0, s);
    SMK_CHECK_CUDA(launch_pdl(mask_upsample_x4_kernel, dim3((unsigned)((n_planes + MU_WARPS - 1) / MU_WARPS)), dim3(MU_WARPS * 32), smem2, s,
                              (const float*)logits_lowres, mask_pred, logits_out, n_planes, hp, wp));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace smk
