// Mask head of the tensor-core modes, in two kernels (maskformer.py:144-162, :223):
//   1. mask logits at patch resolution: queries · memory^T per image as ONE batched tcgen05 GEMM (smk_gemm_tc.cu, gemm_tc_batched):
//      per image a 128-row tile holds the L·nq query rows of all decoder layers (120 of 128 rows at L = 6, nq = 20), N = the image's
//      patch tokens, K = D with the 3-term bf16 split (hi·hi + hi·lo + lo·hi, fp32 accumulate in TMEM ≈ fp32 product): the query rows
//      arrive already split [hi | hi | lo] from the decoder's last LayerNorm, the memory tokens [hi | lo] from the encoder's final
//      LayerNorm; operands by TMA, the logits leave through a 3-D TMA store that clips the pad rows / columns.
//      (Round 1 ran this on mma.sync: 89 us per step at batch 256 against ~20 us here.)
//   2. mask_upsample_x4_kernel: the pixel decoder's bilinear x4 applied to the nq-channel logits (bilinear is linear and
//      per-channel, so it commutes with the contraction — SURVEY.md K12), sigmoid, 16-byte streaming stores of mask_pred.
// The fused CUDA-core kernel (smk_simt.cu mask_head_kernel) stays for the fp32 / bf16x3 modes and other scale factors.
#include "smk_kernels.h"

namespace smk {

namespace {

// planes [n_planes][hp][wp] (logits) → out [n_planes][4hp][4wp] = sigmoid(bilinear x4); one warp per plane, a lane owns a source
// column i (output pixels 4i .. 4i+3, one 16-byte store per row) and every kRG-th row of source cells: the horizontal blends are
// shared by the <= 4 output rows of a cell row (same arithmetic as quad4 / bilerp, so identical values to the fused kernel).
constexpr int MU_WARPS = 8;
__global__ void __launch_bounds__(MU_WARPS * 32)
mask_upsample_x4_kernel(const float* __restrict__ planes, int64_t plane_stride, float* __restrict__ out, float* __restrict__ logits_out,
                        int64_t n_planes, int hp, int wp) {
  extern __shared__ float mu_smem[];
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pi = (int64_t)blockIdx.x * MU_WARPS + warp;
  if (pi >= n_planes) return;
  const int hw = hp * wp, Ho = 4 * hp, Wo = 4 * wp;
  float* pl = mu_smem + warp * hw;
  const float* src = planes + pi * plane_stride;
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

// queries (split rows [hi | hi | lo] of image b: rows b·q_batch_rows + layer0·nq .. + L·nq of q3) x memory tokens ([hi | lo] rows of
// tok16, image b at rows b·(hw+1) + 1) → logits [B, L*nq, ld_logits] → mask_pred [B, L, nq, 4hp, 4wp]
int mask_head_tc(const __nv_bfloat16* q3, int q_batch_rows, const __nv_bfloat16* tok16, float* logits_lowres, float* mask_pred,
                 float* logits_out, int B, int L, int layer0, int nq, int D, int hp, int wp, cudaStream_t s) {
  if (B == 0) return SMK_OK;
  const int hw = hp * wp, R = L * nq, N = hw + 1;
  const int ld = (hw + 3) / 4 * 4;                        // TMA store: 16-byte aligned logit rows
  SMK_REQUIRE(D % 64 == 0 && B <= 65535, "mask_head_tc: D=%d / B=%d unsupported", D, B);
  {
    TagScope tg(TAG_MASK_LOGITS);
    const GemmTerms t3{3, {0, 0, 2 * D}, {0, D, 0}};      // q_hi·t_hi + q_hi·t_lo + q_lo·t_hi (the duplicate hi columns are not read)
    SMK_PROPAGATE(gemm_tc_batched(q3, 3 * (int64_t)D, (int64_t)B * q_batch_rows, q_batch_rows, layer0 * nq, R, tok16, 2 * (int64_t)D, (int64_t)B * N, N, 1, ld,
                                  logits_lowres, B, D, 0, t3, s));
  }
  const int64_t n_planes = (int64_t)B * R;
  const size_t smem2 = (size_t)MU_WARPS * hw * sizeof(float);
  SMK_REQUIRE(smem2 <= 200 * 1024, "mask_head_tc: %d x %d planes do not fit shared memory", hp, wp);
  if (smem2 > 48 * 1024) SMK_CHECK_CUDA(cudaFuncSetAttribute(mask_upsample_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  {
    TagScope tg(TAG_MASK_UPSAMPLE);
    ProfScope prof(PROF_MASK_HEAD, (double)n_planes * ((double)hw + 16.0 * hw) * 4.0, s);
    SMK_CHECK_CUDA(launch_pdl(mask_upsample_x4_kernel, dim3((unsigned)((n_planes + MU_WARPS - 1) / MU_WARPS)), dim3(MU_WARPS * 32), smem2, s,
                              (const float*)logits_lowres, (int64_t)ld, mask_pred, logits_out, n_planes, hp, wp));
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

}  // namespace smk
