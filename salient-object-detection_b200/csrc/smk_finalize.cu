// Device-side finalisation of the per-mask records: integer counts / histograms / moment sums → the reference's metric
// VALUES, with every floating-point operation individually rounded in the reference's order (no FMA contraction):
//   iou.py:31, pixel_acc.py:10-14, f_measure.py:24-81 (255 thresholds, beta_square**2 quirk), mae.py:9 in float32;
//   s_measure.py:108-124 in float64 (the reference returns a python float).
// It restates salient-object-detection_b200/metrics.py::finalize (numpy) bit for bit — tests/test_gpu_kernels.py checks
// equality — so that the host side of an evaluation sweep is reduced to the ordered running means of average_meter.py.
// One CTA per evaluated mask: thread k < 255 owns threshold k of F-max, thread 0 the scalar metrics.
#include "smk_common.cuh"

namespace smk {

namespace {

struct FinalizeConst { float c1, c2, eps; };   // float32(1 + beta_square**2), float32(beta_square**2), float32(1e-7)

__device__ __forceinline__ float f_from_counts(int tp, int tp_fp, int tp_fn, const FinalizeConst& fc) {
  const float t = __int2float_rn(tp);
  const float prec = __fdiv_rn(t, __fadd_rn(__int2float_rn(tp_fp), fc.eps));
  const float rec = __fdiv_rn(t, __fadd_rn(__int2float_rn(tp_fn), fc.eps));
  const float num = __fmul_rn(__fmul_rn(fc.c1, prec), rec);
  const float den = __fadd_rn(__fadd_rn(__fmul_rn(fc.c2, prec), rec), fc.eps);
  return __fdiv_rn(num, den);
}

// s_measure.py:54-60 with the unbiased std of torch (.std(), correction = 1)
__device__ __forceinline__ double s_object(double sum1, double sum2, double cnt) {
  const double mu = __ddiv_rn(sum1, cnt);
  const double var = cnt > 1.0 ? __ddiv_rn(__dsub_rn(sum2, __dmul_rn(__dmul_rn(cnt, mu), mu)), __dsub_rn(cnt, 1.0)) : nan("");
  const double sd = isnan(var) ? var : __dsqrt_rn(var > 0.0 ? var : 0.0);
  return __ddiv_rn(__dmul_rn(2.0, mu), __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(mu, mu), 1.0), sd), 1e-20));
}

// s_measure.py:33-52 from the quadrant's moment sums {N, Σp, Σp², Σg, Σpg}
__device__ __forceinline__ double s_ssim(const double* q) {
  const double N = q[0], sp = q[1], sp2 = q[2], sg = q[3], spg = q[4];
  if (N == 0.0) return nan("");
  const double x = __ddiv_rn(sp, N), y = __ddiv_rn(sg, N);
  const double den = __dadd_rn(__dsub_rn(N, 1.0), 1e-20);
  const double sx2 = __ddiv_rn(__dsub_rn(sp2, __dmul_rn(__dmul_rn(N, x), x)), den);
  const double sy2 = __ddiv_rn(__dsub_rn(sg, __dmul_rn(__dmul_rn(N, y), y)), den);
  const double sxy = __ddiv_rn(__dsub_rn(spg, __dmul_rn(__dmul_rn(N, x), y)), den);
  double a = __dmul_rn(__dmul_rn(__dmul_rn(4.0, x), y), sxy);
  const double b = __dmul_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dadd_rn(sx2, sy2));
  if ((y == 0.0 || y == 1.0) && !isnan(a)) a = 0.0;   // (g - ȳ) ≡ 0 → the reference's σxy is an exact zero
  if (a != 0.0) return __ddiv_rn(a, __dadd_rn(b, 1e-20));
  return b == 0.0 ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256)
finalize_kernel(const int32_t* __restrict__ m_counts, const double* __restrict__ m_sums, double* __restrict__ out, FinalizeConst fc) {
  __shared__ int s_fg[256], s_all[256];
  __shared__ float s_max[8];
  pdl_wait();
  pdl_trigger();
  const int64_t mask = blockIdx.x;
  const int32_t* c = m_counts + mask * SMK_MCOUNT_STRIDE;
  const double* s = m_sums + mask * SMK_MSUM_STRIDE;
  const int k = threadIdx.x;
  s_fg[k] = c[k];
  s_all[k] = c[k] + c[256 + k];
  __syncthreads();
  const int G = c[514];
  // F-max: count(p > t_k) = Σ_{b>k} hist[b], k = 0..254 (f_measure.py:52-69)
  float fk = -INFINITY;
  if (k < 255) {
    int fg = 0, all = 0;
    for (int b = k + 1; b < 256; ++b) { fg += s_fg[b]; all += s_all[b]; }
    fk = f_from_counts(fg, all, G, fc);
  }
  fk = warp_max(fk);
  if ((k & 31) == 0) s_max[k >> 5] = fk;
  __syncthreads();
  if (k != 0) return;
  float fmax = s_max[0];
  for (int w = 1; w < 8; ++w) fmax = fmaxf(fmax, s_max[w]);

  const int tp05 = c[512], tpfp05 = c[513], tpm = c[515], tpfpm = c[516], n_i = c[519];
  const int uni = tpfp05 + G - tp05, wrong = tpfp05 + G - 2 * tp05;
  double* o = out + mask * 8;
  o[0] = (double)__fdiv_rn(__int2float_rn(tp05), __fadd_rn(__int2float_rn(uni), fc.eps));     // iou
  o[1] = (double)__fdiv_rn(__int2float_rn(n_i - wrong), __int2float_rn(n_i));                    // pixel accuracy
  o[2] = (double)f_from_counts(tp05, tpfp05, G, fc);                                             // F @ 0.5
  o[3] = (double)fmax;
  o[4] = (double)f_from_counts(tpm, tpfpm, G, fc);                                               // F @ 2·mean(p)
  const double n = (double)n_i, Gd = (double)G;
  o[5] = (double)__double2float_rn(__ddiv_rn(s[1], n));                                          // MAE
  // S-measure
  const double mean_p = __ddiv_rn(s[0], n);
  double sm;
  if (G == 0) sm = __dsub_rn(1.0, mean_p);
  else if (G == n_i) sm = mean_p;
  else {
    const double u = __ddiv_rn(Gd, n);
    const double s_obj = __dadd_rn(__dmul_rn(u, s_object(s[3], s[4], Gd)), __dmul_rn(__dsub_rn(1.0, u), s_object(s[5], s[6], __dsub_rn(n, Gd))));
    const float hw = __int2float_rn(n_i);
    const float w1 = __fdiv_rn(__fmul_rn(__int2float_rn(c[517]), __int2float_rn(c[518])), hw);
    const float w2 = __fdiv_rn(__double2float_rn(s[8 + 5]), hw);
    const float w3 = __fdiv_rn(__double2float_rn(s[8 + 10]), hw);
    const float w4 = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, w1), w2), w3);
    const double q0 = s_ssim(s + 8), q1 = s_ssim(s + 13), q2 = s_ssim(s + 18), q3 = s_ssim(s + 23);
    const double s_reg = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn((double)w1, q0), __dmul_rn((double)w2, q1)), __dmul_rn((double)w3, q2)),
                                   __dmul_rn((double)w4, q3));
    const double val = __dadd_rn(__dmul_rn(0.5, s_obj), __dmul_rn(0.5, s_reg));
    sm = val < 0.0 ? 0.0 : val;
  }
  o[6] = sm;
  o[7] = 0.0;
}

}  // namespace
}  // namespace smk

extern "C" int smk_finalize_records(const int32_t* m_counts, const double* m_sums, int64_t n_masks, float c1, float c2, float eps,
                                    double* out, void* stream) {
  SMK_REQUIRE(m_counts && m_sums && out, "smk_finalize_records: null pointer");
  SMK_REQUIRE(n_masks >= 0 && n_masks < (1ll << 31), "smk_finalize_records: bad mask count");
  if (n_masks == 0) return SMK_OK;
  smk::FinalizeConst fc{c1, c2, eps};
  SMK_CHECK_CUDA(smk::launch_pdl(smk::finalize_kernel, dim3((unsigned)n_masks), dim3(256), 0, (cudaStream_t)stream, m_counts, m_sums, out, fc));
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}
