// Evaluation kernels: x4 bilinear upsample of probabilities fused with the per-query IoU counts, the
// upper-bound / objectness query selection and every reduction behind IoU, F-measure (F@0.5, F-max over
// 255 thresholds via 256-bin histograms, F-mean), MAE, pixel accuracy and S-measure.
//
// Reference semantics: evaluator.pyc@L199-226 and metrics/{iou,f_measure,mae,pixel_acc,s_measure}.py
// (SURVEY.md §8a').  All integer outputs are exact; the host forms the float32 ratios from them.
// HBM-bound byte work: the full-resolution masks [B,nq,H,W] are never written — each CTA keeps one
// low-resolution probability plane in shared memory and re-creates full-resolution pixels on the fly.
#include <stdlib.h>

#include "smk_common.cuh"

namespace smk {

// F-max thresholds t_k = float32(k * (1/255)), the product formed in double (f_measure.py:65 `torch.arange(0, 1, 1/255)`, ATen CPU
// arange accumulates in double): computed in the kernel — IEEE double multiply + round-to-nearest conversion, bit-identical to the
// host expression — so the library keeps no per-process device table (no allocation, no synchronisation, any device, capturable).
__device__ __forceinline__ float fmax_threshold(int k) { return __double2float_rn(__dmul_rn((double)k, 1.0 / 255.0)); }


constexpr int kEvalThreads = 256;
constexpr int kEvalWarps = kEvalThreads / 32;

__device__ __forceinline__ int block_sum_int(int v, int* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int i = 0; i < kEvalWarps; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ double block_sum_double(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < kEvalWarps; ++i) t += red[i];   // fixed order → run-to-run deterministic
  return t;
}

// Value of full-resolution pixel (y,x) of a plane held at low resolution in `pl` ([hp,wp]).
struct RowTap {
  const float* r0;
  const float* r1;
  float ly0, ly1;
};
__device__ __forceinline__ RowTap row_tap(const float* pl, int y, int hp, int wp, float rscale) {
  Tap t = make_tap(y, rscale, hp);
  return RowTap{pl + t.i0 * wp, pl + t.i1 * wp, t.l0, t.l1};
}
__device__ __forceinline__ float pixel(const RowTap& r, int x, int wp, float rscale) {
  Tap t = make_tap(x, rscale, wp);
  return bilerp(r.r0[t.i0], r.r0[t.i1], r.r1[t.i0], r.r1[t.i1], t.l0, t.l1, r.ly0, r.ly1);
}

// ------------------------------------------------------------------------------------------------
// E1: per (image, query) intersection / union at threshold 0.5  (evaluator.pyc@L113-120, iou.py:22-30)
// grid (nq, B), 256 threads.  dynamic smem: hp*wp floats.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEvalThreads)
query_iou_kernel(const float* __restrict__ mask_pred, int64_t batch_stride, const uint8_t* __restrict__ gt,
                 int nq, int hp, int wp, int up, int H, int W, int32_t* __restrict__ q_counts) {
  extern __shared__ float plane[];
  __shared__ int red[kEvalWarps];
  const int q = blockIdx.x, b = blockIdx.y;
  const float* src = mask_pred + (int64_t)b * batch_stride + (int64_t)q * hp * wp;
  for (int i = threadIdx.x; i < hp * wp; i += kEvalThreads) plane[i] = src[i];
  __syncthreads();
  const float rscale = 1.0f / (float)up;
  const uint8_t* g = gt + (int64_t)b * H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int inter = 0, npred = 0, ngt = 0;
  for (int y = warp; y < H; y += kEvalWarps) {
    RowTap r = row_tap(plane, y, hp, wp, rscale);
    const uint8_t* grow = g + (int64_t)y * W;
    for (int x = lane; x < W; x += 32) {
      float v = pixel(r, x, wp, rscale);
      int p = v > 0.5f, t = grow[x] != 0;
      inter += p & t;
      npred += p;
      ngt += t;
    }
  }
  inter = block_sum_int(inter, red);
  npred = block_sum_int(npred, red);
  ngt = block_sum_int(ngt, red);
  if (threadIdx.x == 0) {
    int32_t* o = q_counts + ((int64_t)b * nq + q) * SMK_QCOUNT_STRIDE;
    o[0] = inter;
    o[1] = npred + ngt - inter;   // |p ∪ g|
  }
}

// ------------------------------------------------------------------------------------------------
// E2: all reductions for one evaluated mask.  grid (n_masks_per_image, B).
//   kSelect = true : mask index chosen here (blockIdx.x 0 → objectness top-1, 1 → best-IoU query)
//   kSelect = false: blockIdx.x-th plane of `planes` as is (full-resolution masks, up = 1)
// ------------------------------------------------------------------------------------------------
struct MaskAcc {
  double sp = 0, sabs = 0, fg_p = 0, fg_p2 = 0, bg_q = 0, bg_q2 = 0;
  // region sums by inclusion-exclusion: all / left (x<X) / top (y<Y) / top-left
  double p_all = 0, p_l = 0, p_t = 0, p_tl = 0;
  double p2_all = 0, p2_l = 0, p2_t = 0, p2_tl = 0;
  double pg_all = 0, pg_l = 0, pg_t = 0, pg_tl = 0;
  int g_l = 0, g_t = 0, g_tl = 0;
  int tp05 = 0, tpfp05 = 0;
};

template <bool kSelect, bool kSmemPlane>
__global__ void __launch_bounds__(kEvalThreads)
mask_metrics_kernel(const float* __restrict__ planes, int64_t batch_stride, const float* __restrict__ objectness,
                    int64_t obj_stride, const int32_t* __restrict__ q_counts, const uint8_t* __restrict__ gt,
                    int nq, int hp, int wp, int up, int H, int W,
                    int32_t* __restrict__ idx_out, int32_t* __restrict__ m_counts, double* __restrict__ m_sums) {
  extern __shared__ float dyn[];
  __shared__ int hist[kEvalWarps][512];
  __shared__ float thr[256];
  __shared__ int red_i[kEvalWarps];
  __shared__ double red_d[kEvalWarps];
  __shared__ int s_sel;
  const int which = blockIdx.x, b = blockIdx.y;
  const int n_masks = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    int sel = which;
    if (kSelect) {
      if (which == 0) {           // objectness top-1 (evaluator.pyc@L219-221); lowest index on ties
        const float* ob = objectness + (int64_t)b * obj_stride;
        float best = ob[0];
        sel = 0;
        for (int i = 1; i < nq; ++i)
          if (ob[i] > best) { best = ob[i]; sel = i; }
      } else {                    // upper bound: argmax_q I/(U+1e-7) in float32, first maximum
        const int32_t* qc = q_counts + (int64_t)b * nq * SMK_QCOUNT_STRIDE;
        float best = -1.0f;
        sel = 0;
        for (int i = 0; i < nq; ++i) {
          float iou = __fdiv_rn((float)qc[2 * i], __fadd_rn((float)qc[2 * i + 1], 1e-7f));
          if (iou > best) { best = iou; sel = i; }
        }
      }
      idx_out[(int64_t)b * 2 + which] = sel;
    }
    s_sel = sel;
  }
  for (int i = threadIdx.x; i < kEvalWarps * 512; i += kEvalThreads) (&hist[0][0])[i] = 0;
  if (threadIdx.x < 255) thr[threadIdx.x] = fmax_threshold(threadIdx.x);
  if (threadIdx.x == 255) thr[255] = 3.0e38f;
  __syncthreads();
  const int sel = s_sel;
  const float* src = planes + (int64_t)b * batch_stride + (int64_t)sel * hp * wp;
  const float* pl = src;
  if (kSmemPlane) {
    for (int i = threadIdx.x; i < hp * wp; i += kEvalThreads) dyn[i] = src[i];
    pl = dyn;
    __syncthreads();
  }
  const float rscale = 1.0f / (float)up;
  const uint8_t* g = gt + (int64_t)b * H * W;

  // ---- pass A over the ground truth: area and centroid (s_measure.py:11-31) -----------------
  int ng = 0, sxg = 0, syg = 0;
  for (int y = warp; y < H; y += kEvalWarps) {
    const uint8_t* grow = g + (int64_t)y * W;
    int rowc = 0;
    for (int x = lane; x < W; x += 32) {
      int t = grow[x] != 0;
      rowc += t;
      sxg += t * x;
    }
    ng += rowc;
    syg += rowc * y;
  }
  ng = block_sum_int(ng, red_i);
  sxg = block_sum_int(sxg, red_i);
  syg = block_sum_int(syg, red_i);
  int X, Y;
  if (ng == 0) {   // python round(cols / 2): half-to-even
    X = (int)rint((double)W / 2.0);
    Y = (int)rint((double)H / 2.0);
  } else {         // torch.round(float32 ratio): half-to-even
    X = (int)rintf(__fdiv_rn((float)sxg, (float)ng));
    Y = (int)rintf(__fdiv_rn((float)syg, (float)ng));
  }

  // ---- pass B: histograms, counts at 0.5, moments ---------------------------------------------
  MaskAcc a;
  int* myhist = hist[warp];
  for (int y = warp; y < H; y += kEvalWarps) {
    RowTap r = row_tap(pl, y, hp, wp, rscale);
    const uint8_t* grow = g + (int64_t)y * W;
    const bool top = y < Y;
    for (int x0 = 0; x0 < W; x0 += 32) {
      const int x = x0 + lane;
      const bool valid = x < W;
      float v = 0.f;
      int t = 0;
      if (valid) {
        v = (up == 1) ? r.r0[x] : pixel(r, x, wp, rscale);
        t = grow[x] != 0;
      }
      // bin = #{k : t_k < v}, exact float32 thresholds, strict compare (f_measure.py:65, :45)
      int k = min(max((int)(v * 255.0f), 0), 255);
      while (k < 255 && thr[k] < v) ++k;
      while (k > 0 && !(thr[k - 1] < v)) --k;
      const int key = valid ? (t ? k : 256 + k) : 1024;
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&myhist[key], __popc(peers));
      if (valid) {
        const int p = v > 0.5f;
        a.tp05 += p & t;
        a.tpfp05 += p;
        const double dv = (double)v, dv2 = dv * dv, dg = t ? dv : 0.0;
        a.sp += dv;
        a.sabs += fabs(dv - (t ? 1.0 : 0.0));
        if (t) { a.fg_p += dv; a.fg_p2 += dv2; }
        else   { const double w = 1.0 - dv; a.bg_q += w; a.bg_q2 += w * w; }
        const bool left = x < X;
        a.p_all += dv; a.p2_all += dv2; a.pg_all += dg;
        if (left) { a.p_l += dv; a.p2_l += dv2; a.pg_l += dg; a.g_l += t; }
        if (top)  { a.p_t += dv; a.p2_t += dv2; a.pg_t += dg; a.g_t += t; }
        if (left && top) { a.p_tl += dv; a.p2_tl += dv2; a.pg_tl += dg; a.g_tl += t; }
      }
    }
  }
  const int64_t mrow = (int64_t)b * n_masks + which;
  int32_t* oc = m_counts + mrow * SMK_MCOUNT_STRIDE;
  double* os = m_sums + mrow * SMK_MSUM_STRIDE;

  const double sp = block_sum_double(a.sp, red_d);
  const double sabs = block_sum_double(a.sabs, red_d);
  const double fg_p = block_sum_double(a.fg_p, red_d), fg_p2 = block_sum_double(a.fg_p2, red_d);
  const double bg_q = block_sum_double(a.bg_q, red_d), bg_q2 = block_sum_double(a.bg_q2, red_d);
  const double p_all = block_sum_double(a.p_all, red_d), p_l = block_sum_double(a.p_l, red_d);
  const double p_t = block_sum_double(a.p_t, red_d), p_tl = block_sum_double(a.p_tl, red_d);
  const double p2_all = block_sum_double(a.p2_all, red_d), p2_l = block_sum_double(a.p2_l, red_d);
  const double p2_t = block_sum_double(a.p2_t, red_d), p2_tl = block_sum_double(a.p2_tl, red_d);
  const double pg_all = block_sum_double(a.pg_all, red_d), pg_l = block_sum_double(a.pg_l, red_d);
  const double pg_t = block_sum_double(a.pg_t, red_d), pg_tl = block_sum_double(a.pg_tl, red_d);
  const int g_l = block_sum_int(a.g_l, red_i), g_t = block_sum_int(a.g_t, red_i), g_tl = block_sum_int(a.g_tl, red_i);
  const int tp05 = block_sum_int(a.tp05, red_i), tpfp05 = block_sum_int(a.tpfp05, red_i);

  // F-mean threshold (f_measure.py:76): 2 * mean(p) in float32
  const int npix = H * W;
  const float tau = 2.0f * (float)(sp / (double)npix);

  // ---- pass C: counts at the adaptive threshold -----------------------------------------------
  int tpm = 0, tpfpm = 0;
  for (int y = warp; y < H; y += kEvalWarps) {
    RowTap r = row_tap(pl, y, hp, wp, rscale);
    const uint8_t* grow = g + (int64_t)y * W;
    for (int x = lane; x < W; x += 32) {
      float v = (up == 1) ? r.r0[x] : pixel(r, x, wp, rscale);
      int p = v > tau, t = grow[x] != 0;
      tpm += p & t;
      tpfpm += p;
    }
  }
  tpm = block_sum_int(tpm, red_i);
  tpfpm = block_sum_int(tpfpm, red_i);

  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += kEvalThreads) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < kEvalWarps; ++w) s += hist[w][i];
    oc[i] = s;
  }
  if (threadIdx.x == 0) {
    oc[512] = tp05; oc[513] = tpfp05; oc[514] = ng; oc[515] = tpm; oc[516] = tpfpm;
    oc[517] = X; oc[518] = Y; oc[519] = npix; oc[520] = sel;
    for (int i = 521; i < SMK_MCOUNT_STRIDE; ++i) oc[i] = 0;
    os[0] = sp; os[1] = sabs; os[2] = (double)tau; os[3] = fg_p; os[4] = fg_p2; os[5] = bg_q; os[6] = bg_q2; os[7] = 0;
    // quadrants LT, RT, LB, RB (s_measure.py:71-94): [:Y,:X] [:Y,X:] [Y:,:X] [Y:,X:]
    const int Xc = min(max(X, 0), W), Yc = min(max(Y, 0), H);
    const double n_q[4] = {(double)Xc * Yc, (double)(W - Xc) * Yc, (double)Xc * (H - Yc), (double)(W - Xc) * (H - Yc)};
    const double sp_q[4] = {p_tl, p_t - p_tl, p_l - p_tl, p_all - p_t - p_l + p_tl};
    const double sp2_q[4] = {p2_tl, p2_t - p2_tl, p2_l - p2_tl, p2_all - p2_t - p2_l + p2_tl};
    const double spg_q[4] = {pg_tl, pg_t - pg_tl, pg_l - pg_tl, pg_all - pg_t - pg_l + pg_tl};
    const double sg_q[4] = {(double)g_tl, (double)(g_t - g_tl), (double)(g_l - g_tl), (double)(ng - g_t - g_l + g_tl)};
    for (int qd = 0; qd < 4; ++qd) {
      os[8 + 5 * qd + 0] = n_q[qd]; os[8 + 5 * qd + 1] = sp_q[qd]; os[8 + 5 * qd + 2] = sp2_q[qd];
      os[8 + 5 * qd + 3] = sg_q[qd]; os[8 + 5 * qd + 4] = spg_q[qd];
    }
    for (int i = 28; i < SMK_MSUM_STRIDE; ++i) os[i] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// x4 fast paths (up == 4, W % 4 == 0, W <= 1024): a thread owns one source column i — the 4 full-resolution pixels
// 4i .. 4i+3 of a row come from 6 shared-memory reads (quad4) and one 4-byte ground-truth load — and walks rows
// y = ys, ys + YS, ...  Same arithmetic as the generic kernels (make_tap + bilerp), so every count is identical.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEvalThreads)
query_iou_x4_kernel(const float* __restrict__ mask_pred, int64_t batch_stride, const uint8_t* __restrict__ gt,
                    int nq, int hp, int wp, int H, int W, int32_t* __restrict__ q_counts) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float plane[];
  __shared__ int red[kEvalWarps];
  const int q = blockIdx.x, b = blockIdx.y;
  const float* src = mask_pred + (int64_t)b * batch_stride + (int64_t)q * hp * wp;
  for (int i = threadIdx.x; i < hp * wp; i += kEvalThreads) plane[i] = src[i];
  __syncthreads();
  const int Q4 = W >> 2, YS = kEvalThreads / Q4;
  const int i = threadIdx.x % Q4, ys = threadIdx.x / Q4;
  const uint8_t* g = gt + (int64_t)b * H * W;
  int inter = 0, npred = 0, ngt = 0;
  if (ys < YS) {
    const QuadX qx = make_quadx(i, wp);
    for (int y = ys; y < H; y += YS) {
      const Tap ty = make_tap(y, 0.25f, hp);
      float v[4];
      quad4(plane + ty.i0 * wp, plane + ty.i1 * wp, qx, ty.l0, ty.l1, v);
      const uchar4 gq = *reinterpret_cast<const uchar4*>(g + (int64_t)y * W + 4 * i);
      const unsigned pm = (v[0] > 0.5f ? 1u : 0u) | (v[1] > 0.5f ? 2u : 0u) | (v[2] > 0.5f ? 4u : 0u) | (v[3] > 0.5f ? 8u : 0u);
      const unsigned tm = (gq.x ? 1u : 0u) | (gq.y ? 2u : 0u) | (gq.z ? 4u : 0u) | (gq.w ? 8u : 0u);
      inter += __popc(pm & tm);
      npred += __popc(pm);
      ngt += __popc(tm);
    }
  }
  inter = block_sum_int(inter, red);
  npred = block_sum_int(npred, red);
  ngt = block_sum_int(ngt, red);
  if (threadIdx.x == 0) {
    int32_t* o = q_counts + ((int64_t)b * nq + q) * SMK_QCOUNT_STRIDE;
    o[0] = inter;
    o[1] = npred + ngt - inter;   // |p ∪ g|
  }
}

// ------------------------------------------------------------------------------------------------
// E1 by source cells (up == 4), one warp per query.  The x4 bilinear map (align_corners = False) cuts the output into cells:
// columns {0,1}, then {4u-2 .. 4u+1} for u = 1 .. wp-1, then the last two columns (same along y); every pixel of cell (v, u) is
// a blend, with non-negative dyadic weights that sum to 1, of the 4 plane samples at rows {max(v-1,0), min(v,hp-1)} x columns
// {max(u-1,0), min(u,wp-1)}.  Round-to-nearest is monotone, so bilerp() is monotone in each corner, and bilerp(c,c,c,c) is
// exactly 0.5 for c = 0.5 and > 0.5 for c = nextafter(0.5) (checked for every weight pair in tests/test_host_cpu.py).
// Hence: 4 corners > 0.5 → every pixel > 0.5; no corner > 0.5 → no pixel > 0.5.  Only cells with mixed corners (2-5 % of the
// cells of a real mask) are evaluated pixel by pixel, with the same make_tap() + bilerp() arithmetic as the generic kernel, so
// every count stays bit-identical to the reference.
// Everything else is bit-vector work: the ground truth is packed to one bit per pixel once per CTA; a warp turns its plane
// into "sample > 0.5" bit rows with ballots, classifies a whole row of cells with shifts / and / or, expands the "all four
// corners" bits to pixel masks (one bit → one nibble, shifted by the 2-pixel cell offset) and counts with popc against the
// ground-truth words.  grid (ceil(nq / 8), B), 8 warps = 8 queries per CTA, no block barrier after the packing.
// ------------------------------------------------------------------------------------------------
constexpr int kCellMaxWords = 4;     // plane width <= 127 samples (bit rows of <= 4 words)

// 8 bits → 8 nibbles (bit j → bits 4j .. 4j+3)
__device__ __forceinline__ uint32_t expand_nibbles(uint32_t x) {
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x * 0xFu;
}
// bits [x, x + n) (n <= 6) of a bit-packed row of ww words
__device__ __forceinline__ unsigned row_bits(const uint32_t* row, int ww, int x, int n) {
  const int w = x >> 5;
  const uint32_t lo = row[w], hi = (w + 1 < ww) ? row[w + 1] : 0u;
  return __funnelshift_r(lo, hi, x & 31) & ((1u << n) - 1u);
}
// 32-bit window starting at bit `pos` of a bit row of nw words (zero beyond the row)
__device__ __forceinline__ uint32_t bit_window(const uint32_t* row, int nw, int pos) {
  const int w = pos >> 5;
  const uint32_t lo = w < nw ? row[w] : 0u, hi = (w + 1 < nw) ? row[w + 1] : 0u;
  return __funnelshift_r(lo, hi, pos & 31);
}

// Ground truth [H,W] bytes → one bit per pixel, rows of ww 32-bit words in shared memory (W % 4 == 0, g 4-byte aligned).
// Block-cooperative; returns this thread's share of the foreground count.  The caller's next barrier publishes gbits.
__device__ __forceinline__ int pack_gt_bits(const uint8_t* __restrict__ g, uint32_t* gbits, int H, int W, int ww) {
  int ngt = 0;
  auto nibble = [](uint32_t px4) { return ((__vcmpne4(px4, 0u) & 0x08040201u) * 0x01010101u) >> 24; };   // byte j non-zero → bit j
  if ((W & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    // 16 pixels → one 16-bit half word; 4 independent 16-byte loads in flight per thread
    const int w16 = W >> 4, n16 = H * w16;
    uint16_t* gb16 = reinterpret_cast<uint16_t*>(gbits);
    const uint4* g16 = reinterpret_cast<const uint4*>(g);
    for (int i0 = threadIdx.x; i0 < n16; i0 += 4 * kEvalThreads) {
      uint4 px[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * kEvalThreads;
        px[j] = i < n16 ? __ldg(g16 + i) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * kEvalThreads;
        if (i < n16) {
          const uint32_t bits = nibble(px[j].x) | (nibble(px[j].y) << 4) | (nibble(px[j].z) << 8) | (nibble(px[j].w) << 12);
          const int y = i / w16, x16 = i - y * w16;
          gb16[y * (2 * ww) + x16] = (uint16_t)bits;
          ngt += __popc(bits);
        }
      }
    }
    if (W & 16)       // odd number of half words per row: clear the padding half of the last word
      for (int y = threadIdx.x; y < H; y += kEvalThreads) gb16[y * (2 * ww) + 2 * ww - 1] = 0;
  } else {
    for (int i = threadIdx.x; i < H * ww; i += kEvalThreads) {
      const int y = i / ww, x0 = (i - y * ww) * 32;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(g + (int64_t)y * W + x0);
      uint32_t bits = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (x0 + 4 * k < W) bits |= nibble(__ldg(src + k)) << (4 * k);
      gbits[i] = bits;
      ngt += __popc(bits);
    }
  }
  return ngt;
}

__global__ void __launch_bounds__(kEvalThreads)
query_iou_cells_kernel(const float* __restrict__ mask_pred, int64_t batch_stride, const uint8_t* __restrict__ gt,
                       int nq, int hp, int wp, int H, int W, int32_t* __restrict__ q_counts) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint32_t dyn_cells[];
  __shared__ int red[kEvalWarps];
  const int ww = (W + 31) >> 5, aw = (wp + 31) >> 5, fw = (wp + 32) >> 5, n_cells = (hp + 1) * (wp + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* gbits = dyn_cells;                                            // [H][ww] ground truth, one bit per pixel
  uint32_t* abits = gbits + H * ww + warp * (hp * aw + 2 * (hp + 1) * fw);   // [hp][aw]  sample > 0.5
  uint32_t* fullw = abits + hp * aw;                                      // [hp+1][fw] cell: all 4 corners above
  uint32_t* mixw = fullw + (hp + 1) * fw;                                 // [hp+1][fw] cell: some, not all, corners above
  uint16_t* mixed = reinterpret_cast<uint16_t*>(gbits + H * ww + kEvalWarps * (hp * aw + 2 * (hp + 1) * fw)) + warp * n_cells;
  const int b = blockIdx.y;
  const uint8_t* g = gt + (int64_t)b * H * W;

  int ngt = pack_gt_bits(g, gbits, H, W, ww);
  ngt = block_sum_int(ngt, red);         // (contains the barrier that publishes gbits)
  const int q = blockIdx.x * kEvalWarps + warp;
  if (q >= nq) return;
  const float* src = mask_pred + (int64_t)b * batch_stride + (int64_t)q * hp * wp;

  // ---- phase 1: sample > 0.5 bit rows ----
  if ((wp & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // 4 consecutive samples of one row per 16-byte load → one nibble, OR-ed into the row's word; 8 loads in flight per lane
    for (int i = lane; i < hp * aw; i += 32) abits[i] = 0u;
    __syncwarp();
    const int w4 = wp >> 2, n4 = hp * w4;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (int i0 = lane; i0 < n4; i0 += 8 * 32) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = i0 + 32 * j < n4 ? __ldg(src4 + i0 + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + 32 * j;
        const uint32_t nib = (v[j].x > 0.5f ? 1u : 0u) | (v[j].y > 0.5f ? 2u : 0u) | (v[j].z > 0.5f ? 4u : 0u) | (v[j].w > 0.5f ? 8u : 0u);
        if (i < n4 && nib) {
          const int r = i / w4, c = 4 * (i - r * w4);
          atomicOr(&abits[r * aw + (c >> 5)], nib << (c & 31));
        }
      }
    }
  } else {
    // (8 rows of loads in flight before the first ballot: the plane comes from L2 / HBM, one latency per batch instead of per word)
    for (int rb = 0; rb < hp; rb += 8) {
      float v[8][kCellMaxWords];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < kCellMaxWords; ++k) {
          const int r = rb + j, c = 32 * k + lane;
          v[j][k] = (r < hp && c < wp) ? __ldg(src + r * wp + c) : 0.f;
        }
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < kCellMaxWords; ++k) {
          const uint32_t bal = __ballot_sync(0xffffffffu, v[j][k] > 0.5f);
          if (lane == 0 && rb + j < hp && k < aw) abits[(rb + j) * aw + k] = bal;
        }
    }
  }
  __syncwarp();

  // ---- phase 2: one lane per row of cells: corner combinations as shifted bit rows ----
  for (int v = lane; v <= hp; v += 32) {
    const uint32_t* a0 = abits + max(v - 1, 0) * aw;
    const uint32_t* a1 = abits + min(v, hp - 1) * aw;
    uint32_t all_[kCellMaxWords + 1], any_[kCellMaxWords + 1];
#pragma unroll
    for (int k = 0; k <= kCellMaxWords; ++k) {
      all_[k] = k < aw ? (a0[k] & a1[k]) : 0u;
      any_[k] = k < aw ? (a0[k] | a1[k]) : 0u;
    }
    // cell u uses columns max(u-1,0) and min(u,wp-1):  L = (S << 1) | S[0],  R = S | (S[wp-1] << wp)
    const uint32_t last0 = a0[(wp - 1) >> 5] >> ((wp - 1) & 31), last1 = a1[(wp - 1) >> 5] >> ((wp - 1) & 31);
    const uint32_t last_all = last0 & last1 & 1u, last_any = (last0 | last1) & 1u;
    uint32_t carry_all = all_[0] & 1u, carry_any = any_[0] & 1u;
#pragma unroll
    for (int k = 0; k <= kCellMaxWords; ++k) {
      if (k < fw) {
        const uint32_t l_all = (all_[k] << 1) | carry_all, l_any = (any_[k] << 1) | carry_any;
        carry_all = all_[k] >> 31;
        carry_any = any_[k] >> 31;
        uint32_t r_all = all_[k], r_any = any_[k];
        if (k == (wp >> 5)) { r_all |= last_all << (wp & 31); r_any |= last_any << (wp & 31); }
        const uint32_t full = l_all & r_all, some = l_any | r_any;
        // keep bits 0 .. wp only
        const int hi_bit = wp - 32 * k;      // index of the last valid bit in this word (>= 0 since k < fw)
        const uint32_t keep = hi_bit >= 31 ? 0xffffffffu : ((2u << hi_bit) - 1u);
        fullw[v * fw + k] = full & keep;
        mixw[v * fw + k] = some & ~full & keep;
      }
    }
  }
  __syncwarp();

  // ---- phase 3: (cell row, pixel word) items: popc of the expanded "full" cells, collection of the mixed ones ----
  int inter = 0, npred = 0, n_mixed = 0;
  const int n_items = (hp + 1) * ww, dv = 32 / ww, dl = 32 - dv * ww;
  int v = lane / ww, l = lane - v * ww;
  for (int it0 = 0; it0 < n_items; it0 += 32, v += dv, l += dl) {
    const int it = it0 + lane;
    uint32_t own = 0;
    if (l >= ww) { l -= ww; ++v; }
    if (it < n_items) {
      const int ya = v == 0 ? 0 : min(4 * v - 2, H), yb = v == 0 ? min(2, H) : min(4 * v + 2, H);
      const uint32_t fwin = bit_window(fullw + v * fw, fw, 8 * l), mwin = bit_window(mixw + v * fw, fw, 8 * l);
      // pixel x of this word lies in cell u = (x + 2) >> 2: nibble expansion in x' = x + 2, shifted back by 2
      uint32_t pm = (expand_nibbles(fwin & 0xFFu) >> 2) | ((fwin & 0x100u) ? 0xC0000000u : 0u);
      if (32 * l + 32 > W) pm &= (1u << (W - 32 * l)) - 1u;
      if (ya < yb) {
        npred += __popc(pm) * (yb - ya);
        for (int y = ya; y < yb; ++y) inter += __popc(pm & gbits[y * ww + l]);
        // a mixed cell belongs to the word that holds its first pixel: u = 8l+1 .. 8l+8 (and u = 0 for l = 0)
        own = ((mwin >> 1) & 0xFFu) << 1 | (l == 0 ? (mwin & 1u) : 0u);
      }
    }
    // warp-aggregated append of (v, u) pairs to this warp's list
    const int cnt = __popc(own);
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    int pos = n_mixed + pre - cnt;
    while (own) {
      const int j = __ffs(own) - 1;
      own &= own - 1;
      mixed[pos++] = (uint16_t)((v << 8) | (8 * l + j));
    }
    n_mixed += total;
  }
  __syncwarp();

  // ---- phase 4: mixed cells, pixel by pixel (shared taps: <= 4 columns x <= 4 rows per cell) ----
  for (int it = lane; it < n_mixed; it += 32) {
    const int v = mixed[it] >> 8, u = mixed[it] & 0xFF;
    const int ya = v == 0 ? 0 : min(4 * v - 2, H), yb = v == 0 ? min(2, H) : min(4 * v + 2, H);
    const int xa = u == 0 ? 0 : min(4 * u - 2, W), xb = u == 0 ? min(2, W) : min(4 * u + 2, W);
    float top[4], bot[4];
    const Tap t0 = make_tap(ya, 0.25f, hp);           // every row of the cell has the same source rows
    const float* r0 = src + t0.i0 * wp;
    const float* r1 = src + t0.i1 * wp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Tap tx = make_tap(min(xa + j, W - 1), 0.25f, wp);
      const float a = __ldg(r0 + tx.i0), bb = __ldg(r0 + tx.i1), c = __ldg(r1 + tx.i0), d = __ldg(r1 + tx.i1);
      top[j] = __fmaf_rn(a, tx.l0, __fmul_rn(bb, tx.l1));
      bot[j] = __fmaf_rn(c, tx.l0, __fmul_rn(d, tx.l1));
    }
    const int nx = xb - xa;
    for (int y = ya; y < yb; ++y) {
      const Tap ty = make_tap(y, 0.25f, hp);
      const unsigned tb = row_bits(gbits + y * ww, ww, xa, nx);
      unsigned pb = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nx && __fmaf_rn(top[j], ty.l0, __fmul_rn(bot[j], ty.l1)) > 0.5f) pb |= 1u << j;
      inter += __popc(pb & tb);
      npred += __popc(pb);
    }
  }
  inter = warp_sum(inter);
  npred = warp_sum(npred);
  if (lane == 0) {
    int32_t* o = q_counts + ((int64_t)b * nq + q) * SMK_QCOUNT_STRIDE;
    o[0] = inter;
    o[1] = npred + ngt - inter;   // |p ∪ g|
  }
}

// Block-cooperative count of the pixels above `thr` (and of those that are also foreground) of the plane in shared memory,
// by source cells like query_iou_cells_kernel.  A general threshold is not a power of two, so bilerp(c,c,c,c) may differ from
// c by an ulp: cells are skipped only behind a guard band — all 4 corners > t_hi = thr·(1+2e-6) (every pixel is a convex
// blend with <= 3 roundings of relative error 2^-24, hence > thr) or all 4 corners <= t_lo = thr·(1-2e-6) (no pixel > thr);
// every other cell is evaluated pixel by pixel with the exact compare.  Scratch: bit rows / lists in shared memory.
struct CellScratch {
  uint32_t *hi, *lo;        // [hp][aw] sample > t_hi / sample > t_lo
  uint32_t *fullw, *mixw;   // [hp+1][fw]
  uint16_t* mixed;          // [(hp+1)*(wp+1)]
  int* n_mixed;
};
__device__ __forceinline__ void block_cells_count(const float* plane, const uint32_t* gbits, int hp, int wp, int H, int W, float thr,
                                                  const CellScratch& sc, int& inter_out, int& npred_out) {
  const int ww = (W + 31) >> 5, aw = (wp + 31) >> 5, fw = (wp + 32) >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool guard_ok = thr > 1e-20f && thr < 1e20f;
  const float t_hi = guard_ok ? thr * (1.0f + 2e-6f) : 3.0e38f, t_lo = guard_ok ? thr * (1.0f - 2e-6f) : -3.0e38f;
  if (threadIdx.x == 0) *sc.n_mixed = 0;
  for (int r = warp; r < hp; r += kEvalWarps)
    for (int k = 0; k < aw; ++k) {
      const int c = 32 * k + lane;
      const float v = c < wp ? plane[r * wp + c] : -3.0e38f;
      const uint32_t bh = __ballot_sync(0xffffffffu, v > t_hi), bl = __ballot_sync(0xffffffffu, v > t_lo);
      if (lane == 0) { sc.hi[r * aw + k] = bh; sc.lo[r * aw + k] = bl; }
    }
  __syncthreads();
  for (int v = threadIdx.x; v <= hp; v += kEvalThreads) {
    const int ra = max(v - 1, 0) * aw, rb = min(v, hp - 1) * aw;
    const int lw = (wp - 1) >> 5, ls = (wp - 1) & 31;
    const uint32_t last_all = (sc.hi[ra + lw] >> ls) & (sc.hi[rb + lw] >> ls) & 1u, last_any = ((sc.lo[ra + lw] | sc.lo[rb + lw]) >> ls) & 1u;
    uint32_t carry_all = sc.hi[ra] & sc.hi[rb] & 1u, carry_any = (sc.lo[ra] | sc.lo[rb]) & 1u;
    for (int k = 0; k < fw; ++k) {
      const uint32_t all_k = k < aw ? (sc.hi[ra + k] & sc.hi[rb + k]) : 0u, any_k = k < aw ? (sc.lo[ra + k] | sc.lo[rb + k]) : 0u;
      const uint32_t l_all = (all_k << 1) | carry_all, l_any = (any_k << 1) | carry_any;
      carry_all = all_k >> 31;
      carry_any = any_k >> 31;
      uint32_t r_all = all_k, r_any = any_k;
      if (k == (wp >> 5)) { r_all |= last_all << (wp & 31); r_any |= last_any << (wp & 31); }
      const uint32_t full = l_all & r_all, some = l_any | r_any;
      const int hi_bit = wp - 32 * k;
      const uint32_t keep = hi_bit >= 31 ? 0xffffffffu : ((2u << hi_bit) - 1u);
      sc.fullw[v * fw + k] = full & keep;
      sc.mixw[v * fw + k] = some & ~full & keep;
    }
  }
  __syncthreads();
  int inter = 0, npred = 0;
  const int n_items = (hp + 1) * ww;
  for (int it = threadIdx.x; it < n_items; it += kEvalThreads) {
    const int v = it / ww, l = it - v * ww;
    const int ya = v == 0 ? 0 : min(4 * v - 2, H), yb = v == 0 ? min(2, H) : min(4 * v + 2, H);
    if (ya >= yb) continue;
    const uint32_t fwin = bit_window(sc.fullw + v * fw, fw, 8 * l), mwin = bit_window(sc.mixw + v * fw, fw, 8 * l);
    uint32_t pm = (expand_nibbles(fwin & 0xFFu) >> 2) | ((fwin & 0x100u) ? 0xC0000000u : 0u);
    if (32 * l + 32 > W) pm &= (1u << (W - 32 * l)) - 1u;
    npred += __popc(pm) * (yb - ya);
    for (int y = ya; y < yb; ++y) inter += __popc(pm & gbits[y * ww + l]);
    uint32_t own = ((mwin >> 1) & 0xFFu) << 1 | (l == 0 ? (mwin & 1u) : 0u);
    if (own) {
      int pos = atomicAdd(sc.n_mixed, __popc(own));
      while (own) {
        const int j = __ffs(own) - 1;
        own &= own - 1;
        sc.mixed[pos++] = (uint16_t)((v << 8) | (8 * l + j));
      }
    }
  }
  __syncthreads();
  const int n_mixed = *sc.n_mixed;
  for (int it = threadIdx.x; it < n_mixed; it += kEvalThreads) {
    const int v = sc.mixed[it] >> 8, u = sc.mixed[it] & 0xFF;
    const int ya = v == 0 ? 0 : min(4 * v - 2, H), yb = v == 0 ? min(2, H) : min(4 * v + 2, H);
    const int xa = u == 0 ? 0 : min(4 * u - 2, W), xb = u == 0 ? min(2, W) : min(4 * u + 2, W);
    float top[4], bot[4];
    const Tap t0 = make_tap(ya, 0.25f, hp);
    const float* r0 = plane + t0.i0 * wp;
    const float* r1 = plane + t0.i1 * wp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Tap tx = make_tap(min(xa + j, W - 1), 0.25f, wp);
      top[j] = __fmaf_rn(r0[tx.i0], tx.l0, __fmul_rn(r0[tx.i1], tx.l1));
      bot[j] = __fmaf_rn(r1[tx.i0], tx.l0, __fmul_rn(r1[tx.i1], tx.l1));
    }
    const int nx = xb - xa;
    for (int y = ya; y < yb; ++y) {
      const Tap ty = make_tap(y, 0.25f, hp);
      const unsigned tb = row_bits(gbits + y * ww, ww, xa, nx);
      unsigned pb = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nx && __fmaf_rn(top[j], ty.l0, __fmul_rn(bot[j], ty.l1)) > thr) pb |= 1u << j;
      inter += __popc(pb & tb);
      npred += __popc(pb);
    }
  }
  inter_out = inter;
  npred_out = npred;
}

// sum of the positions of the set bits of a word
__device__ __forceinline__ int bitpos_sum(uint32_t w) {
  return __popc(w & 0xAAAAAAAAu) + 2 * __popc(w & 0xCCCCCCCCu) + 4 * __popc(w & 0xF0F0F0F0u) + 8 * __popc(w & 0xFF00FF00u) +
         16 * __popc(w & 0xFFFF0000u);
}

// All reductions for the objectness-selected (blockIdx.x == 0) and best-IoU (1) mask of image blockIdx.y, x4 path.
// A thread owns the 4 pixels x = 4qi .. 4qi+3 and walks rows of source cells: the horizontal blends (top / bot) are shared by
// the <= 4 pixel rows of a cell row, the ground truth comes from the bit-packed copy, histogram updates are run-length
// aggregated per thread (neighbouring pixels of a smooth mask fall into the same bin), the float64 moments are kept for the
// regions {rows above / below the centroid} x {all / left-of-centroid columns} and everything else is derived from them:
//   sum|p-g| = (n_fg - S_fg p) + (S p - S_fg p),  S_bg (1-p) = n_bg - (S p - S_fg p),  S_bg (1-p)^2 = n_bg - 2 S_bg p + S_bg p^2.
// Same pixel arithmetic as the generic kernel (make_tap + bilerp) → identical integer outputs.
__global__ void __launch_bounds__(kEvalThreads)
mask_metrics_x4_kernel(const float* __restrict__ planes, int64_t batch_stride, const float* __restrict__ objectness,
                       int64_t obj_stride, const int32_t* __restrict__ q_counts, const uint8_t* __restrict__ gt,
                       int nq, int hp, int wp, int H, int W,
                       int32_t* __restrict__ idx_out, int32_t* __restrict__ m_counts, double* __restrict__ m_sums) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint32_t dyn_mm[];
  __shared__ int hist[kEvalWarps][512];
  __shared__ float thr[258];
  constexpr int kND = 13, kNI = 8;
  __shared__ double red_d[kND][kEvalWarps];
  __shared__ int red_i[kNI][kEvalWarps];
  __shared__ int s_sel, s_nmixed;
  const int which = blockIdx.x, b = blockIdx.y;
  const int n_masks = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ww = (W + 31) >> 5, aw = (wp + 31) >> 5, fw = (wp + 32) >> 5;
  float* plane = reinterpret_cast<float*>(dyn_mm);            // [hp*wp]
  uint32_t* gbits = dyn_mm + hp * wp;                         // [H][ww]
  CellScratch sc;
  sc.hi = gbits + H * ww;
  sc.lo = sc.hi + hp * aw;
  sc.fullw = sc.lo + hp * aw;
  sc.mixw = sc.fullw + (hp + 1) * fw;
  sc.mixed = reinterpret_cast<uint16_t*>(sc.mixw + (hp + 1) * fw);
  sc.n_mixed = &s_nmixed;

  if (threadIdx.x == 0) {
    int sel = 0;
    if (which == 0) {           // objectness top-1 (evaluator.pyc@L219-221); lowest index on ties
      const float* ob = objectness + (int64_t)b * obj_stride;
      float best = ob[0];
      for (int i = 1; i < nq; ++i)
        if (ob[i] > best) { best = ob[i]; sel = i; }
    } else {                    // upper bound: argmax_q I/(U+1e-7) in float32, first maximum
      const int32_t* qc = q_counts + (int64_t)b * nq * SMK_QCOUNT_STRIDE;
      float best = -1.0f;
      for (int i = 0; i < nq; ++i) {
        float iou = __fdiv_rn((float)qc[2 * i], __fadd_rn((float)qc[2 * i + 1], 1e-7f));
        if (iou > best) { best = iou; sel = i; }
      }
    }
    idx_out[(int64_t)b * 2 + which] = sel;
    s_sel = sel;
  }
  for (int i = threadIdx.x; i < kEvalWarps * 512; i += kEvalThreads) (&hist[0][0])[i] = 0;
  for (int i = threadIdx.x; i < 258; i += kEvalThreads) thr[i] = i < 255 ? fmax_threshold(i) : 3.0e38f;
  const uint8_t* g = gt + (int64_t)b * H * W;
  int ng = pack_gt_bits(g, gbits, H, W, ww);
  __syncthreads();
  const int sel = s_sel;
  const float* src = planes + (int64_t)b * batch_stride + (int64_t)sel * hp * wp;
  for (int i = threadIdx.x; i < hp * wp; i += kEvalThreads) plane[i] = src[i];

  // ---- pass A over the packed ground truth: area and centroid (s_measure.py:11-31) -----------------
  int sxg = 0, syg = 0;
  for (int i = threadIdx.x; i < H * ww; i += kEvalThreads) {
    const uint32_t w = gbits[i];
    if (w) {
      const int y = i / ww, c = __popc(w);
      sxg += c * ((i - y * ww) * 32) + bitpos_sum(w);
      syg += c * y;
    }
  }
  ng = warp_sum(ng); sxg = warp_sum(sxg); syg = warp_sum(syg);
  if (lane == 0) { red_i[0][warp] = ng; red_i[1][warp] = sxg; red_i[2][warp] = syg; }
  __syncthreads();               // also publishes the plane
  ng = sxg = syg = 0;
#pragma unroll
  for (int w = 0; w < kEvalWarps; ++w) { ng += red_i[0][w]; sxg += red_i[1][w]; syg += red_i[2][w]; }
  int X, Y;
  if (ng == 0) {   // python round(cols / 2): half-to-even
    X = (int)rint((double)W / 2.0);
    Y = (int)rint((double)H / 2.0);
  } else {         // torch.round(float32 ratio): half-to-even
    X = (int)rintf(__fdiv_rn((float)sxg, (float)ng));
    Y = (int)rintf(__fdiv_rn((float)syg, (float)ng));
  }
  __syncthreads();               // red_i is reused below

  // ---- pass B: histograms, counts at 0.5, float64 moments ---------------------------------------
  const int Q4 = W >> 2, YS = kEvalThreads / Q4;
  const int qi = threadIdx.x % Q4, ys = threadIdx.x / Q4;
  const int x0 = 4 * qi;
  const int n_left = min(max(X - x0, 0), 4);   // how many of this thread's 4 pixels lie left of the centroid column
  const unsigned lmask = (1u << n_left) - 1u;
  double Tp = 0, Tp2 = 0, Tpg = 0, TLp = 0, TLp2 = 0, TLpg = 0, Bp = 0, Bp2 = 0, Bpg = 0, BLp = 0, BLp2 = 0, BLpg = 0, fgp2 = 0;
  int g_l = 0, g_t = 0, g_tl = 0, tp05 = 0, tpfp05 = 0;
  int cur_key = 0, cur_cnt = 0;
  int* myhist = hist[warp];
  if (ys < YS) {
    const QuadX qx = make_quadx(qi, wp);
    for (int v = ys; v <= hp; v += YS) {
      const int ya = v == 0 ? 0 : min(4 * v - 2, H), yb = v == 0 ? min(2, H) : min(4 * v + 2, H);
      if (ya >= yb) break;
      const float* r0 = plane + max(v - 1, 0) * wp;
      const float* r1 = plane + min(v, hp - 1) * wp;
      const float am = r0[qx.cm], ac = r0[qx.cc], ap = r0[qx.cp], bm = r1[qx.cm], bc = r1[qx.cc], bp = r1[qx.cp];
      float top[4], bot[4];
      top[0] = __fmaf_rn(am, qx.l0[0], __fmul_rn(ac, qx.l1[0])); bot[0] = __fmaf_rn(bm, qx.l0[0], __fmul_rn(bc, qx.l1[0]));
      top[1] = __fmaf_rn(am, qx.l0[1], __fmul_rn(ac, qx.l1[1])); bot[1] = __fmaf_rn(bm, qx.l0[1], __fmul_rn(bc, qx.l1[1]));
      top[2] = __fmaf_rn(ac, qx.l0[2], __fmul_rn(ap, qx.l1[2])); bot[2] = __fmaf_rn(bc, qx.l0[2], __fmul_rn(bp, qx.l1[2]));
      top[3] = __fmaf_rn(ac, qx.l0[3], __fmul_rn(ap, qx.l1[3])); bot[3] = __fmaf_rn(bc, qx.l0[3], __fmul_rn(bp, qx.l1[3]));
      for (int y = ya; y < yb; ++y) {
        const Tap ty = make_tap(y, 0.25f, hp);
        const unsigned tm = (gbits[y * ww + (x0 >> 5)] >> (x0 & 31)) & 0xFu;
        double qp = 0, qp2 = 0, qpg = 0, lp = 0, lp2 = 0, lpg = 0;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          const float vv = __fmaf_rn(top[f], ty.l0, __fmul_rn(bot[f], ty.l1));
          const int t = (tm >> f) & 1;
          // bin = #{k : t_k < v}, exact float32 thresholds, strict compare (f_measure.py:65, :45); trunc(v*255) is at most 2 low
          const int k0 = min(max((int)(vv * 255.0f), 0), 255);
          const int key = k0 + (thr[k0] < vv ? 1 : 0) + (thr[k0 + 1] < vv ? 1 : 0) + (t ? 0 : 256);
          if (key == cur_key) {
            ++cur_cnt;
          } else {
            if (cur_cnt) atomicAdd(&myhist[cur_key], cur_cnt);
            cur_key = key;
            cur_cnt = 1;
          }
          const int p = vv > 0.5f;
          tp05 += p & t;
          tpfp05 += p;
          const double dv = (double)vv;
          qp += dv;
          qp2 = fma(dv, dv, qp2);
          if (t) { qpg += dv; fgp2 = fma(dv, dv, fgp2); }
          if (f == n_left - 1) { lp = qp; lp2 = qp2; lpg = qpg; }     // prefix over the pixels left of the centroid column
        }
        const int gl = __popc(tm & lmask);
        g_l += gl;
        if (y < Y) { Tp += qp; Tp2 += qp2; Tpg += qpg; TLp += lp; TLp2 += lp2; TLpg += lpg; g_t += __popc(tm); g_tl += gl; }
        else       { Bp += qp; Bp2 += qp2; Bpg += qpg; BLp += lp; BLp2 += lp2; BLpg += lpg; }
      }
    }
    if (cur_cnt) atomicAdd(&myhist[cur_key], cur_cnt);
  }
  {
    const double dv_[kND] = {Tp, Tp2, Tpg, TLp, TLp2, TLpg, Bp, Bp2, Bpg, BLp, BLp2, BLpg, fgp2};
#pragma unroll
    for (int i = 0; i < kND; ++i) {
      const double r = warp_sum(dv_[i]);
      if (lane == 0) red_d[i][warp] = r;
    }
    const int iv_[5] = {g_l, g_t, g_tl, tp05, tpfp05};
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int r = warp_sum(iv_[i]);
      if (lane == 0) red_i[i][warp] = r;
    }
  }
  __syncthreads();
  double sd[kND];
#pragma unroll
  for (int i = 0; i < kND; ++i) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kEvalWarps; ++w) t += red_d[i][w];     // fixed order → run-to-run deterministic
    sd[i] = t;
  }
  int si[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kEvalWarps; ++w) t += red_i[i][w];
    si[i] = t;
  }
  const double p_all = sd[0] + sd[6], p2_all = sd[1] + sd[7], pg_all = sd[2] + sd[8];
  const double p_t = sd[0], p2_t = sd[1], pg_t = sd[2], p_tl = sd[3], p2_tl = sd[4], pg_tl = sd[5];
  const double p_l = sd[3] + sd[9], p2_l = sd[4] + sd[10], pg_l = sd[5] + sd[11];
  const double sp = p_all, fg_p = pg_all, fg_p2 = sd[12];
  g_l = si[0]; g_t = si[1]; g_tl = si[2]; tp05 = si[3]; tpfp05 = si[4];
  const int npix = H * W;
  const double n_fg = (double)ng, n_bg = (double)(npix - ng), bg_p = sp - fg_p;
  const double sabs = (n_fg - fg_p) + bg_p, bg_q = n_bg - bg_p, bg_q2 = n_bg - 2.0 * bg_p + (p2_all - fg_p2);

  // F-mean threshold (f_measure.py:76): 2 * mean(p) in float32
  const float tau = 2.0f * (float)(sp / (double)npix);

  // ---- pass C: counts at the adaptive threshold (cells behind a guard band, the rest per pixel) -----
  int tpm = 0, tpfpm = 0;
  block_cells_count(plane, gbits, hp, wp, H, W, tau, sc, tpm, tpfpm);
  tpm = warp_sum(tpm);
  tpfpm = warp_sum(tpfpm);
  __syncthreads();
  if (lane == 0) { red_i[5][warp] = tpm; red_i[6][warp] = tpfpm; }
  __syncthreads();
  tpm = tpfpm = 0;
#pragma unroll
  for (int w = 0; w < kEvalWarps; ++w) { tpm += red_i[5][w]; tpfpm += red_i[6][w]; }

  const int64_t mrow = (int64_t)b * n_masks + which;
  int32_t* oc = m_counts + mrow * SMK_MCOUNT_STRIDE;
  double* os = m_sums + mrow * SMK_MSUM_STRIDE;
  for (int i = threadIdx.x; i < 512; i += kEvalThreads) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < kEvalWarps; ++w) s += hist[w][i];
    oc[i] = s;
  }
  if (threadIdx.x == 0) {
    oc[512] = tp05; oc[513] = tpfp05; oc[514] = ng; oc[515] = tpm; oc[516] = tpfpm;
    oc[517] = X; oc[518] = Y; oc[519] = npix; oc[520] = sel;
    for (int i = 521; i < SMK_MCOUNT_STRIDE; ++i) oc[i] = 0;
    os[0] = sp; os[1] = sabs; os[2] = (double)tau; os[3] = fg_p; os[4] = fg_p2; os[5] = bg_q; os[6] = bg_q2; os[7] = 0;
    // quadrants LT, RT, LB, RB (s_measure.py:71-94): [:Y,:X] [:Y,X:] [Y:,:X] [Y:,X:]
    const int Xc = min(max(X, 0), W), Yc = min(max(Y, 0), H);
    const double n_q[4] = {(double)Xc * Yc, (double)(W - Xc) * Yc, (double)Xc * (H - Yc), (double)(W - Xc) * (H - Yc)};
    const double sp_q[4] = {p_tl, p_t - p_tl, p_l - p_tl, p_all - p_t - p_l + p_tl};
    const double sp2_q[4] = {p2_tl, p2_t - p2_tl, p2_l - p2_tl, p2_all - p2_t - p2_l + p2_tl};
    const double spg_q[4] = {pg_tl, pg_t - pg_tl, pg_l - pg_tl, pg_all - pg_t - pg_l + pg_tl};
    const double sg_q[4] = {(double)g_tl, (double)(g_t - g_tl), (double)(g_l - g_tl), (double)(ng - g_t - g_l + g_tl)};
    for (int qd = 0; qd < 4; ++qd) {
      os[8 + 5 * qd + 0] = n_q[qd]; os[8 + 5 * qd + 1] = sp_q[qd]; os[8 + 5 * qd + 2] = sp2_q[qd];
      os[8 + 5 * qd + 3] = sg_q[qd]; os[8 + 5 * qd + 4] = spg_q[qd];
    }
    for (int i = 28; i < SMK_MSUM_STRIDE; ++i) os[i] = 0;
  }
}

__global__ void upsample_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w, int up,
                                         int H, int W) {
  // grid (ceil(H/8), n); 256 threads = 8 rows of one plane per CTA, lanes along x
  const int64_t n = blockIdx.y;
  const int y = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (y >= H) return;
  const float rscale = 1.0f / (float)up;
  RowTap r = row_tap(in + n * h * w, y, h, w, rscale);
  float* o = out + (n * H + y) * (int64_t)W;
  for (int x = threadIdx.x & 31; x < W; x += 32) o[x] = pixel(r, x, w, rscale);
}

}  // namespace smk

using namespace smk;

extern "C" int smk_eval_batch(const float* mask_pred, int64_t batch_stride, const float* objectness, int64_t obj_stride,
                              const uint8_t* gt, int B, int nq, int hp, int wp, int up, int H, int W,
                              int32_t* q_counts, int32_t* idx, int32_t* m_counts, double* m_sums, void* stream) {
  SMK_REQUIRE(mask_pred && objectness && gt && q_counts && idx && m_counts && m_sums, "smk_eval_batch: null pointer");
  SMK_REQUIRE(B >= 0 && nq > 0 && hp > 0 && wp > 0 && up >= 1, "smk_eval_batch: bad sizes");
  SMK_REQUIRE(H > 0 && W > 0 && H <= hp * up && W <= wp * up, "smk_eval_batch: gt %dx%d larger than masks %dx%d x%d", H, W, hp, wp, up);
  SMK_REQUIRE(B <= 65535, "smk_eval_batch: B > 65535");
  if (B == 0) return SMK_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t plane_bytes = (size_t)hp * wp * sizeof(float);
  SMK_REQUIRE(plane_bytes <= 160 * 1024, "smk_eval_batch: mask plane %dx%d does not fit shared memory", hp, wp);
  if (plane_bytes > 24 * 1024) {   // static (17.5 KB) + dynamic shared memory beyond the 48 KB default needs the opt-in
    SMK_CHECK_CUDA(cudaFuncSetAttribute(query_iou_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_bytes));
    SMK_CHECK_CUDA(cudaFuncSetAttribute(mask_metrics_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_bytes));
  }
  const bool x4 = up == 4 && (W % 4) == 0 && W <= 4 * kEvalThreads && ((uintptr_t)gt % 4) == 0;
  if (x4 && plane_bytes > 24 * 1024)
    SMK_CHECK_CUDA(cudaFuncSetAttribute(query_iou_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_bytes));
  // x4 metrics kernel: plane + bit-packed GT + cell bit rows + mixed-cell list (static: 16 KB of histograms + reductions)
  const size_t mm_bytes = plane_bytes + (size_t)H * ((W + 31) / 32) * 4 + (2 * (size_t)hp * ((wp + 31) / 32) + 2 * (size_t)(hp + 1) * ((wp + 32) / 32)) * 4 +
                          (size_t)(hp + 1) * (wp + 1) * 2;
  const bool x4m = x4 && wp <= 255 && hp <= 255 && mm_bytes <= 180 * 1024;
  if (x4m && mm_bytes > 24 * 1024)
    SMK_CHECK_CUDA(cudaFuncSetAttribute(mask_metrics_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mm_bytes));
  // cell-classified IoU kernel (one warp per query): bit-packed GT + per-warp bit rows and mixed-cell lists in shared memory
  const int n_cells = (hp + 1) * (wp + 1);
  const size_t cells_bytes = (size_t)H * ((W + 31) / 32) * 4 +
                             (size_t)kEvalWarps * (((size_t)hp * ((wp + 31) / 32) + 2 * (size_t)(hp + 1) * ((wp + 32) / 32)) * 4 + (size_t)n_cells * 2);
  static const bool cells_on = !(getenv("SMK_EVAL_CELLS") && atoi(getenv("SMK_EVAL_CELLS")) == 0);   // A/B switch (tuning)
  const bool cells = x4 && cells_on && wp <= 32 * kCellMaxWords - 1 && hp <= 255 && cells_bytes <= 200 * 1024;
  if (cells && cells_bytes > 40 * 1024)
    SMK_CHECK_CUDA(cudaFuncSetAttribute(query_iou_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cells_bytes));
  {
    // algorithmic bytes "as if materialised" (SURVEY.md §8d): every full-resolution mask pixel (fp32) + the GT plane
    ProfScope prof(PROF_EVAL, (double)B * ((double)nq * H * W * 4.0 + (double)H * W), s);
    if (cells)
      SMK_CHECK_CUDA(launch_pdl(query_iou_cells_kernel, dim3((nq + kEvalWarps - 1) / kEvalWarps, B), dim3(kEvalThreads), cells_bytes, s, mask_pred,
                                batch_stride, gt, nq, hp, wp, H, W, q_counts));
    else if (x4)
      SMK_CHECK_CUDA(launch_pdl(query_iou_x4_kernel, dim3(nq, B), dim3(kEvalThreads), (size_t)plane_bytes, s, mask_pred, batch_stride, gt, nq, hp, wp, H, W,
                                q_counts));
    else
      query_iou_kernel<<<dim3(nq, B), kEvalThreads, plane_bytes, s>>>(mask_pred, batch_stride, gt, nq, hp, wp, up, H, W, q_counts);
  }
  SMK_CHECK_LAUNCH();
  {
    ProfScope prof(PROF_EVAL, (double)B * 2.0 * ((double)H * W * 4.0 + (double)H * W), s);
    if (x4m)
      SMK_CHECK_CUDA(launch_pdl(mask_metrics_x4_kernel, dim3(2, B), dim3(kEvalThreads), mm_bytes, s, mask_pred, batch_stride, objectness,
                                obj_stride, q_counts, gt, nq, hp, wp, H, W, idx, m_counts, m_sums));
    else
      mask_metrics_kernel<true, true><<<dim3(2, B), kEvalThreads, plane_bytes, s>>>(
          mask_pred, batch_stride, objectness, obj_stride, q_counts, gt, nq, hp, wp, up, H, W, idx, m_counts, m_sums);
  }
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

extern "C" int smk_mask_metrics(const float* pred, const uint8_t* gt, int n, int H, int W, int32_t* m_counts,
                                double* m_sums, void* stream) {
  SMK_REQUIRE(pred && gt && m_counts && m_sums, "smk_mask_metrics: null pointer");
  SMK_REQUIRE(n >= 0 && n <= 65535 && H > 0 && W > 0, "smk_mask_metrics: bad sizes");
  if (n == 0) return SMK_OK;
  // one full-resolution plane per "image"; the plane is read from global memory (L1/L2), up = 1
  mask_metrics_kernel<false, false><<<dim3(1, n), kEvalThreads, 0, (cudaStream_t)stream>>>(
      pred, (int64_t)H * W, nullptr, 0, nullptr, gt, 1, H, W, 1, H, W, nullptr, m_counts, m_sums);
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}

extern "C" int smk_upsample_bilinear(const float* in, float* out, int64_t n, int h, int w, int scale, int H, int W,
                                     void* stream) {
  SMK_REQUIRE(in && out, "smk_upsample_bilinear: null pointer");
  SMK_REQUIRE(n >= 0 && n <= 65535 && h > 0 && w > 0 && scale >= 1 && H > 0 && W > 0 && H <= h * scale && W <= w * scale,
              "smk_upsample_bilinear: bad sizes");
  if (n == 0) return SMK_OK;
  upsample_bilinear_kernel<<<dim3((H + 7) / 8, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(in, out, h, w, scale, H, W);
  SMK_CHECK_LAUNCH();
  return SMK_OK;
}
