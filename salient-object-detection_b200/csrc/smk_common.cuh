// Shared helpers for the selfmask_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/selfmask_b200.h"

namespace smk {

void set_error(const char* fmt, ...);
void count_launch();

// Optional per-category device timing (CUDA events on the launching stream), for bench.py's roofline block.
enum ProfCat { PROF_GEMM_TC = 0, PROF_ATTENTION = 1, PROF_GEMM_F32 = 2, PROF_LAYERNORM = 3, PROF_EVAL = 4, PROF_MASK_HEAD = 5,
               PROF_OTHER = 6, PROF_ATTENTION_TC = 7, PROF_NUM = 8 };
struct ProfScope {
  int slot;
  cudaStream_t stream;
  // work = ALGORITHMIC FLOPs (tensor categories) or bytes (memory ones); issued = what the kernel really executes when that differs
  // (split-operand GEMMs issue 2-3 tensor-core terms per algorithmic FLOP), 0 = same as work
  ProfScope(int cat, double work, cudaStream_t s, double issued = 0.0);
  ~ProfScope();
};
// Sub-category of the next launches for the per-kernel roofline rows of bench.py (host side, set by the model orchestration)
enum ProfTag { TAG_NONE = 0, TAG_PATCH_EMBED, TAG_QKV, TAG_ATTN, TAG_PROJ, TAG_FC1, TAG_FC2, TAG_LN, TAG_KV, TAG_DEC_GEMM, TAG_DEC_ATTN,
               TAG_DEC_LN, TAG_MASK_LOGITS, TAG_MASK_UPSAMPLE, TAG_OBJECTNESS, TAG_EVAL_IOU, TAG_EVAL_METRICS, TAG_IM2COL, TAG_NUM };
extern thread_local int g_prof_tag;
struct TagScope {
  int prev;
  explicit TagScope(int t) : prev(g_prof_tag) { g_prof_tag = t; }
  ~TagScope() { g_prof_tag = prev; }
};

// Per-device one-time state.  The library is used by one host thread per GPU, but a process may touch several devices in turn
// (tests, the reference-style single-process callers): anything a kernel needs set once (cudaFuncSetAttribute, SM count) is keyed
// by the current device ordinal instead of a process-wide flag.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d >= 0 && d < kMaxDevices ? d : 0;
}
struct DeviceOnce {
  bool done[kMaxDevices] = {};
  bool first() {                       // true exactly once per device
    const int d = current_device();
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
int device_sm_count();                 // multiprocessors of the current device (cached per device)

#define SMK_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      smk::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SMK_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define SMK_CHECK_LAUNCH()       \
  do {                           \
    smk::count_launch();         \
    SMK_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

#define SMK_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      smk::set_error(__VA_ARGS__);  \
      return SMK_ERR_INVALID;       \
    }                               \
  } while (0)

#define SMK_PROPAGATE(expr)   \
  do {                        \
    int _s = (expr);          \
    if (_s != SMK_OK) return _s; \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// A step is ~170 small-to-medium kernels in one stream.  Launched with the programmatic-stream-serialization attribute, a
// kernel's CTAs may be scheduled while the previous kernel's last CTAs are still running: launch latency, barrier / TMEM
// set-up and tensor-map prefetch overlap that tail.  Contract: a kernel launched through launch_pdl() executes
// pdl_wait() before its first global-memory access (it returns once the previous grid has completed and flushed), and
// pdl_trigger() right after (lets its own successor be scheduled as soon as SM resources free up).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // SMK_PDL=0 switches the attribute off (A/B measurements)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }

// 16-bit operand types of the tensor-core paths: pack two floats (round to nearest even) / unpack, and the 2-term split
// x ≈ hi + lo with hi = T(x), lo = T(x − hi)  (bf16: ~16 significant bits, fp16: ~22)
template <typename T> struct Pack16;
template <> struct Pack16<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }
};
template <> struct Pack16<__half> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    const __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
};
template <typename T>
__device__ __forceinline__ void split16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = Pack16<T>::pack(a, b);
  const float2 hf = Pack16<T>::unpack(hi);
  lo = Pack16<T>::pack(a - hf.x, b - hf.y);
}

// ---- fp8 correction operands of the fp16s GEMMs (smk_gemm_tc.cu TcGemmParams::q8) -------------------------------------------
// A split row is [hi fp16 (K) | 2K bytes of e4m3]: per 32 operand columns, 32 bytes "first" then 32 bytes "second" —
// activations: first = e4m3(hi), second = e4m3(lo·2^11); weights: first = e4m3(lo·2^15), second = e4m3(hi·2^4) — so that the fp8
// contraction over the 64 bytes is (hi·lo + lo·hi)·2^15 for those 32 columns.  Conversions saturate (graceful: the terms are
// corrections of relative size 2^-11).
constexpr float kQ8ActLo = 2048.f, kQ8WLo = 32768.f, kQ8WHi = 16.f;
__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {     // memory order a, b, c, d
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  return (uint32_t)lo | ((uint32_t)hi << 16);
}
__device__ __forceinline__ int q8_byte_off(int col) { return ((col >> 5) << 6) + (col & 31); }   // of `first`; `second` is 32 bytes further
// four consecutive values → fp16 hi pairs + the two e4m3 words
template <bool kWeight>
__device__ __forceinline__ void split_q8x4(float a, float b, float c, float d, uint2& hi16, uint32_t& first8, uint32_t& second8) {
  hi16.x = Pack16<__half>::pack(a, b);
  hi16.y = Pack16<__half>::pack(c, d);
  const float2 h0 = Pack16<__half>::unpack(hi16.x), h1 = Pack16<__half>::unpack(hi16.y);
  if constexpr (kWeight) {
    first8 = e4m3x4((a - h0.x) * kQ8WLo, (b - h0.y) * kQ8WLo, (c - h1.x) * kQ8WLo, (d - h1.y) * kQ8WLo);
    second8 = e4m3x4(h0.x * kQ8WHi, h0.y * kQ8WHi, h1.x * kQ8WHi, h1.y * kQ8WHi);
  } else {
    first8 = e4m3x4(h0.x, h0.y, h1.x, h1.y);
    second8 = e4m3x4((a - h0.x) * kQ8ActLo, (b - h0.y) * kQ8ActLo, (c - h1.x) * kQ8ActLo, (d - h1.y) * kQ8ActLo);
  }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ATen bilinear source index (align_corners=False): src = max((dst+0.5)/scale - 0.5, 0)
struct Tap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Tap make_tap(int dst, float rscale, int n) {
  float s = fmaxf(__fmaf_rn((float)dst + 0.5f, rscale, -0.5f), 0.0f);
  Tap t;
  t.i0 = min((int)s, n - 1);
  t.i1 = min(t.i0 + 1, n - 1);
  t.l1 = s - (float)t.i0;
  t.l0 = 1.0f - t.l1;
  return t;
}
// "nested" rounding = ATen for the evaluator's shapes (oracle/selfmask_oracle.py upsample_bilinear)
__device__ __forceinline__ float bilerp(float a, float b, float c, float d, float lx0, float lx1, float ly0, float ly1) {
  float top = __fmaf_rn(a, lx0, __fmul_rn(b, lx1));
  float bot = __fmaf_rn(c, lx0, __fmul_rn(d, lx1));
  return __fmaf_rn(top, ly0, __fmul_rn(bot, ly1));
}

// x4 specialisation (the evaluator's scale_factor=4 and the shipped pixel decoder): the 4 outputs x = 4i .. 4i+3 of one row
// only read source columns {max(i-1,0), i, min(i+1,w-1)} — outputs 0,1 blend (left, centre), outputs 2,3 blend (centre,
// right).  Where make_tap() clamps (first / last source column) the clamped pair holds the same value twice, so the
// result is bit-identical to the generic tap (a*l0 + a*l1 paths only differ when the weights are exactly 1 and 0).
struct QuadX {
  int cm, cc, cp;
  float l0[4], l1[4];
};
__device__ __forceinline__ QuadX make_quadx(int i, int w) {
  QuadX q;
  q.cm = max(i - 1, 0);
  q.cc = min(i, w - 1);
  q.cp = min(i + 1, w - 1);
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const Tap t = make_tap(4 * i + f, 0.25f, w);
    q.l0[f] = t.l0;
    q.l1[f] = t.l1;
  }
  return q;
}
__device__ __forceinline__ void quad4(const float* __restrict__ r0, const float* __restrict__ r1, const QuadX& q, float ly0, float ly1,
                                      float (&out)[4]) {
  const float am = r0[q.cm], ac = r0[q.cc], ap = r0[q.cp];
  const float bm = r1[q.cm], bc = r1[q.cc], bp = r1[q.cp];
  out[0] = bilerp(am, ac, bm, bc, q.l0[0], q.l1[0], ly0, ly1);
  out[1] = bilerp(am, ac, bm, bc, q.l0[1], q.l1[1], ly0, ly1);
  out[2] = bilerp(ac, ap, bc, bp, q.l0[2], q.l1[2], ly0, ly1);
  out[3] = bilerp(ac, ap, bc, bp, q.l0[3], q.l1[3], ly0, ly1);
}
// sigmoid with MUFU.EX2 + MUFU.RCP (≈3 ulp): 1 / (1 + 2^(-x·log2 e))
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace smk
