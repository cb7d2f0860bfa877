// common.cuh -- shared device helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/nbody_b200.h"
#ifndef __CUDA_ARCH__
#include <nvtx3/nvToolsExt.h>     // header-only NVTX v3: ranges cost ~nothing unless a profiler is attached
#endif

namespace nb {

// NVTX range around the host-side launch sequence of one phase (prepare / main / MEGNO / finalize / large-N passes),
// so that an nsys / ncu timeline shows the reference's call structure
struct NvtxRange {
  explicit NvtxRange(const char* name) {
#ifndef __CUDA_ARCH__
    nvtxRangePushA(name);
#endif
  }
  ~NvtxRange() {
#ifndef __CUDA_ARCH__
    nvtxRangePop();
#endif
  }
};

void set_error(const char* msg);
int cuda_fail(cudaError_t e, const char* where);

#define NB_CUDA_CHECK(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return nb::cuda_fail(_e, #expr);   \
  } while (0)

// ------------------------------------------------------------------------------------------------
// fp64 reciprocal square root: MUFU.RSQ64H seed (rel. err ~2^-22) + one cubic (Halley) correction
//   e = 1 - x y0^2 ;  y = y0 (1 + e/2 + 3 e^2/8)      -> rel. err ~ (5/16) e^3 < 2^-60
// 4 DP-pipe ops + 1 MUFU.  x <= 0 / denormal -> 0, matching the reference's masked inv_r3
// (geometry_cache.py:33-36); NaN propagates.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt_seed(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  return y0;
}

template <bool GUARD>
__device__ __forceinline__ double rsqrt_f64(double x) {
  double y0 = rsqrt_seed(x);
  double t = x * y0;
  double e = fma(-t, y0, 1.0);
  double p = fma(0.375, e, 0.5);
  double y = fma(y0 * e, p, y0);
  if (GUARD) {
    // x < 2^-1022 (zero, denormal or negative): the reference leaves inv_r3 = 0 there
    if (__double2hiint(x) < 0x00100000) y = 0.0;
  }
  return y;
}

// x^(-3/2) directly from the same seed, one FP64 operation fewer than rsqrt_f64 followed by two multiplies:
//   u = y0^2 ; e = 1 - x u (one fma) ; x^(-3/2) = y0^3 (1 - e)^(-3/2) = y0^3 (1 + 3e/2 + 15 e^2/8 + O(e^3))
// 6 DP-pipe ops + 1 MUFU; relative error ~ (35/16) e^3 + 3 ulp with |e| <~ 2^-21.  Same x <= 0 / denormal -> 0 rule.
template <bool GUARD>
__device__ __forceinline__ double rsqrt3_f64(double x) {
  const double y0 = rsqrt_seed(x);
  const double u = y0 * y0;
  const double e = fma(-x, u, 1.0);
  const double y03 = u * y0;
  const double p = fma(1.875, e, 1.5);
  double w3 = fma(y03 * e, p, y03);
  if (GUARD) {
    if (__double2hiint(x) < 0x00100000) w3 = 0.0;
  }
  return w3;
}

// ------------------------------------------------------------------------------------------------
// double-double arithmetic (only used for the E0/E1 energy reductions that the reference does in
// long double + Kahan, diagnostics.py:457-549)
// ------------------------------------------------------------------------------------------------
struct dd {
  double hi, lo;
};
__device__ __forceinline__ dd dd_make(double a) { return dd{a, 0.0}; }
__device__ __forceinline__ dd two_sum(double a, double b) {
  double s = __dadd_rn(a, b);
  double bb = __dadd_rn(s, -a);
  double e = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
  return dd{s, e};
}
__device__ __forceinline__ dd quick_two_sum(double a, double b) {
  double s = __dadd_rn(a, b);
  double e = __dadd_rn(b, -__dadd_rn(s, -a));
  return dd{s, e};
}
__device__ __forceinline__ dd two_prod(double a, double b) {
  double p = __dmul_rn(a, b);
  double e = __fma_rn(a, b, -p);
  return dd{p, e};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  dd t = two_sum(a.lo, b.lo);
  s.lo = __dadd_rn(s.lo, t.hi);
  s = quick_two_sum(s.hi, s.lo);
  s.lo = __dadd_rn(s.lo, t.lo);
  return quick_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
__device__ __forceinline__ dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
__device__ __forceinline__ dd dd_add_d(dd a, double b) { return dd_add(a, dd_make(b)); }
__device__ __forceinline__ dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo = __fma_rn(a.hi, b.lo, p.lo);
  p.lo = __fma_rn(a.lo, b.hi, p.lo);
  return quick_two_sum(p.hi, p.lo);
}
__device__ __forceinline__ dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo = __fma_rn(a.lo, b, p.lo);
  return quick_two_sum(p.hi, p.lo);
}
__device__ __forceinline__ dd dd_div(dd a, dd b) {
  double q1 = __ddiv_rn(a.hi, b.hi);
  dd r = dd_sub(a, dd_mul_d(b, q1));
  double q2 = __ddiv_rn(r.hi, b.hi);
  r = dd_sub(r, dd_mul_d(b, q2));
  double q3 = __ddiv_rn(r.hi, b.hi);
  dd q = quick_two_sum(q1, q2);
  return dd_add_d(q, q3);
}
__device__ __forceinline__ dd dd_sqrt(dd a) {
  if (!(a.hi > 0.0)) return dd_make(0.0);
  double x = 1.0 / sqrt(a.hi);
  double ax = __dmul_rn(a.hi, x);
  dd err = dd_sub(a, two_prod(ax, ax));
  return quick_two_sum(ax, __dmul_rn(err.hi, __dmul_rn(x, 0.5)));
}
__device__ __forceinline__ double dd_to_double(dd a) { return __dadd_rn(a.hi, a.lo); }

// Kernel-side time stamps for the roofline figures: the first CTA to start / the last warp to finish publish
// %globaltimer (ns) with one atomic each.  Unlike CUDA events recorded between launches this adds no stream operation
// (timing events on the bucket streams serialised the concurrent launches: +20 % on the C3 step).
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void stamp_begin(unsigned long long* ts) {
  if (ts && threadIdx.x == 0) atomicMin(ts, global_timer_ns());
}
__device__ __forceinline__ void stamp_end(unsigned long long* ts) {
  if (ts && (threadIdx.x & 31) == 0) atomicMax(ts + 1, global_timer_ns());
}

__device__ __forceinline__ bool is_finite(double x) { return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000; }

}  // namespace nb
