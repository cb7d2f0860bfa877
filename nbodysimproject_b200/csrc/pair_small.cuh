// pair_small.cuh -- register-resident pair passes for one small system per thread (N = 2..8).
//
// Restates (not ports) geometry_cache.py:24-39 + forces.py:63-75 + simulation.py:551-552 (a1-a3),
// forces.py:77-112 / potential.py:23-64 (a4, a5) and tangent_map.py:21-59 (a6) for a thread that holds
// the whole system in registers.  Every unordered pair is visited once (Newton's third law), the
// loops are fully unrolled, and body i receives its j-terms in ascending j like the reference's
// axis-1 sum.
#pragma once
#include "common.cuh"

namespace nb {

template <int N>
struct SysState {
  double gm[N];          // G * m_i
  double x[N], y[N];
  double vx[N], vy[N];
  double ax[N], ay[N];   // acceleration at the current positions (FSAL)
  double eps2;
};

// acceleration (and optionally the variational acceleration for the tangent vector dr) at the
// current positions.  14 algorithmic flops per ordered pair (+19 for the tangent term).
template <int N, bool TANGENT, bool GUARD>
__device__ __forceinline__ void pair_pass(SysState<N>& s, const double* __restrict__ drx,
                                          const double* __restrict__ dry, double* __restrict__ dax,
                                          double* __restrict__ day) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.ax[i] = 0.0;
    s.ay[i] = 0.0;
    if (TANGENT) {
      dax[i] = 0.0;
      day[i] = 0.0;
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      const double dx = s.x[i] - s.x[j];
      const double dy = s.y[i] - s.y[j];
      const double r2 = fma(dx, dx, fma(dy, dy, s.eps2));
      double w2, w3;
      if (TANGENT) {
        const double w = rsqrt_f64<GUARD>(r2);
        w2 = w * w;
        w3 = w2 * w;
      } else {
        w2 = 0.0;
        w3 = rsqrt3_f64<GUARD>(r2);      // rho^-3 directly: 16 instead of 17 FP64 operations per unordered pair
      }
      const double cj = s.gm[j] * w3;
      const double ci = s.gm[i] * w3;
      s.ax[i] = fma(-cj, dx, s.ax[i]);
      s.ay[i] = fma(-cj, dy, s.ay[i]);
      s.ax[j] = fma(ci, dx, s.ax[j]);
      s.ay[j] = fma(ci, dy, s.ay[j]);
      if (TANGENT) {
        // D = q_j - q_i = -(dx,dy); d = dr_j - dr_i; term = d w^3 - 3 (D.d) w^5 D
        const double ex = drx[j] - drx[i];
        const double ey = dry[j] - dry[i];
        const double dot = -fma(dx, ex, dy * ey);
        const double c5 = 3.0 * dot * w2 * w3;
        const double tx = fma(ex, w3, c5 * dx);   // -c5 * D_x = +c5 * dx
        const double ty = fma(ey, w3, c5 * dy);
        dax[i] = fma(s.gm[j], tx, dax[i]);
        day[i] = fma(s.gm[j], ty, day[i]);
        dax[j] = fma(-s.gm[i], tx, dax[j]);
        day[j] = fma(-s.gm[i], ty, day[j]);
      }
    }
  }
}

// U = -G sum_{i<j} m_i m_j / rho  and  S3 = G sum_{i<j} m_i m_j / rho^3   (dV/deps = eps * S3)
template <int N, bool GUARD>
__device__ __forceinline__ void pair_scalars(const double* gm, const double* m, const double* x,
                                             const double* y, double eps2, double& U, double& S3) {
  double u = 0.0, s3 = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      const double dx = x[i] - x[j];
      const double dy = y[i] - y[j];
      const double r2 = fma(dx, dx, fma(dy, dy, eps2));
      const double w = rsqrt_f64<GUARD>(r2);
      const double mm = gm[i] * m[j];
      u = fma(mm, w, u);
      s3 = fma(mm, w * w * w, s3);
    }
  }
  U = -u;
  S3 = s3;
}

}  // namespace nb
