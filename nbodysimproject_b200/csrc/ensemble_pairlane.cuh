// ensemble_pairlane.cuh -- shortest-critical-path mapping for the sub-step-heavy systems with N >= 5:
// one unordered PAIR per lane, floor(32 / P) systems per warp (P = N(N-1)/2).
//
// A system's sub-steps are sequential, so the run time of a launch is bounded below by
//   (sub-steps of the heaviest system) x (latency of one force evaluation).
// The lane-per-body mapping (ensemble_group.cuh) still issues ~14 N FP64 instructions per evaluation and
// warp (N = 8: 112, i.e. >= 224 FP64-pipe cycles, more when the SMSP is shared with bulk warps).  Here a lane
// evaluates ONE pair (13 FP64 instructions), publishes the mass-free pair vector (w^3 dx, w^3 dy) in shared
// memory, and after one __syncwarp every lane sums the accelerations of ITS OWN two bodies from the N-1 pair
// vectors each needs (4 (N-1) DFMA).  Each lane carries the full state of its two bodies and integrates them
// redundantly, so positions never have to be exchanged: one shared-memory round trip per force evaluation.
// ~53 FP64 warp-instructions per evaluation at N = 8 instead of 112, and a dependency chain of ~200 cycles.
//
// Arithmetic: the same formulas as pair_small.cuh; body i still receives its terms in ascending j
// (geometry_cache.py:24-39 axis-1 order) but as fma(-+G m_j, w^3 dx, a) instead of fma(-(G m_j w^3), dx, a):
// a last-bit difference (<= 1e-16 relative per evaluation), inside every stated tolerance.
// Used by phase 0 (the main loop, 95 % of a run); the MEGNO phase keeps the lane-per-body mapping.
#pragma once
#include "pair_small.cuh"
#include "args.cuh"

namespace nb {

template <int N>
struct PairLane {
  static constexpr int P = N * (N - 1) / 2;
  static constexpr int S = 32 / P;                 // systems per warp
  static constexpr int WARPS = 4;                  // warps per CTA (128 threads)
  static constexpr int SMEM_DOUBLES = WARPS * (2 * S * P * 2 + S * N * 3);
};

template <int N>
static inline int pairlane_blocks_for(int B) {
  const int warps = (B + PairLane<N>::S - 1) / PairLane<N>::S;
  return (warps + PairLane<N>::WARPS - 1) / PairLane<N>::WARPS;
}

__device__ __forceinline__ int pl_pair_index(int N, int a, int b) {   // a < b
  return a * (2 * N - a - 1) / 2 + (b - a - 1);
}

template <int N, int MODE, bool GUARD>
__device__ __forceinline__ void pairlane_main(const RunArgs& a, int write_state, double* smem, int bid) {
  constexpr int P = PairLane<N>::P, S = PairLane<N>::S;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (bid * (int)blockDim.x + (int)threadIdx.x) >> 5;
  const int n_heavy = min(*a.n_heavy, a.B);
  if (warp * S >= n_heavy) return;                 // warp-uniform exit
  // spare lanes and empty slots shadow a live (system, pair): they compute and publish identical values and
  // never write to global memory, so the warp stays converged
  int g = lane / P;
  int p = lane - g * P;
  bool live = g < S;
  if (!live) { g = 0; p = 0; }
  if (warp * S + g >= n_heavy) { g = 0; live = false; }
  const int slot = warp * S + g;
  const int sys = a.perm[slot];
  int bi = 0, bj;
  {
    int rem = p;                                   // p = bi (2N - bi - 1)/2 + (bj - bi - 1)
    while (rem >= N - 1 - bi) { rem -= N - 1 - bi; ++bi; }
    bj = bi + 1 + rem;
  }
  double2* pbuf = reinterpret_cast<double2*>(smem) + (size_t)wib * (2 * S * P) + g * P;   // [parity][S][P]
  double* sbuf = smem + (size_t)PairLane<N>::WARPS * (2 * S * P * 2) + ((size_t)wib * S + g) * (N * 3);

  // per-lane summation tables: partner k of body i in ascending order, its pair slot and signed G m_k
  double cfi[N - 1], cfj[N - 1];
  int ofi[N - 1], ofj[N - 1];
#pragma unroll
  for (int t = 0; t < N - 1; ++t) {
    const int ki = t < bi ? t : t + 1;
    const int kj = t < bj ? t : t + 1;
    const double gmi = a.G * a.m[(size_t)sys * N + ki];
    const double gmj = a.G * a.m[(size_t)sys * N + kj];
    cfi[t] = bi < ki ? -gmi : gmi;
    cfj[t] = bj < kj ? -gmj : gmj;
    ofi[t] = bi < ki ? pl_pair_index(N, bi, ki) : pl_pair_index(N, ki, bi);
    ofj[t] = bj < kj ? pl_pair_index(N, bj, kj) : pl_pair_index(N, kj, bj);
  }
  const double mi = a.m[(size_t)sys * N + bi], mj = a.m[(size_t)sys * N + bj];
  double xi = a.q[((size_t)sys * N + bi) * 2 + 0], yi = a.q[((size_t)sys * N + bi) * 2 + 1];
  double xj = a.q[((size_t)sys * N + bj) * 2 + 0], yj = a.q[((size_t)sys * N + bj) * 2 + 1];
  double ui = a.v[((size_t)sys * N + bi) * 2 + 0], wi = a.v[((size_t)sys * N + bi) * 2 + 1];
  double uj = a.v[((size_t)sys * N + bj) * 2 + 0], wj = a.v[((size_t)sys * N + bj) * 2 + 1];
  const double eps = a.eps[sys];
  const double eps2 = eps * eps;
  const int n_sub = max(1, a.n_sub[sys]);
  const int n_sub_warp = __reduce_max_sync(0xffffffffu, n_sub);
  const double h = a.dt / (double)n_sub;
  double axi, ayi, axj, ayj;
  int par = 0;

  auto accel = [&](double pxi, double pyi, double pxj, double pyj, double& oxi, double& oyi, double& oxj, double& oyj) {
    const double dx = pxi - pxj, dy = pyi - pyj;
    const double r2 = fma(dx, dx, fma(dy, dy, eps2));
    const double w = rsqrt_f64<GUARD>(r2);
    const double w3 = w * w * w;
    double2* buf = pbuf + par * (S * P);
    buf[p] = make_double2(w3 * dx, w3 * dy);
    par ^= 1;
    __syncwarp();
    double sxi = 0.0, syi = 0.0, sxj = 0.0, syj = 0.0;
#pragma unroll
    for (int t = 0; t < N - 1; ++t) {
      const double2 fi = buf[ofi[t]];
      const double2 fj = buf[ofj[t]];
      sxi = fma(cfi[t], fi.x, sxi);
      syi = fma(cfi[t], fi.y, syi);
      sxj = fma(cfj[t], fj.x, sxj);
      syj = fma(cfj[t], fj.y, syj);
    }
    oxi = sxi; oyi = syi; oxj = sxj; oyj = syj;
  };
  // one velocity-Verlet kernel of size hh on both bodies, FSAL (integration_scheme_base.py:129-149)
  auto vv = [&](double hh, double& pxi, double& pyi, double& pui, double& pwi, double& pxj, double& pyj, double& puj,
                double& pwj, double& qxi, double& qyi, double& qxj, double& qyj) {
    const double h2 = 0.5 * hh;
    pui = fma(h2, qxi, pui); pwi = fma(h2, qyi, pwi);
    puj = fma(h2, qxj, puj); pwj = fma(h2, qyj, pwj);
    pxi = fma(hh, pui, pxi); pyi = fma(hh, pwi, pyi);
    pxj = fma(hh, puj, pxj); pyj = fma(hh, pwj, pyj);
    accel(pxi, pyi, pxj, pyj, qxi, qyi, qxj, qyj);
    pui = fma(h2, qxi, pui); pwi = fma(h2, qyi, pwi);
    puj = fma(h2, qxj, puj); pwj = fma(h2, qyj, pwj);
  };

  accel(xi, yi, xj, yj, axi, ayi, axj, ayj);
  double* f = a.dyn ? a.dyn + (size_t)sys * NB_N_DYN : nullptr;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  // body ownership for per-body reductions and the final store: pair (b, b+1) owns b; pair (N-2, N-1) also owns N-1
  const bool own_i = (bj == bi + 1);
  const bool own_j = (bi == N - 2);

  double com_sum = 0.0, com_max = -1.0, var_sum = 0.0, var_max = -1.0, cos_sum = 0.0, cos_min = 2.0, th_sum = 0.0;
  double Lfirst = 0.0;
  bool have_first = false, cos_nan = false;
  int n_samp = 0, next_sample = 0;
  const int interval = a.sample_interval;
  const double theta_eps = (eps != 0.0) ? atan2(0.0, eps) : nan;   // diagnostics.py:246-249 with pi = 0: loop-invariant
  const double cbrt2 = 1.2599210498948731648;
  const double ha = (1.0 / (2.0 - cbrt2)) * h, hb = (-cbrt2 / (2.0 - cbrt2)) * h;
  for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll 1
    for (int k = 0; k < n_sub_warp; ++k) {
      double txi = xi, tyi = yi, tui = ui, twi = wi, txj = xj, tyj = yj, tuj = uj, twj = wj;
      double bxi = axi, byi = ayi, bxj = axj, byj = ayj;
      if (MODE == NB_MODE_VERLET) {
        vv(h, txi, tyi, tui, twi, txj, tyj, tuj, twj, bxi, byi, bxj, byj);
      } else {
        vv(ha, txi, tyi, tui, twi, txj, tyj, tuj, twj, bxi, byi, bxj, byj);
        vv(hb, txi, tyi, tui, twi, txj, tyj, tuj, twj, bxi, byi, bxj, byj);
        vv(ha, txi, tyi, tui, twi, txj, tyj, tuj, twj, bxi, byi, bxj, byj);
      }
      if (k < n_sub) {
        xi = txi; yi = tyi; ui = tui; wi = twi; xj = txj; yj = tyj; uj = tuj; wj = twj;
        axi = bxi; ayi = byi; axj = bxj; ayj = byj;
      }
    }
    if (interval > 0 && step == next_sample) {       // diagnostics.py:241-285
      next_sample += interval;
      __syncwarp();
      if (own_i) { sbuf[bi * 3 + 0] = mi * xi; sbuf[bi * 3 + 1] = mi * yi; sbuf[bi * 3 + 2] = mi * (xi * wi - yi * ui); }
      if (own_j) { sbuf[bj * 3 + 0] = mj * xj; sbuf[bj * 3 + 1] = mj * yj; sbuf[bj * 3 + 2] = mj * (xj * wj - yj * uj); }
      __syncwarp();
      double cx = 0.0, cy = 0.0, Lt = 0.0;
#pragma unroll
      for (int b = 0; b < N; ++b) { cx += sbuf[b * 3 + 0]; cy += sbuf[b * 3 + 1]; Lt += sbuf[b * 3 + 2]; }
      const double mean = Lt / N;
      double var = 0.0;
#pragma unroll
      for (int b = 0; b < N; ++b) { const double d = sbuf[b * 3 + 2] - mean; var += d * d; }
      var /= N;
      const double com = sqrt(cx * cx + cy * cy);
      if (!have_first) { Lfirst = Lt; have_first = true; }
      double cc;
      if (Lfirst != 0.0 && Lt != 0.0) cc = (Lt * Lfirst) / (fabs(Lt) * fabs(Lfirst));
      else { cc = 0.0; cos_nan = true; }
      com_sum += com; com_max = fmax(com_max, com);
      var_sum += var; var_max = fmax(var_max, var);
      cos_sum += cc; cos_min = fmin(cos_min, cc);
      th_sum += theta_eps;
      ++n_samp;
    }
  }
  if (f && p == 0 && live) {
    const double inv = n_samp > 0 ? 1.0 / (double)n_samp : nan;
    f[NB_F_COM_MEAN] = n_samp > 0 ? com_sum * inv : nan;
    f[NB_F_COM_MAX] = n_samp > 0 ? com_max : nan;
    f[NB_F_JEPS_MEAN] = n_samp > 0 ? 0.0 : nan;
    f[NB_F_JEPS_STD] = n_samp > 0 ? 0.0 : nan;
    f[NB_F_THETA_MEAN] = n_samp > 0 ? th_sum * inv : nan;
    f[NB_F_THETA_STD] = n_samp > 0 ? ((eps != 0.0) ? 0.0 : nan) : nan;
    f[NB_F_COS_MEAN] = (n_samp > 0 && !cos_nan) ? cos_sum * inv : nan;
    f[NB_F_COS_MIN] = (n_samp > 0 && !cos_nan) ? cos_min : nan;
    f[NB_F_VARL_MEAN] = n_samp > 0 ? var_sum * inv : nan;
    f[NB_F_VARL_MAX] = n_samp > 0 ? var_max : nan;
    f[NB_F_TIDAL_MEAN] = n_samp > 0 ? 0.0 : nan;
    f[NB_F_TIDAL_MAX] = n_samp > 0 ? 0.0 : nan;
  }
  const bool fin = is_finite(xi) && is_finite(yi) && is_finite(ui) && is_finite(wi) && is_finite(xj) && is_finite(yj) &&
                   is_finite(uj) && is_finite(wj);
  const unsigned grp = (P >= 32 ? 0xffffffffu : ((1u << P) - 1u)) << (g * P);
  const unsigned bad = __ballot_sync(0xffffffffu, !fin) & grp;
  if (write_state && live) {
    if (own_i) {
      a.q[((size_t)sys * N + bi) * 2 + 0] = xi; a.q[((size_t)sys * N + bi) * 2 + 1] = yi;
      a.v[((size_t)sys * N + bi) * 2 + 0] = ui; a.v[((size_t)sys * N + bi) * 2 + 1] = wi;
    }
    if (own_j) {
      a.q[((size_t)sys * N + bj) * 2 + 0] = xj; a.q[((size_t)sys * N + bj) * 2 + 1] = yj;
      a.v[((size_t)sys * N + bj) * 2 + 0] = uj; a.v[((size_t)sys * N + bj) * 2 + 1] = wj;
    }
  }
  if (a.status && p == 0 && live) a.status[sys] = bad ? NB_STATUS_NONFINITE : 0;
}

}  // namespace nb
