// ensemble_run.cuh -- batched ensembles of small 2-D systems, one system per thread, state in registers.
//
//   prepare kernel : construction-time work of NBodySimulation (simulation.py:39-162): COM velocity
//                    removal, corrector half kicks, frozen sub-step schedule (timestep_manager.py:139-253)
//                    and the 25 static features (dynamical_features.py:27-155).
//   run kernel     : StabilityAnalyzer.run_stability_analysis (stability_analyzer.py:69-259) for one
//                    system per thread: n_steps macro steps of verlet / yoshida4 / whfast with the
//                    reference's sub-stepping (integrator.py:78-104), step_metrics sampling
//                    (diagnostics.py:241-285), long-double-equivalent E0/E1 (diagnostics.py:457-549),
//                    then the tangent-map MEGNO loop (evolution_features.py:34-66).
//
// Design notes (B200): the whole working set of a system (<= 56 doubles + 32 for the tangent vectors)
// lives in registers for the entire run, so the kernel is bound by the FP64 pipe, not by HBM; each
// unordered pair is evaluated once; the acceleration at the end of a velocity-Verlet kernel is reused
// as the start acceleration of the next one (bit-identical to the reference's recomputation); the last
// force evaluation of a MEGNO step is fused with the variational acceleration so they share rho^-1.
#pragma once
#include "pair_small.cuh"
#include "kepler.cuh"
#include "args.cuh"
#include "ensemble_group.cuh"
#include "ensemble_pairlane.cuh"

#ifndef NB_WH_MINB
#define NB_WH_MINB 5          // whfast main kernel: resident CTAs per SM the register allocation must allow (N <= NB_WH_MINB_MAXN).
                              // 96 registers + 300-500 B of spills instead of 154-166 registers and none: a kernel alone is 0-6 %
                              // slower, the three concurrent C4 buckets 13 % faster (1.12e9 -> 1.27e9; 3: 1.12, 4: 1.22,
                              // 5 / 6 / 8: 1.27) -- the strictly rounded Newton chain is latency-bound (top stall `wait`)
#endif
#ifndef NB_WH_MINB_MAXN
#define NB_WH_MINB_MAXN 5
#endif


#ifndef NB_Y4_ROLL_MINN
#define NB_Y4_ROLL_MINN 6     // yoshida4 main kernel: body counts from which the three stages run as a rolled loop (one
                              // instance of the unrolled pair loop: C3 step 54.7 -> 53.9 ms; rolling N <= 5 as well: 54.1)
#endif

namespace nb {

// ---------------------------------------------------------------------------------------------
// helpers on a register-resident system
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void kick(SysState<N>& s, double h) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.vx[i] = fma(h, s.ax[i], s.vx[i]);
    s.vy[i] = fma(h, s.ay[i], s.vy[i]);
  }
}
template <int N>
__device__ __forceinline__ void drift(SysState<N>& s, double h) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.x[i] = fma(h, s.vx[i], s.x[i]);
    s.y[i] = fma(h, s.vy[i], s.y[i]);
  }
}

// whfast_scheme.py:22-37 + simulation.py:487-534: pseudo-Jacobi Kepler drift.  `m` are plain masses.
template <int N, bool EXACT>
__device__ __forceinline__ int kepler_drift(SysState<N>& s, const double* m, double G, double tau, int& iters) {
  double jx[N], jy[N], jvx[N], jvy[N];
  int worst = 0;
  {
    double Rx = m[0] * s.x[0], Ry = m[0] * s.y[0], Vx = m[0] * s.vx[0], Vy = m[0] * s.vy[0], M = m[0];
    jx[0] = s.x[0]; jy[0] = s.y[0]; jvx[0] = s.vx[0]; jvy[0] = s.vy[0];
#pragma unroll
    for (int i = 1; i < N; ++i) {
      jx[i] = s.x[i] - Rx / M;
      jy[i] = s.y[i] - Ry / M;
      jvx[i] = s.vx[i] - Vx / M;
      jvy[i] = s.vy[i] - Vy / M;
      Rx = __dadd_rn(Rx, __dmul_rn(m[i], s.x[i]));
      Ry = __dadd_rn(Ry, __dmul_rn(m[i], s.y[i]));
      Vx = __dadd_rn(Vx, __dmul_rn(m[i], s.vx[i]));
      Vy = __dadd_rn(Vy, __dmul_rn(m[i], s.vy[i]));
      M += m[i];
    }
  }
  jx[0] = __dadd_rn(jx[0], __dmul_rn(jvx[0], tau));
  jy[0] = __dadd_rn(jy[0], __dmul_rn(jvy[0], tau));
  {
    double cum = m[0];
#pragma unroll 1
    for (int i = 1; i < N; ++i) {
      const double mu = G * (cum + m[i]);
      cum += m[i];
      double rx = jx[i], ry = jy[i], ux = jvx[i], uy = jvy[i];
      int done = 0;
      const int it = EXACT ? kepler_exact(rx, ry, ux, uy, mu, tau) : kepler_reference(rx, ry, ux, uy, mu, tau, done);
      worst = max(worst, it);
      iters += EXACT ? min(it, 64) : done;
      jx[i] = rx; jy[i] = ry; jvx[i] = ux; jvy[i] = uy;
    }
  }
  {
    s.x[0] = jx[0]; s.y[0] = jy[0]; s.vx[0] = jvx[0]; s.vy[0] = jvy[0];
    double Rx = m[0] * s.x[0], Ry = m[0] * s.y[0], Vx = m[0] * s.vx[0], Vy = m[0] * s.vy[0], M = m[0];
#pragma unroll
    for (int i = 1; i < N; ++i) {
      s.x[i] = jx[i] + Rx / M;
      s.y[i] = jy[i] + Ry / M;
      s.vx[i] = jvx[i] + Vx / M;
      s.vy[i] = jvy[i] + Vy / M;
      Rx = __dadd_rn(Rx, __dmul_rn(m[i], s.x[i]));
      Ry = __dadd_rn(Ry, __dmul_rn(m[i], s.y[i]));
      Vx = __dadd_rn(Vx, __dmul_rn(m[i], s.vx[i]));
      Vy = __dadd_rn(Vy, __dmul_rn(m[i], s.vy[i]));
      M += m[i];
    }
  }
  return worst;
}

// whfast_scheme.py:39-69: the "interaction acceleration" (only live use: the whfast corrector :95-123)
template <int N>
__device__ __forceinline__ void wh_interaction_accel(const SysState<N>& s, const double* m, double G, double* ax,
                                                     double* ay) {
  double cum[N];
  cum[0] = m[0];
#pragma unroll
  for (int i = 1; i < N; ++i) cum[i] = cum[i - 1] + m[i];
  double jx[N], jy[N];
  {
    double Rx = m[0] * s.x[0], Ry = m[0] * s.y[0], M = m[0];
    jx[0] = s.x[0]; jy[0] = s.y[0];
#pragma unroll
    for (int i = 1; i < N; ++i) {
      jx[i] = s.x[i] - Rx / M;
      jy[i] = s.y[i] - Ry / M;
      Rx += m[i] * s.x[i];
      Ry += m[i] * s.y[i];
      M += m[i];
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) { ax[i] = 0.0; ay[i] = 0.0; }
#pragma unroll
  for (int i = 2; i < N; ++i) {
    const double rn2 = jx[i] * jx[i] + jy[i] * jy[i] + s.eps2;
    if (rn2 > 0.0) {
      const double f = G * cum[i - 1] / (rn2 * sqrt(rn2));
      const double gx = f * jx[i], gy = f * jy[i];
#pragma unroll
      for (int k = 0; k < i; ++k) {
        const double w = m[i] * (m[k] / cum[i - 1]);
        ax[k] -= w * gx;
        ay[k] -= w * gy;
      }
      ax[i] += cum[i - 1] * gx;
      ay[i] += cum[i - 1] * gy;
    }
  }
#pragma unroll
  for (int i = 1; i < N; ++i) {
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      const double dx = s.x[j] - s.x[i], dy = s.y[j] - s.y[i];
      const double r2 = dx * dx + dy * dy + s.eps2;
      const double w = G / (r2 * sqrt(r2));
      ax[i] -= m[j] * w * dx;
      ay[i] -= m[j] * w * dy;
      ax[j] += m[i] * w * dx;
      ay[j] += m[i] * w * dy;
    }
  }
}

// one sub-step of size h.  `tan_*` non-null => this is the last sub-step of a MEGNO step and the final
// force evaluation is fused with the variational acceleration (classic modes).
template <int N, int MODE, bool GUARD, bool EXACT, bool TANGENT>
__device__ __forceinline__ int substep(SysState<N>& s, const double* m, double G, double h, const double* drx,
                                       const double* dry, double* dax, double* day, int& iters) {
  int kep = 0;
  if (MODE == NB_MODE_VERLET) {
    const double h2 = 0.5 * h;
    kick<N>(s, h2);
    drift<N>(s, h);
    pair_pass<N, TANGENT, GUARD>(s, drx, dry, dax, day);
    kick<N>(s, h2);
  } else if (MODE == NB_MODE_YOSHIDA4) {
    const double cbrt2 = 1.2599210498948731648;  // 2^(1/3)
    const double w1 = 1.0 / (2.0 - cbrt2);
    const double w2 = -cbrt2 / (2.0 - cbrt2);
    const double ha = w1 * h, hb = w2 * h;
    // the closing half kick of one Verlet kernel and the opening half kick of the next use the same acceleration:
    // v += (ha/2) a ; v += (hb/2) a  is issued as one fma with (ha + hb)/2 (differs from the reference's two
    // roundings by <= 1 ulp of the kick; 4N FP64 operations fewer per sub-step)
    const double hab = 0.5 * ha + 0.5 * hb;
    kick<N>(s, 0.5 * ha);
    if (!TANGENT && N >= NB_Y4_ROLL_MINN) {
      // the three stages as trips of one loop: one instance of the unrolled pair loop per kernel instead of three
      // (same operations in the same order: bit-identical)
#pragma unroll 1
      for (int st = 0; st < 3; ++st) {
        drift<N>(s, st == 1 ? hb : ha);
        pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);
        kick<N>(s, st == 2 ? 0.5 * ha : hab);
      }
    } else {
      drift<N>(s, ha);
      pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);
      kick<N>(s, hab);
      drift<N>(s, hb);
      pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);
      kick<N>(s, hab);
      drift<N>(s, ha);
      pair_pass<N, TANGENT, GUARD>(s, drx, dry, dax, day);
      kick<N>(s, 0.5 * ha);
    }
  } else {  // NB_MODE_WHFAST: Kepler(h/2) . full-force kick(h) . Kepler(h/2)   whfast_scheme.py:71-93
    // two trips of one loop, so that the Kepler solver -- by far the largest piece of code -- is instantiated once per
    // kernel (+5 % on the C4 cohort: instruction fetch).  Tried and dropped (r2, measured): solving the planets of a
    // system W = 2 / 3 at a time in lock-step so that their dependency chains interleave -- bit-identical, but 1.6x
    // (N = 3) to 5x (N = 4, 5) SLOWER: every lane then runs to the slower planet's iteration count, and a converged
    // lane divides 0 by f', which is the generic division's slow path.
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      kep = max(kep, kepler_drift<N, EXACT>(s, m, G, 0.5 * h, iters));
      if (half == 0) {
        if (EXACT) {
          // physically correct Wisdom-Holman: kick with the INTERACTION acceleration only (the star-planet Kepler
          // terms are already in the drift); the reference kicks with the full force (whfast_scheme.py:85-88)
          double ix[N], iy[N];
          wh_interaction_accel<N>(s, m, G, ix, iy);
#pragma unroll
          for (int i = 0; i < N; ++i) { s.ax[i] = ix[i]; s.ay[i] = iy[i]; }
        } else {
          pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);
        }
        kick<N>(s, h);
      }
    }
    if (TANGENT) pair_pass<N, true, GUARD>(s, drx, dry, dax, day);
  }
  return kep;
}

__device__ __forceinline__ double drift_of(double a0, double a1) {  // stability_analyzer.py:147-170
  if (is_finite(a0) && fabs(a0) > 0.0 && is_finite(a1)) return fabs((a1 - a0) / a0);
  if (is_finite(a0) && is_finite(a1)) return fabs(a1 - a0);
  return __longlong_as_double(0x7ff0000000000000LL);
}

template <int N>
__device__ __forceinline__ void load_state(const RunArgs& a, int sys, SysState<N>& s, double* m) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    m[i] = a.m[(size_t)sys * N + i];
    s.gm[i] = a.G * m[i];
    s.x[i] = a.q[((size_t)sys * N + i) * 2 + 0];
    s.y[i] = a.q[((size_t)sys * N + i) * 2 + 1];
    s.vx[i] = a.v[((size_t)sys * N + i) * 2 + 0];
    s.vy[i] = a.v[((size_t)sys * N + i) * 2 + 1];
  }
  const double eps = a.eps[sys];
  s.eps2 = eps * eps;
}

template <int N>
__device__ __forceinline__ int store_state(const RunArgs& a, int sys, const SysState<N>& s, bool write) {
  bool finite = true;
#pragma unroll
  for (int i = 0; i < N; ++i)
    finite = finite && is_finite(s.x[i]) && is_finite(s.y[i]) && is_finite(s.vx[i]) && is_finite(s.vy[i]);
  if (write) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      a.q[((size_t)sys * N + i) * 2 + 0] = s.x[i];
      a.q[((size_t)sys * N + i) * 2 + 1] = s.y[i];
      a.v[((size_t)sys * N + i) * 2 + 0] = s.vx[i];
      a.v[((size_t)sys * N + i) * 2 + 1] = s.vy[i];
    }
  }
  return finite ? 0 : NB_STATUS_NONFINITE;
}

template <int N, int MODE>
__host__ __device__ constexpr bool use_pairlane() { return N >= 5 && MODE != NB_MODE_WHFAST; }

// ---------------------------------------------------------------------------------------------
// phase 1: the main loop, n_steps macro steps with step_metrics sampling (stability_analyzer.py:113-128).
// Nothing in here takes the address of the state, so it stays in registers for the whole run.
// ---------------------------------------------------------------------------------------------
template <int N, int MODE, bool GUARD, bool EXACT>
__global__ void __launch_bounds__(128, (MODE == NB_MODE_WHFAST && N <= NB_WH_MINB_MAXN) ? NB_WH_MINB : 1) ensemble_main_kernel(RunArgs a, int write_state) {
  const int bid = (int)blockIdx.x + a.block0;                          // logical CTA (the launch may be split head / rest)
  stamp_begin(a.tstamp);
  if (MODE != NB_MODE_WHFAST && bid < a.group_blocks) {                // latency-optimised mappings for the n_sub-heavy head
    if constexpr (use_pairlane<N, MODE>()) {
      __shared__ __align__(16) double pl_smem[PairLane<N>::SMEM_DOUBLES];
      pairlane_main<N, MODE, GUARD>(a, write_state, pl_smem, bid);     // one pair per lane (N >= 5)
    } else {
      group_body<N, MODE == NB_MODE_WHFAST ? NB_MODE_VERLET : MODE, GUARD>(a, 0, write_state, bid);   // one body per lane
    }
    stamp_end(a.tstamp);
    return;
  }
  const int nh = (a.group_blocks > 0) ? min(*a.n_heavy, a.B) : 0;
  const int t = (bid - a.group_blocks) * blockDim.x + threadIdx.x;
  if (t >= a.B - nh) return;
  const int sys = a.perm ? a.perm[t + nh] : t;
  SysState<N> s;
  double m[N];
  load_state<N>(a, sys, s, m);
  const double G = a.G;
  const double eps = a.eps[sys];
  const int n_sub = a.n_sub ? max(1, a.n_sub[sys]) : 1;
  const double h = a.dt / (double)n_sub;
  int kep_worst = 0, kep_iters = 0;
  if (MODE != NB_MODE_WHFAST) pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);   // FSAL start

  double com_sum = 0.0, com_max = -1.0, var_sum = 0.0, var_max = -1.0, cos_sum = 0.0, cos_min = 2.0, th_sum = 0.0;
  double Lfirst = 0.0;
  bool have_first = false, cos_nan = false;
  int n_samp = 0, next_sample = 0;
  const int interval = a.sample_interval;
  // diagnostics.py:246-249 with pi = 0: loop-invariant
  const double theta_eps = (eps != 0.0) ? atan2(0.0, eps) : __longlong_as_double(0x7ff8000000000000LL);
  for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll 1
    for (int k = 0; k < n_sub; ++k)
      kep_worst = max(kep_worst, substep<N, MODE, GUARD, EXACT, false>(s, m, G, h, nullptr, nullptr, nullptr, nullptr, kep_iters));
    if (interval > 0 && step == next_sample) {       // diagnostics.py:241-285
      next_sample += interval;
      double cx = 0.0, cy = 0.0, Lt = 0.0, Li[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        cx += m[i] * s.x[i];
        cy += m[i] * s.y[i];
        Li[i] = m[i] * (s.x[i] * s.vy[i] - s.y[i] * s.vx[i]);
        Lt += Li[i];
      }
      const double com = sqrt(cx * cx + cy * cy);
      const double mean = Lt / N;
      double var = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) var += (Li[i] - mean) * (Li[i] - mean);
      var /= N;
      if (!have_first) { Lfirst = Lt; have_first = true; }
      double c;
      if (Lfirst != 0.0 && Lt != 0.0) c = (Lt * Lfirst) / (fabs(Lt) * fabs(Lfirst));
      else { c = 0.0; cos_nan = true; }
      com_sum += com; com_max = fmax(com_max, com);
      var_sum += var; var_max = fmax(var_max, var);
      cos_sum += c; cos_min = fmin(cos_min, c);
      th_sum += theta_eps;
      ++n_samp;
    }
  }
  int st = store_state<N>(a, sys, s, write_state != 0);
  if (MODE == NB_MODE_WHFAST && kep_worst > 64) st |= NB_STATUS_KEPLER_NOCONV;
  if (a.status) a.status[sys] = st;
  if (MODE == NB_MODE_WHFAST && a.work) {   // counted work for the roofline (mean Newton iterations)
    a.work[2 * (size_t)sys] = (double)kep_iters;
    a.work[2 * (size_t)sys + 1] = 2.0 * (N - 1) * (double)n_sub * (double)a.n_steps;
  }
  if (a.dyn) {
    double* f = a.dyn + (size_t)sys * NB_N_DYN;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double inv = n_samp > 0 ? 1.0 / (double)n_samp : nan;
    f[NB_F_COM_MEAN] = n_samp > 0 ? com_sum * inv : nan;
    f[NB_F_COM_MAX] = n_samp > 0 ? com_max : nan;
    f[NB_F_JEPS_MEAN] = n_samp > 0 ? 0.0 : nan;   // classic: pi = 0, mu_soft = 1 (diagnostics.py:246-249)
    f[NB_F_JEPS_STD] = n_samp > 0 ? 0.0 : nan;
    f[NB_F_THETA_MEAN] = n_samp > 0 ? th_sum * inv : nan;
    f[NB_F_THETA_STD] = n_samp > 0 ? ((eps != 0.0) ? 0.0 : nan) : nan;
    f[NB_F_COS_MEAN] = (n_samp > 0 && !cos_nan) ? cos_sum * inv : nan;
    f[NB_F_COS_MIN] = (n_samp > 0 && !cos_nan) ? cos_min : nan;
    f[NB_F_VARL_MEAN] = n_samp > 0 ? var_sum * inv : nan;
    f[NB_F_VARL_MAX] = n_samp > 0 ? var_max : nan;
    f[NB_F_TIDAL_MEAN] = n_samp > 0 ? 0.0 : nan;  // integrator.py:48 -- _last_tr_hessian is never updated
    f[NB_F_TIDAL_MAX] = n_samp > 0 ? 0.0 : nan;
  }
  stamp_end(a.tstamp);
}

// ---------------------------------------------------------------------------------------------
// phase 2: MEGNO (evolution_features.py:34-66): n_megno further steps with the tangent map
// ---------------------------------------------------------------------------------------------
template <int N, int MODE, bool GUARD, bool EXACT>
__global__ void __launch_bounds__(128) ensemble_megno_kernel(RunArgs a, int write_state) {
  if (MODE != NB_MODE_WHFAST && (int)blockIdx.x < a.group_blocks) {   // lane-per-body mapping for the n_sub-heavy head
    group_body<N, MODE == NB_MODE_WHFAST ? NB_MODE_VERLET : MODE, GUARD>(a, 1, write_state, (int)blockIdx.x);
    return;
  }
  const int nh = (a.group_blocks > 0) ? min(*a.n_heavy, a.B) : 0;
  const int t = ((int)blockIdx.x - a.group_blocks) * blockDim.x + threadIdx.x;
  if (t >= a.B - nh) return;
  const int sys = a.perm ? a.perm[t + nh] : t;
  SysState<N> s;
  double m[N];
  load_state<N>(a, sys, s, m);
  const double G = a.G;
  const int n_sub = a.n_sub ? max(1, a.n_sub[sys]) : 1;
  const double h = a.dt / (double)n_sub;
  int kep_worst = 0, kep_iters = 0;
  if (MODE != NB_MODE_WHFAST) pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);

  double drx[N], dry[N], dvx[N], dvy[N], dax[N], day[N];
  {
    double M = 0.0, cx = 0.0, cy = 0.0, ux = 0.0, uy = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      drx[i] = a.raw_dr[((size_t)sys * N + i) * 2 + 0];
      dry[i] = a.raw_dr[((size_t)sys * N + i) * 2 + 1];
      dvx[i] = a.raw_dv[((size_t)sys * N + i) * 2 + 0];
      dvy[i] = a.raw_dv[((size_t)sys * N + i) * 2 + 1];
      M += m[i];
      cx += m[i] * drx[i]; cy += m[i] * dry[i];
      ux += m[i] * dvx[i]; uy += m[i] * dvy[i];
    }
    cx /= M; cy /= M; ux /= M; uy /= M;
    double nr = 0.0, nv = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      drx[i] -= cx; dry[i] -= cy; dvx[i] -= ux; dvy[i] -= uy;
      nr += drx[i] * drx[i] + dry[i] * dry[i];
      nv += dvx[i] * dvx[i] + dvy[i] * dvy[i];
    }
    nr = sqrt(nr); nv = sqrt(nv);
#pragma unroll
    for (int i = 0; i < N; ++i) { drx[i] /= nr; dry[i] /= nr; dvx[i] /= nv; dvy[i] /= nv; }
  }
  double tt = 0.0, accum = 0.0;
  const double dt = a.dt;
  for (int step = 0; step < a.n_megno; ++step) {
#pragma unroll 1
    for (int k = 0; k < n_sub - 1; ++k)
      kep_worst = max(kep_worst, substep<N, MODE, GUARD, EXACT, false>(s, m, G, h, nullptr, nullptr, nullptr, nullptr, kep_iters));
    // delta_r += delta_v dt does not depend on the step, so it can precede the fused last evaluation
#pragma unroll
    for (int i = 0; i < N; ++i) { drx[i] = fma(dvx[i], dt, drx[i]); dry[i] = fma(dvy[i], dt, dry[i]); }
    kep_worst = max(kep_worst, substep<N, MODE, GUARD, EXACT, true>(s, m, G, h, drx, dry, dax, day, kep_iters));
    double nr = 0.0, nv = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      dvx[i] = fma(dax[i], dt, dvx[i]);
      dvy[i] = fma(day[i], dt, dvy[i]);
      nr += drx[i] * drx[i] + dry[i] * dry[i];
    }
    tt += dt;
    nr = sqrt(nr);
    if (nr < 1e-12) {
#pragma unroll
      for (int i = 0; i < N; ++i) { drx[i] /= nr; dry[i] /= nr; dvx[i] /= nr; dvy[i] /= nr; }
      nr = 1.0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) nv += dvx[i] * dvx[i] + dvy[i] * dvy[i];
    nv = sqrt(nv);
    accum += (nv / nr) * tt * dt;
  }
  const double megno = 2.0 * accum / tt;
  const double lyap = (megno == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : tt / fabs(megno);
  int st = store_state<N>(a, sys, s, write_state != 0);
  if (MODE == NB_MODE_WHFAST && kep_worst > 64) st |= NB_STATUS_KEPLER_NOCONV;
  if (a.status) a.status[sys] |= st;
  if (a.dyn) {
    double* f = a.dyn + (size_t)sys * NB_N_DYN;
    f[NB_F_MEGNO] = megno;
    f[NB_F_LYAP_TIME] = lyap;
    f[NB_F_T_END] = tt;
  }
}

template <int N, int MODE>
static int launch_run_mode(const RunArgs& a_in, int phase, int write_state, cudaStream_t st) {
  RunArgs a = a_in;
  const int threads = 128;
  a.group_blocks = (MODE != NB_MODE_WHFAST && a.n_heavy && a.perm)
                       ? ((phase == 0 && use_pairlane<N, MODE>()) ? pairlane_blocks_for<N>(a.B) : group_blocks_for<N>(a.B))
                       : 0;
  int blocks = a.group_blocks + (a.B + threads - 1) / threads;
  if (phase == 0 && a.block_count > 0) blocks = min(a.block_count, blocks - a.block0);   // head or rest of a split launch
  if (phase != 0) { a.block0 = 0; a.block_count = 0; }
  if (blocks <= 0) return NB_OK;
  const bool exact = MODE == NB_MODE_WHFAST && (a.flags & NB_RUN_KEPLER_EXACT) != 0;
  if (phase == 0) {
    if (exact) ensemble_main_kernel<N, MODE, true, MODE == NB_MODE_WHFAST><<<blocks, threads, 0, st>>>(a, write_state);
    else ensemble_main_kernel<N, MODE, true, false><<<blocks, threads, 0, st>>>(a, write_state);
  } else {
    if (exact) ensemble_megno_kernel<N, MODE, true, MODE == NB_MODE_WHFAST><<<blocks, threads, 0, st>>>(a, write_state);
    else ensemble_megno_kernel<N, MODE, true, false><<<blocks, threads, 0, st>>>(a, write_state);
  }
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

// logical CTAs of the phase-0 launch and how many of them form the latency-mapped prefix
template <int N, int MODE>
static void main_blocks_mode(const RunArgs& a, int* total, int* prefix) {
  const int g = (MODE != NB_MODE_WHFAST && a.n_heavy && a.perm)
                    ? (use_pairlane<N, MODE>() ? pairlane_blocks_for<N>(a.B) : group_blocks_for<N>(a.B)) : 0;
  *prefix = g;
  *total = g + (a.B + 127) / 128;
}
template <int MODE>
static void main_blocks_n(const RunArgs& a, int N, int* total, int* prefix) {
  switch (N) {
    case 2: main_blocks_mode<2, MODE>(a, total, prefix); break;
    case 3: main_blocks_mode<3, MODE>(a, total, prefix); break;
    case 4: main_blocks_mode<4, MODE>(a, total, prefix); break;
    case 5: main_blocks_mode<5, MODE>(a, total, prefix); break;
    case 6: main_blocks_mode<6, MODE>(a, total, prefix); break;
    case 7: main_blocks_mode<7, MODE>(a, total, prefix); break;
    default: main_blocks_mode<8, MODE>(a, total, prefix); break;
  }
}

template <int MODE>
static int launch_run_n(const RunArgs& a, int N, int phase, int write_state, cudaStream_t st) {
  switch (N) {
    case 2: return launch_run_mode<2, MODE>(a, phase, write_state, st);
    case 3: return launch_run_mode<3, MODE>(a, phase, write_state, st);
    case 4: return launch_run_mode<4, MODE>(a, phase, write_state, st);
    case 5: return launch_run_mode<5, MODE>(a, phase, write_state, st);
    case 6: return launch_run_mode<6, MODE>(a, phase, write_state, st);
    case 7: return launch_run_mode<7, MODE>(a, phase, write_state, st);
    case 8: return launch_run_mode<8, MODE>(a, phase, write_state, st);
    default: set_error("N must be in 2..8"); return NB_ERR_ARG;
  }
}

}  // namespace nb
