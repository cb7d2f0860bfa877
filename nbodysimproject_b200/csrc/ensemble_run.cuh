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
__device__ __forceinline__ int kepler_drift(SysState<N>& s, const double* m, double G, double tau) {
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
      const int it = EXACT ? kepler_exact(rx, ry, ux, uy, mu, tau) : kepler_reference(rx, ry, ux, uy, mu, tau);
      worst = max(worst, it);
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
                                       const double* dry, double* dax, double* day) {
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
    kick<N>(s, 0.5 * ha);
    drift<N>(s, ha);
    pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);
    kick<N>(s, 0.5 * ha);
    kick<N>(s, 0.5 * hb);
    drift<N>(s, hb);
    pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);
    kick<N>(s, 0.5 * hb);
    kick<N>(s, 0.5 * ha);
    drift<N>(s, ha);
    pair_pass<N, TANGENT, GUARD>(s, drx, dry, dax, day);
    kick<N>(s, 0.5 * ha);
  } else {  // NB_MODE_WHFAST: Kepler(h/2) . full-force kick(h) . Kepler(h/2)   whfast_scheme.py:71-93
    kep = kepler_drift<N, EXACT>(s, m, G, 0.5 * h);
    pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);
    kick<N>(s, h);
    kep = max(kep, kepler_drift<N, EXACT>(s, m, G, 0.5 * h));
    if (TANGENT) pair_pass<N, true, GUARD>(s, drx, dry, dax, day);
  }
  return kep;
}

// T + U with double-double accumulation; each part rounded to fp64 and then added, like
// diagnostics.py:543-549 does with its long-double Kahan sums.
template <int N>
__device__ __noinline__ double energy_dd(const double* m, const double* x, const double* y, const double* vx,
                                         const double* vy, double eps, double G) {
  dd T = dd_make(0.0);
  for (int i = 0; i < N; ++i) {
    dd v2 = dd_add(two_prod(vx[i], vx[i]), two_prod(vy[i], vy[i]));
    T = dd_add(T, dd_mul_d(dd_mul_d(v2, m[i]), 0.5));
  }
  dd S = dd_make(0.0);
  const dd e2 = two_prod(eps, eps);
  for (int i = 0; i < N; ++i)
    for (int j = i + 1; j < N; ++j) {
      dd dx = two_sum(x[i], -x[j]);
      dd dy = two_sum(y[i], -y[j]);
      dd r2 = dd_add(dd_add(dd_mul(dx, dx), dd_mul(dy, dy)), e2);
      if (!(r2.hi > 0.0)) r2 = dd_make(1e-300);
      dd inv = dd_div(dd_make(1.0), dd_sqrt(r2));
      S = dd_add(S, dd_mul(two_prod(m[i], m[j]), inv));
    }
  const double Tf = dd_to_double(T);
  const double Vf = dd_to_double(dd_mul_d(S, -G));
  return Tf + Vf;
}

__device__ __forceinline__ double drift_of(double a0, double a1) {  // stability_analyzer.py:147-170
  if (is_finite(a0) && fabs(a0) > 0.0 && is_finite(a1)) return fabs((a1 - a0) / a0);
  if (is_finite(a0) && is_finite(a1)) return fabs(a1 - a0);
  return __longlong_as_double(0x7ff0000000000000LL);
}

template <int N>
__device__ __forceinline__ double angmom(const double* m, const SysState<N>& s) {  // diagnostics.py:553-557
  double L = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) L += m[i] * (s.x[i] * s.vy[i] - s.y[i] * s.vx[i]);
  return L;
}

// ---------------------------------------------------------------------------------------------
// run kernel
// ---------------------------------------------------------------------------------------------
template <int N, int MODE, bool GUARD, bool EXACT>
__global__ void __launch_bounds__(128) ensemble_run_kernel(RunArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.B) return;
  const int sys = a.perm ? a.perm[t] : t;
  SysState<N> s;
  double m[N];
  const double G = a.G;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    m[i] = a.m[(size_t)sys * N + i];
    s.gm[i] = G * m[i];
    s.x[i] = a.q[((size_t)sys * N + i) * 2 + 0];
    s.y[i] = a.q[((size_t)sys * N + i) * 2 + 1];
    s.vx[i] = a.v[((size_t)sys * N + i) * 2 + 0];
    s.vy[i] = a.v[((size_t)sys * N + i) * 2 + 1];
  }
  const double eps = a.eps[sys];
  s.eps2 = eps * eps;
  const int n_sub = a.n_sub ? max(1, a.n_sub[sys]) : 1;
  const double h = a.dt / (double)n_sub;
  const bool want_energy = (a.flags & NB_RUN_ENERGY) != 0;
  int st = 0;
  int kep_worst = 0;

  double E0 = 0.0, L0 = 0.0;
  if (want_energy) {
    E0 = energy_dd<N>(m, s.x, s.y, s.vx, s.vy, eps, G);
    L0 = angmom<N>(m, s);
  }
  // FSAL start acceleration
  if (MODE != NB_MODE_WHFAST) pair_pass<N, false, GUARD>(s, nullptr, nullptr, nullptr, nullptr);

  // ---- main loop with step_metrics sampling (stability_analyzer.py:113-128)
  double com_sum = 0.0, com_max = -1.0, var_sum = 0.0, var_max = -1.0, cos_sum = 0.0, cos_min = 2.0;
  double th_sum = 0.0;
  double Lfirst = 0.0;
  bool have_first = false, cos_nan = false;
  int n_samp = 0;
  int next_sample = 0;
  const int interval = a.sample_interval;
  for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll 1
    for (int k = 0; k < n_sub; ++k)
      kep_worst = max(kep_worst, substep<N, MODE, GUARD, EXACT, false>(s, m, G, h, nullptr, nullptr, nullptr, nullptr));
    if (interval > 0 && step == next_sample) {
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
      th_sum += (eps != 0.0) ? atan2(0.0, eps) : __longlong_as_double(0x7ff8000000000000LL);
      ++n_samp;
    }
  }

  double E1 = 0.0, L1 = 0.0;
  if (want_energy) {
    E1 = energy_dd<N>(m, s.x, s.y, s.vx, s.vy, eps, G);
    L1 = angmom<N>(m, s);
  }

  // ---- MEGNO (evolution_features.py:34-66)
  double megno = 2.0, lyap = __longlong_as_double(0x7ff0000000000000LL), t_end = 0.0;
  if (a.n_megno > 0) {
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
        kep_worst = max(kep_worst, substep<N, MODE, GUARD, EXACT, false>(s, m, G, h, nullptr, nullptr, nullptr, nullptr));
      // delta_r += delta_v dt does not depend on the step, so it can precede the fused last evaluation
#pragma unroll
      for (int i = 0; i < N; ++i) { drx[i] = fma(dvx[i], dt, drx[i]); dry[i] = fma(dvy[i], dt, dry[i]); }
      kep_worst = max(kep_worst, substep<N, MODE, GUARD, EXACT, true>(s, m, G, h, drx, dry, dax, day));
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
    megno = 2.0 * accum / tt;
    lyap = (megno == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : tt / fabs(megno);
    t_end = tt;
  }

  // ---- outputs
  bool finite = true;
#pragma unroll
  for (int i = 0; i < N; ++i)
    finite = finite && is_finite(s.x[i]) && is_finite(s.y[i]) && is_finite(s.vx[i]) && is_finite(s.vy[i]);
  if (!finite) st |= NB_STATUS_NONFINITE;
  if (MODE == NB_MODE_WHFAST && kep_worst >= 64) st |= NB_STATUS_KEPLER_NOCONV;
  if (a.status) a.status[sys] = st;

  if (a.flags & NB_RUN_WRITE_STATE) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      a.q[((size_t)sys * N + i) * 2 + 0] = s.x[i];
      a.q[((size_t)sys * N + i) * 2 + 1] = s.y[i];
      a.v[((size_t)sys * N + i) * 2 + 0] = s.vx[i];
      a.v[((size_t)sys * N + i) * 2 + 1] = s.vy[i];
    }
  }
  if (a.dyn) {
    double* f = a.dyn + (size_t)sys * NB_N_DYN;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double ed = want_energy ? drift_of(E0, E1) : nan;
    const double ld = want_energy ? drift_of(L0, L1) : nan;
    const double inv = n_samp > 0 ? 1.0 / (double)n_samp : nan;
    const double com_mean = n_samp > 0 ? com_sum * inv : nan;
    f[NB_F_ENERGY_DRIFT] = ed;
    f[NB_F_ANGMOM_DRIFT] = ld;
    f[NB_F_COM_MEAN] = com_mean;
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
    f[NB_F_MEGNO] = megno;
    f[NB_F_LYAP_TIME] = lyap;
    f[NB_F_IS_STABLE] = ((ed < 0.01) && (ld < 0.01) && (com_mean < 1.0) && (megno < 10.0)) ? 1.0 : 0.0;
    f[NB_F_E0] = E0;
    f[NB_F_E1] = E1;
    f[NB_F_L0] = L0;
    f[NB_F_L1] = L1;
    f[NB_F_T_END] = t_end;
  }
}


template <int N, int MODE>
static int launch_run_mode(const RunArgs& a, cudaStream_t st) {
  const int threads = 128;
  const int blocks = (a.B + threads - 1) / threads;
  const bool exact = (a.flags & NB_RUN_KEPLER_EXACT) != 0;
  if (MODE == NB_MODE_WHFAST && exact)
    ensemble_run_kernel<N, MODE, true, true><<<blocks, threads, 0, st>>>(a);
  else
    ensemble_run_kernel<N, MODE, true, false><<<blocks, threads, 0, st>>>(a);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

template <int MODE>
static int launch_run_n(const RunArgs& a, int N, cudaStream_t st) {
  switch (N) {
    case 2: return launch_run_mode<2, MODE>(a, st);
    case 3: return launch_run_mode<3, MODE>(a, st);
    case 4: return launch_run_mode<4, MODE>(a, st);
    case 5: return launch_run_mode<5, MODE>(a, st);
    case 6: return launch_run_mode<6, MODE>(a, st);
    case 7: return launch_run_mode<7, MODE>(a, st);
    case 8: return launch_run_mode<8, MODE>(a, st);
    default: set_error("N must be in 2..8"); return NB_ERR_ARG;
  }
}

}  // namespace nb
