// ensemble_adaptive.cu -- classic ADAPTIVE softening for verlet / yoshida4 ensembles.
//
// Reference flow (integrator.py:78-136, 200-227; softening_manager.py:186-199, 246-257, 298-336, 423-471, 541-547):
// after EVERY sub-step the softening is re-derived from the current minimum separation,
//     eps_new = clamp(max(eps_min, r_min / softening_scale), <= 10 s0), limited to [eps/2, 2 eps],
// the potential-energy jump G sum_{i<j} m_i m_j (1/rho_new - 1/rho_old) (+ the barrier-energy difference) is booked
// into softening_energy_delta, and the following force evaluations use the new epsilon.  One thread per system,
// state in registers like the fixed-softening kernel; because epsilon changes between sub-steps the first
// acceleration of a sub-step is re-evaluated (the FSAL reuse is only valid inside a sub-step).
#include "pair_small.cuh"
#include "args.cuh"

namespace nb {

struct AdaptArgs {
  const double* m;
  double* q;
  double* v;
  double* eps;              // [B] current softening manager.s, in/out
  const double* soft_par;   // [B][3]: s0, min_softening, softening_scale
  double G;
  int B;
  double dt;
  int n_steps;
  const int32_t* n_sub;
  double k_wall;
  int n_exp;
  double* e_delta;          // [B] softening_energy_delta, in/out
  double* eps_hist;         // [B][n_steps] softening after each macro step (optional)
  int32_t* status;
  // analysis (run_stability_analysis on an adaptive copy): step_metrics sampling and the MEGNO phase
  int sample_interval;
  double* dyn;              // [B][NB_N_DYN] or null
  const double* eps_energy; // [B] sim._epsilon: the constant epsilon the reference's diagnostics use (diagnostics.py:474)
  int n_megno;
  const double* raw_dr;
  const double* raw_dv;
};

__device__ __forceinline__ double barrier_energy_dev(double eps, double a, double b, double k_wall, int n) {  // barrier.py:35-63
  if (!(is_finite(k_wall) && k_wall > 0.0 && n >= 2)) return 0.0;
  if (b < a) { const double t = a; a = b; b = t; }
  const int p = n - 1;
  const double l = fmax(0.0, a - eps), r = fmax(0.0, eps - b);
  double lp = 1.0, rp = 1.0;
  for (int i = 0; i < p; ++i) { lp *= l; rp *= r; }
  return (k_wall / (double)p) * (lp + rp);
}

// PHASE 0: n_steps macro steps (+ step_metrics sampling when a.dyn); PHASE 1: n_megno steps with the tangent map
// (evolution_features.py:34-66; the variational equations use the CURRENT manager.step_s2, tangent_map.py:21-59)
template <int N, int MODE, int PHASE>
__global__ void __launch_bounds__(128) ensemble_adaptive_kernel(AdaptArgs a) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= a.B) return;
  SysState<N> s;
  double m[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    m[i] = a.m[(size_t)sys * N + i];
    s.gm[i] = a.G * m[i];
    s.x[i] = a.q[((size_t)sys * N + i) * 2 + 0];
    s.y[i] = a.q[((size_t)sys * N + i) * 2 + 1];
    s.vx[i] = a.v[((size_t)sys * N + i) * 2 + 0];
    s.vy[i] = a.v[((size_t)sys * N + i) * 2 + 1];
  }
  double soft = a.eps[sys];
  const double s0 = a.soft_par[(size_t)sys * 3 + 0], eps_min = a.soft_par[(size_t)sys * 3 + 1];
  const double scale = a.soft_par[(size_t)sys * 3 + 2];
  const double eps_cap = 10.0 * s0;
  double e_delta = a.e_delta ? a.e_delta[sys] : 0.0;
  const int n_sub = a.n_sub ? max(1, a.n_sub[sys]) : 1;
  const double h = a.dt / (double)n_sub;
  const double cbrt2 = 1.2599210498948731648;
  const double ha = (1.0 / (2.0 - cbrt2)) * h, hb = (-cbrt2 / (2.0 - cbrt2)) * h;

  auto vkernel = [&](double hh) {     // integration_scheme_base.py:129-149 with the start acceleration already in s.ax/ay
    const double h2 = 0.5 * hh;
#pragma unroll
    for (int i = 0; i < N; ++i) { s.vx[i] = fma(h2, s.ax[i], s.vx[i]); s.vy[i] = fma(h2, s.ay[i], s.vy[i]); }
#pragma unroll
    for (int i = 0; i < N; ++i) { s.x[i] = fma(hh, s.vx[i], s.x[i]); s.y[i] = fma(hh, s.vy[i], s.y[i]); }
    pair_pass<N, false, true>(s, nullptr, nullptr, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < N; ++i) { s.vx[i] = fma(h2, s.ax[i], s.vx[i]); s.vy[i] = fma(h2, s.ay[i], s.vy[i]); }
  };

  // one macro step: n_sub sub-steps, each followed by the softening refresh
  auto macro_step = [&]() {
    double pending = 0.0;                                   // begin_step
#pragma unroll 1
    for (int k = 0; k < n_sub; ++k) {
      const double ef = sqrt(soft * soft);                  // simulation.py:539-581: eps = sqrt(manager.step_s2)
      s.eps2 = ef * ef;
      pair_pass<N, false, true>(s, nullptr, nullptr, nullptr, nullptr);
      if (MODE == NB_MODE_YOSHIDA4) { vkernel(ha); vkernel(hb); vkernel(ha); }
      else vkernel(h);
      // ---- refresh_softening(softening_from_min_sep(min separation))
      double r2min = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j) {
          const double dx = s.x[i] - s.x[j], dy = s.y[i] - s.y[j];
          r2min = fmin(r2min, dx * dx + dy * dy);
        }
      const double min_sep = fmax(sqrt(r2min), 1e-12);      // simulation.py:659-665
      double eps_new = soft;
      if (is_finite(min_sep) && min_sep > 0.0) {
        double prop = fmax(eps_min, min_sep / scale);
        prop = fmin(prop, eps_cap);
        eps_new = fmax(soft / 2.0, fmin(soft * 2.0, prop));
      }
      if (eps_new != soft && is_finite(soft) && is_finite(eps_new)) {     // softening_manager.py:423-471
        double dE = 0.0;
        const double e2o = soft * soft, e2n = eps_new * eps_new;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = i + 1; j < N; ++j) {
            const double dx = s.x[i] - s.x[j], dy = s.y[i] - s.y[j];
            const double r2 = dx * dx + dy * dy;
            const double wn = rsqrt_f64<true>(r2 + e2n), wo = rsqrt_f64<true>(r2 + e2o);
            dE = fma(m[i] * m[j], wn - wo, dE);
          }
        dE *= a.G;
        dE += barrier_energy_dev(eps_new, eps_min, eps_cap, a.k_wall, a.n_exp) -
              barrier_energy_dev(soft, eps_min, eps_cap, a.k_wall, a.n_exp);
        if (is_finite(dE)) pending += dE;
      }
      soft = eps_new;
      if (pending != 0.0) { e_delta += pending; pending = 0.0; }          // commit_substep
    }
  };

  if (PHASE == 0) {
    double com_sum = 0.0, com_max = -1.0, var_sum = 0.0, var_max = -1.0, cos_sum = 0.0, cos_min = 2.0, th_sum = 0.0;
    double Lfirst = 0.0;
    bool have_first = false, cos_nan = false;
    int n_samp = 0, next_sample = 0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double eps_d = a.eps_energy ? a.eps_energy[sys] : soft;
    const double theta_eps = (eps_d != 0.0) ? atan2(0.0, eps_d) : nan;   // diagnostics.py:246-249 with pi = 0
    const int interval = a.dyn ? a.sample_interval : 0;
    for (int step = 0; step < a.n_steps; ++step) {
      macro_step();
      if (a.eps_hist) a.eps_hist[(size_t)sys * a.n_steps + step] = soft;
      if (interval > 0 && step == next_sample) {           // diagnostics.py:241-285
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
    if (a.dyn) {
      double* f = a.dyn + (size_t)sys * NB_N_DYN;
      const double inv = n_samp > 0 ? 1.0 / (double)n_samp : nan;
      f[NB_F_COM_MEAN] = n_samp > 0 ? com_sum * inv : nan;
      f[NB_F_COM_MAX] = n_samp > 0 ? com_max : nan;
      f[NB_F_JEPS_MEAN] = n_samp > 0 ? 0.0 : nan;
      f[NB_F_JEPS_STD] = n_samp > 0 ? 0.0 : nan;
      f[NB_F_THETA_MEAN] = n_samp > 0 ? th_sum * inv : nan;
      f[NB_F_THETA_STD] = n_samp > 0 ? ((eps_d != 0.0) ? 0.0 : nan) : nan;
      f[NB_F_COS_MEAN] = (n_samp > 0 && !cos_nan) ? cos_sum * inv : nan;
      f[NB_F_COS_MIN] = (n_samp > 0 && !cos_nan) ? cos_min : nan;
      f[NB_F_VARL_MEAN] = n_samp > 0 ? var_sum * inv : nan;
      f[NB_F_VARL_MAX] = n_samp > 0 ? var_max : nan;
      f[NB_F_TIDAL_MEAN] = n_samp > 0 ? 0.0 : nan;
      f[NB_F_TIDAL_MAX] = n_samp > 0 ? 0.0 : nan;
    }
  } else {
    double drx[N], dry[N], dvx[N], dvy[N], dax[N], day[N];
    {
      double M = 0.0, cx = 0.0, cy = 0.0, ux = 0.0, uy = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        drx[i] = a.raw_dr[((size_t)sys * N + i) * 2 + 0]; dry[i] = a.raw_dr[((size_t)sys * N + i) * 2 + 1];
        dvx[i] = a.raw_dv[((size_t)sys * N + i) * 2 + 0]; dvy[i] = a.raw_dv[((size_t)sys * N + i) * 2 + 1];
        M += m[i];
        cx += m[i] * drx[i]; cy += m[i] * dry[i]; ux += m[i] * dvx[i]; uy += m[i] * dvy[i];
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
      macro_step();
#pragma unroll
      for (int i = 0; i < N; ++i) { drx[i] = fma(dvx[i], dt, drx[i]); dry[i] = fma(dvy[i], dt, dry[i]); }
      s.eps2 = soft * soft;                                 // manager.step_s2 AFTER the step's last refresh
      pair_pass<N, true, true>(s, drx, dry, dax, day);
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
      accum += (sqrt(nv) / nr) * tt * dt;
    }
    if (a.dyn && a.n_megno > 0) {
      double* f = a.dyn + (size_t)sys * NB_N_DYN;
      const double megno = 2.0 * accum / tt;
      f[NB_F_MEGNO] = megno;
      f[NB_F_LYAP_TIME] = (megno == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : tt / fabs(megno);
      f[NB_F_T_END] = tt;
    }
  }
  bool finite = true;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    finite = finite && is_finite(s.x[i]) && is_finite(s.y[i]) && is_finite(s.vx[i]) && is_finite(s.vy[i]);
    a.q[((size_t)sys * N + i) * 2 + 0] = s.x[i];
    a.q[((size_t)sys * N + i) * 2 + 1] = s.y[i];
    a.v[((size_t)sys * N + i) * 2 + 0] = s.vx[i];
    a.v[((size_t)sys * N + i) * 2 + 1] = s.vy[i];
  }
  a.eps[sys] = soft;
  if (a.e_delta) a.e_delta[sys] = e_delta;
  if (a.status) {
    const int st = finite ? 0 : NB_STATUS_NONFINITE;
    if (PHASE == 0) a.status[sys] = st; else a.status[sys] |= st;
  }
}

// ---------------------------------------------------------------------------------------------
// 9 .. 64 bodies: the same flow with a run-time body count, one thread per system, state in local memory.  The
// reference accepts any body count (simulation.py:39-162); this serves single simulations and small batches (facade,
// analysis of a handful of mid-sized systems), not a throughput benchmark.  The pair loops run in the order of
// pair_pass<N> (i < j ascending), so a body's sums are formed exactly as in the register-resident kernel.
// ---------------------------------------------------------------------------------------------
constexpr int AD_MAX = NB_MAX_N_MID;

struct AdState {
  double x[AD_MAX], y[AD_MAX], vx[AD_MAX], vy[AD_MAX], ax[AD_MAX], ay[AD_MAX], gm[AD_MAX], m[AD_MAX];
};

template <bool TANGENT>
__device__ __noinline__ void ad_pair_pass(AdState& s, int n, double eps2, const double* drx, const double* dry,
                                          double* dax, double* day) {
  for (int i = 0; i < n; ++i) {
    s.ax[i] = 0.0; s.ay[i] = 0.0;
    if (TANGENT) { dax[i] = 0.0; day[i] = 0.0; }
  }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const double dx = s.x[i] - s.x[j], dy = s.y[i] - s.y[j];
      const double r2 = fma(dx, dx, fma(dy, dy, eps2));
      double w2, w3;
      if (TANGENT) {
        const double w = rsqrt_f64<true>(r2);
        w2 = w * w;
        w3 = w2 * w;
      } else {
        w2 = 0.0;
        w3 = rsqrt3_f64<true>(r2);
      }
      const double cj = s.gm[j] * w3, ci = s.gm[i] * w3;
      s.ax[i] = fma(-cj, dx, s.ax[i]);
      s.ay[i] = fma(-cj, dy, s.ay[i]);
      s.ax[j] = fma(ci, dx, s.ax[j]);
      s.ay[j] = fma(ci, dy, s.ay[j]);
      if (TANGENT) {
        const double ex = drx[j] - drx[i], ey = dry[j] - dry[i];
        const double dot = -fma(dx, ex, dy * ey);
        const double c5 = 3.0 * dot * w2 * w3;
        const double tx = fma(ex, w3, c5 * dx), ty = fma(ey, w3, c5 * dy);
        dax[i] = fma(s.gm[j], tx, dax[i]);
        day[i] = fma(s.gm[j], ty, day[i]);
        dax[j] = fma(-s.gm[i], tx, dax[j]);
        day[j] = fma(-s.gm[i], ty, day[j]);
      }
    }
}

template <int MODE, int PHASE>
__global__ void __launch_bounds__(32) ensemble_adaptive_rt_kernel(AdaptArgs a, int n) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= a.B) return;
  AdState s;
  for (int i = 0; i < n; ++i) {
    s.m[i] = a.m[(size_t)sys * n + i];
    s.gm[i] = a.G * s.m[i];
    s.x[i] = a.q[((size_t)sys * n + i) * 2 + 0];
    s.y[i] = a.q[((size_t)sys * n + i) * 2 + 1];
    s.vx[i] = a.v[((size_t)sys * n + i) * 2 + 0];
    s.vy[i] = a.v[((size_t)sys * n + i) * 2 + 1];
  }
  double soft = a.eps[sys];
  const double s0 = a.soft_par[(size_t)sys * 3 + 0], eps_min = a.soft_par[(size_t)sys * 3 + 1];
  const double scale = a.soft_par[(size_t)sys * 3 + 2];
  const double eps_cap = 10.0 * s0;
  double e_delta = a.e_delta ? a.e_delta[sys] : 0.0;
  const int n_sub = a.n_sub ? max(1, a.n_sub[sys]) : 1;
  const double h = a.dt / (double)n_sub;
  const double cbrt2 = 1.2599210498948731648;
  const double ha = (1.0 / (2.0 - cbrt2)) * h, hb = (-cbrt2 / (2.0 - cbrt2)) * h;
  double eps2 = 0.0;

  auto vkernel = [&](double hh) {     // integration_scheme_base.py:129-149 with the start acceleration already in s.ax/ay
    const double h2 = 0.5 * hh;
    for (int i = 0; i < n; ++i) { s.vx[i] = fma(h2, s.ax[i], s.vx[i]); s.vy[i] = fma(h2, s.ay[i], s.vy[i]); }
    for (int i = 0; i < n; ++i) { s.x[i] = fma(hh, s.vx[i], s.x[i]); s.y[i] = fma(hh, s.vy[i], s.y[i]); }
    ad_pair_pass<false>(s, n, eps2, nullptr, nullptr, nullptr, nullptr);
    for (int i = 0; i < n; ++i) { s.vx[i] = fma(h2, s.ax[i], s.vx[i]); s.vy[i] = fma(h2, s.ay[i], s.vy[i]); }
  };
  auto macro_step = [&]() {           // n_sub sub-steps, each followed by the softening refresh (see the kernel above)
    double pending = 0.0;
    for (int k = 0; k < n_sub; ++k) {
      const double ef = sqrt(soft * soft);
      eps2 = ef * ef;
      ad_pair_pass<false>(s, n, eps2, nullptr, nullptr, nullptr, nullptr);
      if (MODE == NB_MODE_YOSHIDA4) { vkernel(ha); vkernel(hb); vkernel(ha); }
      else vkernel(h);
      double r2min = __longlong_as_double(0x7ff0000000000000LL);
      for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
          const double dx = s.x[i] - s.x[j], dy = s.y[i] - s.y[j];
          r2min = fmin(r2min, dx * dx + dy * dy);
        }
      const double min_sep = fmax(sqrt(r2min), 1e-12);
      double eps_new = soft;
      if (is_finite(min_sep) && min_sep > 0.0) {
        double prop = fmax(eps_min, min_sep / scale);
        prop = fmin(prop, eps_cap);
        eps_new = fmax(soft / 2.0, fmin(soft * 2.0, prop));
      }
      if (eps_new != soft && is_finite(soft) && is_finite(eps_new)) {
        double dE = 0.0;
        const double e2o = soft * soft, e2n = eps_new * eps_new;
        for (int i = 0; i < n; ++i)
          for (int j = i + 1; j < n; ++j) {
            const double dx = s.x[i] - s.x[j], dy = s.y[i] - s.y[j];
            const double r2 = dx * dx + dy * dy;
            const double wn = rsqrt_f64<true>(r2 + e2n), wo = rsqrt_f64<true>(r2 + e2o);
            dE = fma(s.m[i] * s.m[j], wn - wo, dE);
          }
        dE *= a.G;
        dE += barrier_energy_dev(eps_new, eps_min, eps_cap, a.k_wall, a.n_exp) -
              barrier_energy_dev(soft, eps_min, eps_cap, a.k_wall, a.n_exp);
        if (is_finite(dE)) pending += dE;
      }
      soft = eps_new;
      if (pending != 0.0) { e_delta += pending; pending = 0.0; }
    }
  };

  if (PHASE == 0) {
    double com_sum = 0.0, com_max = -1.0, var_sum = 0.0, var_max = -1.0, cos_sum = 0.0, cos_min = 2.0, th_sum = 0.0;
    double Lfirst = 0.0;
    bool have_first = false, cos_nan = false;
    int n_samp = 0, next_sample = 0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double eps_d = a.eps_energy ? a.eps_energy[sys] : soft;
    const double theta_eps = (eps_d != 0.0) ? atan2(0.0, eps_d) : nan;
    const int interval = a.dyn ? a.sample_interval : 0;
    for (int step = 0; step < a.n_steps; ++step) {
      macro_step();
      if (a.eps_hist) a.eps_hist[(size_t)sys * a.n_steps + step] = soft;
      if (interval > 0 && step == next_sample) {           // diagnostics.py:241-285
        next_sample += interval;
        double cx = 0.0, cy = 0.0, Lt = 0.0;
        for (int i = 0; i < n; ++i) {
          cx += s.m[i] * s.x[i];
          cy += s.m[i] * s.y[i];
          Lt += s.m[i] * (s.x[i] * s.vy[i] - s.y[i] * s.vx[i]);
        }
        const double com = sqrt(cx * cx + cy * cy);
        const double mean = Lt / n;
        double var = 0.0;
        for (int i = 0; i < n; ++i) {
          const double Li = s.m[i] * (s.x[i] * s.vy[i] - s.y[i] * s.vx[i]);
          var += (Li - mean) * (Li - mean);
        }
        var /= n;
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
    if (a.dyn) {
      double* f = a.dyn + (size_t)sys * NB_N_DYN;
      const double inv = n_samp > 0 ? 1.0 / (double)n_samp : nan;
      f[NB_F_COM_MEAN] = n_samp > 0 ? com_sum * inv : nan;
      f[NB_F_COM_MAX] = n_samp > 0 ? com_max : nan;
      f[NB_F_JEPS_MEAN] = n_samp > 0 ? 0.0 : nan;
      f[NB_F_JEPS_STD] = n_samp > 0 ? 0.0 : nan;
      f[NB_F_THETA_MEAN] = n_samp > 0 ? th_sum * inv : nan;
      f[NB_F_THETA_STD] = n_samp > 0 ? ((eps_d != 0.0) ? 0.0 : nan) : nan;
      f[NB_F_COS_MEAN] = (n_samp > 0 && !cos_nan) ? cos_sum * inv : nan;
      f[NB_F_COS_MIN] = (n_samp > 0 && !cos_nan) ? cos_min : nan;
      f[NB_F_VARL_MEAN] = n_samp > 0 ? var_sum * inv : nan;
      f[NB_F_VARL_MAX] = n_samp > 0 ? var_max : nan;
      f[NB_F_TIDAL_MEAN] = n_samp > 0 ? 0.0 : nan;
      f[NB_F_TIDAL_MAX] = n_samp > 0 ? 0.0 : nan;
    }
  } else {
    double drx[AD_MAX], dry[AD_MAX], dvx[AD_MAX], dvy[AD_MAX], dax[AD_MAX], day[AD_MAX];
    {
      double M = 0.0, cx = 0.0, cy = 0.0, ux = 0.0, uy = 0.0;
      for (int i = 0; i < n; ++i) {
        drx[i] = a.raw_dr[((size_t)sys * n + i) * 2 + 0]; dry[i] = a.raw_dr[((size_t)sys * n + i) * 2 + 1];
        dvx[i] = a.raw_dv[((size_t)sys * n + i) * 2 + 0]; dvy[i] = a.raw_dv[((size_t)sys * n + i) * 2 + 1];
        M += s.m[i];
        cx += s.m[i] * drx[i]; cy += s.m[i] * dry[i]; ux += s.m[i] * dvx[i]; uy += s.m[i] * dvy[i];
      }
      cx /= M; cy /= M; ux /= M; uy /= M;
      double nr = 0.0, nv = 0.0;
      for (int i = 0; i < n; ++i) {
        drx[i] -= cx; dry[i] -= cy; dvx[i] -= ux; dvy[i] -= uy;
        nr += drx[i] * drx[i] + dry[i] * dry[i];
        nv += dvx[i] * dvx[i] + dvy[i] * dvy[i];
      }
      nr = sqrt(nr); nv = sqrt(nv);
      for (int i = 0; i < n; ++i) { drx[i] /= nr; dry[i] /= nr; dvx[i] /= nv; dvy[i] /= nv; }
    }
    double tt = 0.0, accum = 0.0;
    const double dt = a.dt;
    for (int step = 0; step < a.n_megno; ++step) {
      macro_step();
      for (int i = 0; i < n; ++i) { drx[i] = fma(dvx[i], dt, drx[i]); dry[i] = fma(dvy[i], dt, dry[i]); }
      ad_pair_pass<true>(s, n, soft * soft, drx, dry, dax, day);     // manager.step_s2 AFTER the step's last refresh
      double nr = 0.0, nv = 0.0;
      for (int i = 0; i < n; ++i) {
        dvx[i] = fma(dax[i], dt, dvx[i]);
        dvy[i] = fma(day[i], dt, dvy[i]);
        nr += drx[i] * drx[i] + dry[i] * dry[i];
      }
      tt += dt;
      nr = sqrt(nr);
      if (nr < 1e-12) {
        for (int i = 0; i < n; ++i) { drx[i] /= nr; dry[i] /= nr; dvx[i] /= nr; dvy[i] /= nr; }
        nr = 1.0;
      }
      for (int i = 0; i < n; ++i) nv += dvx[i] * dvx[i] + dvy[i] * dvy[i];
      accum += (sqrt(nv) / nr) * tt * dt;
    }
    if (a.dyn && a.n_megno > 0) {
      double* f = a.dyn + (size_t)sys * NB_N_DYN;
      const double megno = 2.0 * accum / tt;
      f[NB_F_MEGNO] = megno;
      f[NB_F_LYAP_TIME] = (megno == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : tt / fabs(megno);
      f[NB_F_T_END] = tt;
    }
  }
  bool finite = true;
  for (int i = 0; i < n; ++i) {
    finite = finite && is_finite(s.x[i]) && is_finite(s.y[i]) && is_finite(s.vx[i]) && is_finite(s.vy[i]);
    a.q[((size_t)sys * n + i) * 2 + 0] = s.x[i];
    a.q[((size_t)sys * n + i) * 2 + 1] = s.y[i];
    a.v[((size_t)sys * n + i) * 2 + 0] = s.vx[i];
    a.v[((size_t)sys * n + i) * 2 + 1] = s.vy[i];
  }
  a.eps[sys] = soft;
  if (a.e_delta) a.e_delta[sys] = e_delta;
  if (a.status) {
    const int st = finite ? 0 : NB_STATUS_NONFINITE;
    if (PHASE == 0) a.status[sys] = st; else a.status[sys] |= st;
  }
}

template <int MODE, int PHASE>
static int launch_adaptive(const AdaptArgs& a, int N, cudaStream_t st) {
  const int threads = 128, blocks = (a.B + threads - 1) / threads;
  switch (N) {
    case 2: ensemble_adaptive_kernel<2, MODE, PHASE><<<blocks, threads, 0, st>>>(a); break;
    case 3: ensemble_adaptive_kernel<3, MODE, PHASE><<<blocks, threads, 0, st>>>(a); break;
    case 4: ensemble_adaptive_kernel<4, MODE, PHASE><<<blocks, threads, 0, st>>>(a); break;
    case 5: ensemble_adaptive_kernel<5, MODE, PHASE><<<blocks, threads, 0, st>>>(a); break;
    case 6: ensemble_adaptive_kernel<6, MODE, PHASE><<<blocks, threads, 0, st>>>(a); break;
    case 7: ensemble_adaptive_kernel<7, MODE, PHASE><<<blocks, threads, 0, st>>>(a); break;
    case 8: ensemble_adaptive_kernel<8, MODE, PHASE><<<blocks, threads, 0, st>>>(a); break;
    default:
      if (N > NB_MAX_N && N <= NB_MAX_N_MID) {
        ensemble_adaptive_rt_kernel<MODE, PHASE><<<(a.B + 31) / 32, 32, 0, st>>>(a, N);
        break;
      }
      set_error("N must be in 2..64");
      return NB_ERR_ARG;
  }
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int launch_energy(const double* m, const double* q, const double* v, const double* eps, double G, int B, int N, double* dyn,
                  int slot, cudaStream_t st);
int launch_finalize(double* dyn, int B, int have_energy, int have_megno, cudaStream_t st);

static int check_adaptive_mode(int mode) {
  if (mode != NB_MODE_VERLET && mode != NB_MODE_YOSHIDA4) {
    set_error("adaptive softening exists for verlet and yoshida4 (the reference turns whfast into verlet, "
              "simulation.py:103-107; ham_soft has its own epsilon flow)");
    return NB_ERR_UNSUPPORTED;
  }
  return NB_OK;
}

int ensemble_run_adaptive(const double* m, double* q, double* v, double* eps, const double* soft_par, double G, int B,
                          int N, int mode, double dt, int n_steps, const int32_t* n_sub, double k_wall, int n_exp,
                          double* e_delta, double* eps_hist, int32_t* status, cudaStream_t st) {
  if (!m || !q || !v || !eps || !soft_par || B < 0 || n_steps < 0) {
    set_error("nb_ensemble_run_adaptive_f64: bad arguments");
    return NB_ERR_ARG;
  }
  int rc = check_adaptive_mode(mode);
  if (rc != NB_OK) return rc;
  if (B == 0) return NB_OK;
  AdaptArgs a{m, q, v, eps, soft_par, G, B, dt, n_steps, n_sub, k_wall, n_exp, e_delta, eps_hist, status,
              0, nullptr, nullptr, 0, nullptr, nullptr};
  return mode == NB_MODE_YOSHIDA4 ? launch_adaptive<NB_MODE_YOSHIDA4, 0>(a, N, st) : launch_adaptive<NB_MODE_VERLET, 0>(a, N, st);
}

// run_stability_analysis (stability_analyzer.py:69-259) for adaptive-softening copies: E0 -> main loop with sampling
// -> E1 -> MEGNO -> drifts / is_stable, the energies with the constant eps_energy (diagnostics.py:474)
int ensemble_analyze_adaptive(const double* m, double* q, double* v, double* eps, const double* eps_energy,
                              const double* soft_par, double G, int B, int N, int mode, double dt, int n_steps,
                              int sample_interval, int n_megno, const int32_t* n_sub, const double* raw_dr,
                              const double* raw_dv, double k_wall, int n_exp, double* e_delta, double* dyn, int32_t* status,
                              cudaStream_t st) {
  if (!m || !q || !v || !eps || !eps_energy || !soft_par || !dyn || B < 0 || n_steps < 0 || n_megno < 0 ||
      (n_megno > 0 && (!raw_dr || !raw_dv))) {
    set_error("nb_ensemble_analyze_adaptive_f64: bad arguments");
    return NB_ERR_ARG;
  }
  int rc = check_adaptive_mode(mode);
  if (rc != NB_OK) return rc;
  if (B == 0) return NB_OK;
  AdaptArgs a{m, q, v, eps, soft_par, G, B, dt, n_steps, n_sub, k_wall, n_exp, e_delta, nullptr, status,
              sample_interval, dyn, eps_energy, n_megno, raw_dr, raw_dv};
  rc = launch_energy(m, q, v, eps_energy, G, B, N, dyn, 0, st);
  if (rc != NB_OK) return rc;
  rc = mode == NB_MODE_YOSHIDA4 ? launch_adaptive<NB_MODE_YOSHIDA4, 0>(a, N, st) : launch_adaptive<NB_MODE_VERLET, 0>(a, N, st);
  if (rc != NB_OK) return rc;
  rc = launch_energy(m, q, v, eps_energy, G, B, N, dyn, 1, st);
  if (rc != NB_OK) return rc;
  if (n_megno > 0) {
    rc = mode == NB_MODE_YOSHIDA4 ? launch_adaptive<NB_MODE_YOSHIDA4, 1>(a, N, st) : launch_adaptive<NB_MODE_VERLET, 1>(a, N, st);
    if (rc != NB_OK) return rc;
  }
  return launch_finalize(dyn, B, 1, n_megno > 0 ? 1 : 0, st);
}

}  // namespace nb
