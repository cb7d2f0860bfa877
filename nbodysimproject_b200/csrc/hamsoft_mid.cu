// hamsoft_mid.cu -- ham_soft (Strang split with a dynamical softening length) for systems of 9 .. 64 bodies:
// ONE CTA PER SYSTEM.  ham_soft is the reference's default integrator mode (sim_config.py:38) and the reference accepts
// any body count (simulation.py:39-162), so NBodySimulation(...) with a dozen bodies must step; the lane-per-evaluation
// kernels of hamsoft.cu stop at N = 8 (4 N + 1 <= 33 finite-difference evaluations = one warp).
//
// Mapping.  The system lives once in shared memory (HmSh).  The eps* model -- 4 N + 1 <= 257 independent fixed-point
// solves per S half-flow (hamsoft_eps_model.py:94-234, 316-400) -- is spread over the CTA, ONE EVALUATION PER THREAD
// (strided when 4 N + 1 > 128): a thread runs the whole Jacobi solve of its perturbed configuration serially, pair
// distances formed on the fly from the shared positions, its N smoothing lengths in local memory.  Every other flow
// (central differences, analytic fallback, J cap, V half kick, T drift, tangent map) is body-per-thread with
// deterministic CTA reductions (every thread adds the partials in ascending order, so eps and pi are replicated
// scalars that stay identical across the CTA).  The arithmetic per body / per evaluation is the arithmetic of hamsoft.cu;
// only the summation order of CTA-wide sums differs (rounding level).
// Not a throughput path: it serves single simulations and small batches of mid-sized systems.
//
// Restates the same reference lines as hamsoft.cu: hamsoft_stepper.py:47-308, 543-663, hamsoft_flows.py:427-762,
// 1102-1132, hamsoft_eps_model.py:94-289, 316-400, 451-556, 645-729, softening.py:86-131, barrier.py:35-113,
// hamiltonian_softening_integrator.py:145-296, 986-1221, diagnostics.py:241-285, 457-549, evolution_features.py:34-66.
#include "hamsoft_common.cuh"

namespace nb {

constexpr int HM_MAX = NB_MAX_N_MID;        // 64
constexpr int HM_THREADS = 128;
constexpr int HM_NP = HM_MAX * (HM_MAX - 1) / 2;

struct HmSh {
  double m[HM_MAX], x[HM_MAX], y[HM_MAX], vx[HM_MAX], vy[HM_MAX];
  double gx[HM_MAX], gy[HM_MAX];            // grad eps* of the current S half-flow
  double h[HM_MAX];                         // smoothing lengths of the unperturbed configuration
  double bx[HM_MAX], by[HM_MAX];            // per-body exchange buffers of the analytic fallback
  double drx[HM_MAX], dry[HM_MAX], dvx[HM_MAX], dvy[HM_MAX];   // tangent vectors (MEGNO)
  double f[4 * HM_MAX + 1];                 // eps* of the finite-difference configurations
  double rs[HM_NP];                         // pair separations (median test, legacy gradient)
  double red[HM_THREADS];
  double acc[HS_NACC];
  double inv_alpha;
  HsSpring spr;
  HsPar P;
};

// deterministic CTA reductions: every thread ends up with the same value (ascending order over the threads)
__device__ __forceinline__ double hm_sum(HmSh& sh, double v) {
  __syncthreads();
  sh.red[threadIdx.x] = v;
  __syncthreads();
  double s = 0.0;
  for (int k = 0; k < HM_THREADS; ++k) s += sh.red[k];
  return s;
}
__device__ __forceinline__ double hm_max(HmSh& sh, double v) {
  __syncthreads();
  sh.red[threadIdx.x] = v;
  __syncthreads();
  double s = sh.red[0];
  for (int k = 1; k < HM_THREADS; ++k) s = fmax(s, sh.red[k]);
  return s;
}
__device__ __forceinline__ int hm_pair_index(int n, int a, int b) { return a * n - a * (a + 1) / 2 + (b - a - 1); }

// Jacobi sweeps of hamsoft_eps_model.py:316-400 for the configuration (X, Y) with coordinate (pb, axis) replaced by pv
// (pb < 0: unperturbed), serially on the calling thread; h[n] receives the smoothing lengths, returns the sweeps executed.
// exp(-r^2/h^2) underflows to exactly 0 below -745.13: those pairs skip the exponential (same bits).
__device__ __noinline__ int hm_solve(const double* X, const double* Y, const double* M, int n, int pb, bool on_y,
                                     double pv, double eps_cur, const HsPar& P, double* h) {
  double lo = P.eps_min, hi = P.eps_max;
  if (hi < lo) { const double t = lo; lo = hi; hi = t; }
  const double flo = fmax(lo, 1.0e-12), cap = fmax(flo, hi);
  double h0 = eps_cur;
  if (!is_finite(h0) || h0 <= 0.0) h0 = 1.0;
  h0 = fmin(fmax(h0, flo), cap);
  for (int i = 0; i < n; ++i) h[i] = h0;
  const double eta = P.eta;
  double hn[HM_MAX];
  int it = 0;
  while (it < 8) {
    bool conv = true;
    for (int i = 0; i < n; ++i) {
      const double hj = fmax(h[i], 1.0e-12);
      const double inv = hs_rcp(hj * hj);
      const double c = inv * NB_INV_PI;
      const double nih2 = -inv;
      const double xi = (i == pb && !on_y) ? pv : X[i], yi = (i == pb && on_y) ? pv : Y[i];
      double S = 0.0;
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const double xj = (j == pb && !on_y) ? pv : X[j], yj = (j == pb && on_y) ? pv : Y[j];
        const double dx = xi - xj, dy = yi - yj;
        const double arg = (dx * dx + dy * dy) * nih2;
        if (arg > -746.0) S += M[j] * (c * hs_exp(arg));
      }
      const double Si = fmax(S, 1.0e-30);
      double v = eta * sqrt(hs_div(M[i], Si));
      if (!is_finite(v) || v <= 0.0) v = h[i];
      if (v < flo) v = flo;
      else if (v > cap) v = cap;
      conv = conv && (fabs(v - h[i]) < 1.0e-6 * hj);
      hn[i] = v;
    }
    for (int i = 0; i < n; ++i) h[i] = hn[i];
    ++it;
    if (conv) break;
  }
  return it;
}

// hamsoft_eps_model.py:240-289: soft-min of the h_i with temperature alpha_run (+ clamp under the soft policy)
__device__ __noinline__ double hm_softmin(const double* h, int n, const HsPar& P, double inv_alpha) {
  double tmax = -h[0] * inv_alpha;
  for (int i = 1; i < n; ++i) tmax = fmax(tmax, -h[i] * inv_alpha);
  double s = 0.0;
  for (int i = 0; i < n; ++i) {
    const double d = -h[i] * inv_alpha - tmax;
    s += (d == 0.0) ? 1.0 : ((d > -746.0) ? hs_exp(d) : 0.0);
  }
  double es;
  if (s <= 0.0 || !is_finite(s)) es = P.s0;
  else es = -P.alpha * (tmax + hs_log(s));
  if (P.policy == 0) {
    double lo = P.eps_min, hi = P.eps_max;
    if (hi < lo) { const double t2 = lo; lo = hi; hi = t2; }
    if (es < lo) es = lo;
    else if (es > hi) es = hi;
  }
  return es;
}

// eps_target of the unperturbed configuration, serially on the calling thread (setup, energy taps)
__device__ __noinline__ double hm_eps_target(const double* X, const double* Y, const double* M, int n, double eps_cur,
                                             const HsPar& P, double* h) {
  hm_solve(X, Y, M, n, -1, false, 0.0, eps_cur, P, h);
  return hm_softmin(h, n, P, 1.0 / P.alpha);
}

// hamsoft_eps_model.py:451-556 analytic SPH gradient + softening.py:86-131 sign reference, one body per thread.
// sh.h holds the smoothing lengths of the unperturbed configuration, sh.rs the pair separations.
__device__ __noinline__ void hm_fallback(HmSh& sh, int n) {
  const HsPar& P = sh.P;
  const int tid = threadIdx.x;
  const bool mine = tid < n;
  const int i = mine ? tid : 0;
  const int np = n * (n - 1) / 2;
  const double ninf = __longlong_as_double(0xfff0000000000000LL);
  const double xi = sh.x[i], yi = sh.y[i], mi = sh.m[i];
  const double flo = fmax(P.eps_min, 1.0e-12);
  const double hmin = fmax(1.0e-12, 0.1 * flo);
  const double ti = -sh.h[i] * sh.inv_alpha;
  const double tmax = hm_max(sh, mine ? ti : ninf);
  const double di = ti - tmax;
  const double ei = (di == 0.0) ? 1.0 : ((di > -746.0) ? hs_exp(di) : 0.0);
  __syncthreads();
  if (mine) sh.bx[i] = ei;
  __syncthreads();
  double den = 0.0;
  for (int k = 0; k < n; ++k) den += sh.bx[k];
  const bool bad = (den <= 0.0 || !is_finite(den));
  const double wi = bad ? 0.0 : hs_div(ei, den);
  const double hj = fmax(sh.h[i], hmin);
  const double ihj = hs_rcp(hj), ih2 = ihj * ihj;
  const double c = ih2 * NB_INV_PI;
  double S = 0.0, Sd = 0.0;
  double W[HM_MAX];
  for (int j = 0; j < n; ++j) {
    W[j] = 0.0;
    if (j == i) continue;
    const double dx = xi - sh.x[j], dy = yi - sh.y[j];
    const double rr = dx * dx + dy * dy;
    const double arg = -rr * ih2;
    if (arg > -746.0) {
      W[j] = c * hs_exp(arg);
      S += sh.m[j] * W[j];
      Sd += sh.m[j] * (W[j] * (2.0 * ihj * (rr * ih2 - 1.0)));      // dW/dh = W (-2/h + 2 r^2/h^3)
    }
  }
  const double Si = fmax(S, 1.0e-30);
  double Om = (Sd == 0.0) ? 1.0 : 1.0 + hs_div(hj * Sd, 2.0 * Si);
  if (!is_finite(Om) || Om == 0.0) Om = 1.0;
  const double si = wi * hs_div(hj, 2.0 * Si * Om);           // -w_i P_i with P_i = -h / (2 Sigma Omega)
  __syncthreads();
  if (mine) { sh.bx[i] = si; sh.by[i] = ih2; }
  __syncthreads();
  // the reference scatters s_i m_j coef_i(r_ij) (q_i - q_j) onto i (+) and j (-); gathered per body here
  double gx = 0.0, gy = 0.0;
  for (int j = 0; j < n; ++j) {
    if (j == i || W[j] == 0.0) continue;
    const double rx = xi - sh.x[j], ry = yi - sh.y[j];
    const double coef = -2.0 * W[j] * ih2;
    gx += si * sh.m[j] * (coef * rx);
    gy += si * sh.m[j] * (coef * ry);
  }
  for (int a = 0; a < n; ++a) {
    if (a == i) continue;
    const double ia2 = sh.by[a], sa = sh.bx[a];
    const double rx = sh.x[a] - xi, ry = sh.y[a] - yi;
    const double arg = -(rx * rx + ry * ry) * ia2;
    if (arg > -746.0) {
      const double Wa = (ia2 * NB_INV_PI) * hs_exp(arg);
      const double coef = -2.0 * Wa * ia2;
      gx -= sa * mi * (coef * rx);
      gy -= sa * mi * (coef * ry);
    }
  }
  if (bad || !is_finite(gx)) gx = 0.0;
  if (bad || !is_finite(gy)) gy = 0.0;
  // sign alignment against the legacy gradient: only the SIGN of sum(g_use . g_legacy) is used, and only when the
  // analytic gradient is non-zero
  const bool any_g = hm_max(sh, (mine && (gx != 0.0 || gy != 0.0)) ? 1.0 : 0.0) != 0.0;
  double sg = 1.0;
  {
    double dp = 0.0;
    for (int p = tid; p < np; p += HM_THREADS) dp += hs_rcp(fmax(sh.rs[p], 1.0e-15) + 1.0e-12);
    const double D = hm_sum(sh, dp);
    const double cp = P.lam * ((double)n / (D * D));
    double sx = 0.0, sy = 0.0;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      const int a = i < j ? i : j, b = i < j ? j : i;
      const double r = fmax(sh.rs[hm_pair_index(n, a, b)], 1.0e-15);
      const double dn = r + 1.0e-12;
      const double A = hs_rcp(r * dn * dn);
      sx += A * (xi - sh.x[j]);
      sy += A * (yi - sh.y[j]);
    }
    const double lx = -cp * sx, ly = -cp * sy;
    const double nok = hm_max(sh, (mine && !(is_finite(lx) && is_finite(ly))) ? 1.0 : 0.0);
    const double dot = hm_sum(sh, mine ? gx * lx + gy * ly : 0.0);
    if (any_g && is_finite(D) && D > 0.0 && nok == 0.0 && is_finite(dot) && dot < 0.0) sg = -1.0;
  }
  if (mine) { sh.gx[i] = sg * gx; sh.gy[i] = sg * gy; }
}

// eps*(q) and its gradient (hamsoft_eps_model.py:94-234) for the system in `sh`: evaluation e = 0 is the unperturbed
// configuration, e = 1 + 2 c + s perturbs coordinate c = 2 i + axis by +h (s = 0) / -h (s = 1).  The gradient lands in
// sh.gx / sh.gy; returns eps* on every thread.
__device__ __noinline__ double hm_eps_star_and_grad(HmSh& sh, int n, double eps_cur, bool& used_fallback, int& sweeps) {
  const int tid = threadIdx.x;
  const HsPar& P = sh.P;
  const int ne = 4 * n + 1, np = n * (n - 1) / 2;
  __syncthreads();
  for (int e = tid; e < ne; e += HM_THREADS) {
    int pb = -1;
    bool on_y = false;
    double pv = 0.0;
    if (e >= 1) {
      const int c = (e - 1) >> 1;
      const double sgn = ((e - 1) & 1) ? -1.0 : 1.0;
      pb = c >> 1;
      on_y = (c & 1) != 0;
      const double bv = on_y ? sh.y[pb] : sh.x[pb];
      pv = bv + sgn * hs_fd_step(bv);
    }
    double h[HM_MAX];
    sweeps += hm_solve(sh.x, sh.y, sh.m, n, pb, on_y, pv, eps_cur, P, h);
    if (e == 0)
      for (int i = 0; i < n; ++i) sh.h[i] = h[i];
    sh.f[e] = hm_softmin(h, n, P, sh.inv_alpha);
  }
  const bool mine = tid < n;
  const int i = mine ? tid : 0;
  const double xi = sh.x[i], yi = sh.y[i];
  if (mine)
    for (int j = i + 1; j < n; ++j) {
      const double dx = xi - sh.x[j], dy = yi - sh.y[j];
      sh.rs[hm_pair_index(n, i, j)] = sqrt(dx * dx + dy * dy);
    }
  __syncthreads();
  const double es = sh.f[0];
  const int src = 1 + 4 * i;
  // a clamped eps* has an exactly zero central difference (the common case under the soft policy)
  const double dfx = sh.f[src] - sh.f[src + 1], dfy = sh.f[src + 2] - sh.f[src + 3];
  double gx = (dfx == 0.0) ? 0.0 : dfx / (2.0 * hs_fd_step(xi));
  double gy = (dfy == 0.0) ? 0.0 : dfy / (2.0 * hs_fd_step(yi));
  if (!is_finite(gx)) gx = 0.0;
  if (!is_finite(gy)) gy = 0.0;
  const double gmax = sqrt(hm_max(sh, mine ? gx * gx + gy * gy : 0.0));
  if (mine) { sh.gx[i] = gx; sh.gy[i] = gy; }
  // fallback test: gmax <= 1e-12 or gmax <= 1e-9 median(r_ij).  The median costs a rank selection over up to 2,016
  // separations, so it is formed only when the test can depend on it (median <= max).
  bool fb = gmax <= 1.0e-12;
  if (!fb) {
    double rmx = 0.0;
    for (int p = tid; p < np; p += HM_THREADS) rmx = fmax(rmx, sh.rs[p]);
    const double rmax = hm_max(sh, rmx);
    if (gmax <= 1.0e-9 * rmax) {
      const double ninf = __longlong_as_double(0xfff0000000000000LL);
      const int klo = (np - 1) / 2, khi = np / 2;
      double lo_c = ninf, hi_c = ninf;
      for (int p = tid; p < np; p += HM_THREADS) {
        const double v = sh.rs[p];
        int less = 0, eq = 0;
        for (int r = 0; r < np; ++r) { const double w = sh.rs[r]; less += w < v; eq += w == v; }
        if (less <= klo && klo < less + eq) lo_c = v;
        if (less <= khi && khi < less + eq) hi_c = v;
      }
      const double lo_v = hm_max(sh, lo_c), hi_v = hm_max(sh, hi_c);
      const double rmed = (lo_v > ninf && hi_v > ninf) ? 0.5 * (lo_v + hi_v) : 0.0;
      fb = gmax <= 1.0e-9 * rmed;
    }
  }
  used_fallback = fb;                          // CTA-uniform
  if (fb) hm_fallback(sh, n);
  __syncthreads();
  return es;
}

// S half-flow: hamsoft_stepper.py:47-88 -> spring_oscillation (live definition) hamsoft_flows.py:427-762
__device__ __forceinline__ void hm_s_half(HmSh& sh, int n, double& eps, double& pi, double h, int& sweeps) {
  const HsPar& P = sh.P;
  const int tid = threadIdx.x;
  const double dt = 0.5 * h;
  double eps0 = eps, pi0 = pi;
  hs_fold(eps0, pi0, P);                         // hamsoft_stepper.py:107-113
  if (P.flags & NB_HS_FLAG_FREEZE_S) {           // SimConfig.freeze_s_subsystem (hamsoft_stepper.py:119-124); CTA-uniform
    eps = eps0;
    pi = pi0;
    return;
  }
  bool fb;
  const double es = hm_eps_star_and_grad(sh, n, eps0, fb, sweeps);
  const HsSpring& R = sh.spr;
  const double k = P.k;
  const double sn = R.sn, cs = R.cs;
  const double kick1 = (P.policy == 0) ? 0.5 * dt * hs_barrier_force(eps0, P) : 0.0;
  const double D0 = eps0 - es;
  const double pin = pi0 + kick1;
  double dlt, eta_t, I;
  if (R.rot) {
    dlt = D0 * cs + (pin * R.inv_mu_om) * sn;
    eta_t = pin * cs - R.mo * D0 * sn;
    I = R.den_ok ? (D0 * R.inv_om) * sn + (pin * R.inv_den) * (1.0 - cs) : 0.0;
  } else {
    dlt = D0; eta_t = pin; I = 0.0;
  }
  double eps_rot = es + dlt;
  const double kick2 = (P.policy == 0) ? 0.5 * dt * hs_barrier_force(eps_rot, P) : 0.0;
  const double J = k * I;
  const bool mine = tid < n;
  const int i = mine ? tid : 0;
  const double mi = sh.m[i], vx = sh.vx[i], vy = sh.vy[i], gx = sh.gx[i], gy = sh.gy[i];
  const double px = mi * vx, py = mi * vy;
  const double jx = J * gx, jy = J * gy;
  const double pmax2 = hm_max(sh, mine ? px * px + py * py : 0.0);
  const double dmax2 = hm_max(sh, mine ? jx * jx + jy * jy : 0.0);
  const double p_scale = fmax(sqrt(pmax2), 1.0e-12);
  const double dp_inf = sqrt(dmax2);
  const double thr = P.jcap * p_scale;
  const double Ja = (dp_inf > thr && dp_inf > 0.0) ? J * (thr / dp_inf) : J;
  if (mine && Ja != 0.0) {                      // p += J grad eps*; v = p / m
    const double im = hs_rcp(mi);
    sh.vx[i] = (mi * vx + Ja * gx) * im;
    sh.vy[i] = (mi * vy + Ja * gy) * im;
  }
  double pi_out = eta_t + kick2;
  hs_fold(eps_rot, pi_out, P);                   // hamsoft_stepper.py:72-80
  eps = eps_rot;
  pi = pi_out;
}

// V half-kick: hamsoft_stepper.py:543-663 + pi_half_kick hamsoft_flows.py:1102-1132; body i on thread i
__device__ __forceinline__ void hm_v_half(HmSh& sh, int n, double eps, double& pi, double G, double h) {
  const HsPar& P = sh.P;
  const int tid = threadIdx.x;
  const double hh = 0.5 * h;
  const double e2 = eps * eps;
  const bool mine = tid < n;
  const int i = mine ? tid : 0;
  __syncthreads();
  const double xi = sh.x[i], yi = sh.y[i], mi = sh.m[i];
  double fx = 0.0, fy = 0.0, s3 = 0.0;
  if (G != 0.0) {
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      const double mj = sh.m[j];
      const double dx = xi - sh.x[j], dy = yi - sh.y[j];
      const double w = rsqrt_f64<true>(fma(dx, dx, fma(dy, dy, e2)));
      const double w3 = w * w * w;
      const double mlo = i < j ? mi : mj, mhi = i < j ? mj : mi;
      const double mm = G * mlo * mhi * w3;
      fx -= mm * dx;
      fy -= mm * dy;
      if (j > i) s3 += mlo * mhi * w3;
    }
  }
  if (mine) {
    const double im = hs_rcp(mi);
    sh.vx[i] = (mi * sh.vx[i] + hh * fx) * im;
    sh.vy[i] = (mi * sh.vy[i] + hh * fy) * im;
  }
  const double s3t = hm_sum(sh, mine ? s3 : 0.0);
  const double dU = (eps == 0.0 || G == 0.0) ? 0.0 : G * eps * s3t;
  const double dB = (P.policy == 0) ? -hs_barrier_force(eps, P) : 0.0;
  if (!(P.flags & NB_HS_FLAG_FREEZE_S)) pi = pi - (dU + dB) * hh;            // hamsoft_stepper.py:592-600
}

__device__ __forceinline__ void hm_t_drift(HmSh& sh, int n, double h) {
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid < n) {
    sh.x[tid] = fma(h, sh.vx[tid], sh.x[tid]);
    sh.y[tid] = fma(h, sh.vy[tid], sh.y[tid]);
  }
}

// S(h/2) V(h/2) T(h) V(h/2) S(h/2), hamsoft_stepper.py:247-308
__device__ __forceinline__ void hm_strang(HmSh& sh, int n, double& eps, double& pi, double G, double h, int& sweeps) {
  const HsPar& P = sh.P;
  hs_fold(eps, pi, P);                           // hamsoft_stepper.py:261-264
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    hm_s_half(sh, n, eps, pi, h, sweeps);
    if (half == 0 && !(P.flags & NB_HS_FLAG_S_ONLY)) {             // cfg._validate_S_only: S S (hamsoft_stepper.py:270-284)
      hm_v_half(sh, n, eps, pi, G, h);
      hm_t_drift(sh, n, h);
      hm_v_half(sh, n, eps, pi, G, h);
    }
  }
  hs_fold(eps, pi, P);                           // hamsoft_stepper.py:300-303
}

// diagnostics.py:457-549: T + V (double-double, each rounded to fp64) + pi^2/2mu + k/2 (eps-eps*)^2 + S_bar
__device__ __noinline__ double hm_energy(const HmSh& sh, int n, double eps, double pi, double es, double G) {
  const HsPar& P = sh.P;
  dd T = dd_make(0.0);
  for (int i = 0; i < n; ++i) {
    dd v2 = dd_add(two_prod(sh.vx[i], sh.vx[i]), two_prod(sh.vy[i], sh.vy[i]));
    T = dd_add(T, dd_mul_d(dd_mul_d(v2, sh.m[i]), 0.5));
  }
  dd S = dd_make(0.0);
  const dd e2 = two_prod(eps, eps);
  if (G != 0.0)
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) {
        dd dx = two_sum(sh.x[i], -sh.x[j]);
        dd dy = two_sum(sh.y[i], -sh.y[j]);
        dd r2 = dd_add(dd_add(dd_mul(dx, dx), dd_mul(dy, dy)), e2);
        if (!(r2.hi > 0.0)) r2 = dd_make(1e-300);
        S = dd_add(S, dd_mul(two_prod(sh.m[i], sh.m[j]), dd_div(dd_make(1.0), dd_sqrt(r2))));
      }
  const double Tf = dd_to_double(T), Vf = dd_to_double(dd_mul_d(S, -G));
  const double K = 0.5 * (pi * pi) / P.mu;
  const double d = eps - es;
  const double Sp = 0.5 * (P.k * (d * d));
  return Tf + Vf + K + Sp + hs_barrier_energy(eps, P);
}

__device__ __forceinline__ void hm_load_system(HmSh& sh, int n, const double* m, const double* q, const double* v,
                                               const double* hs, int sys) {
  const int tid = threadIdx.x;
  if (tid < n) {
    sh.m[tid] = m[(size_t)sys * n + tid];
    sh.x[tid] = q[((size_t)sys * n + tid) * 2 + 0];
    sh.y[tid] = q[((size_t)sys * n + tid) * 2 + 1];
    sh.vx[tid] = v ? v[((size_t)sys * n + tid) * 2 + 0] : 0.0;
    sh.vy[tid] = v ? v[((size_t)sys * n + tid) * 2 + 1] : 0.0;
    sh.gx[tid] = 0.0; sh.gy[tid] = 0.0; sh.h[tid] = 0.0;
    sh.drx[tid] = 0.0; sh.dry[tid] = 0.0; sh.dvx[tid] = 0.0; sh.dvy[tid] = 0.0;
  }
  if (tid == 0) {
    sh.P = hs_load(hs + (size_t)sys * NB_HS_NPARAM);
    sh.inv_alpha = 1.0 / sh.P.alpha;
    for (int k = 0; k < HS_NACC; ++k) sh.acc[k] = 0.0;
    sh.acc[HA_COM_MAX] = -1.0; sh.acc[HA_VAR_MAX] = -1.0; sh.acc[HA_COS_MIN] = 2.0;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    sh.acc[HA_E0] = qnan; sh.acc[HA_L0] = qnan; sh.acc[HA_E1] = qnan; sh.acc[HA_L1] = qnan;
  }
}

// ---------------------------------------------------------------------------------------------
// run kernel: one CTA per system
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HM_THREADS) hamsoft_mid_run_kernel(HsArgs a, int n) {
  __shared__ HmSh sh;
  const int tid = threadIdx.x;
  const int sys = a.perm ? a.perm[blockIdx.x] : blockIdx.x;
  stamp_begin(a.tstamp);
  hm_load_system(sh, n, a.m, a.q, a.v, a.hs, sys);
  double eps = a.eps_pi[2 * (size_t)sys];
  double pi = a.eps_pi[2 * (size_t)sys + 1];
  const double G = a.G;
  const bool s_only = ((int)a.hs[(size_t)sys * NB_HS_NPARAM + NB_HS_FLAGS] & NB_HS_FLAG_S_ONLY) != 0;
  const int n_sub = s_only ? 1 : max(1, a.n_sub ? a.n_sub[sys] : 1);   // hamiltonian_softening_integrator.py:804-835
  const double h = a.dt / (double)n_sub;
  const double dt = a.dt;
  __syncthreads();
  if (tid == 0) {
    // hamiltonian_softening_integrator.py:232-242: mu is raised to k (dt/theta_imp)^2 on the first step
    if (a.n_steps + a.n_megno > 0 && is_finite(sh.P.k) && sh.P.k > 0.0) {
      const double mu_macro = sh.P.k * (fabs(dt) / sh.P.theta_imp) * (fabs(dt) / sh.P.theta_imp);
      if (sh.P.mu < mu_macro) sh.P.mu = mu_macro;
    }
    hs_spring_setup(sh.spr, sh.P, h);
  }
  __syncthreads();
  const bool want_energy = (a.flags & NB_RUN_ENERGY) != 0;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  int sweeps = 0, next_sample = 0;
  double tt = 0.0, accum = 0.0;
  const int n_total = a.n_steps + a.n_megno;
  const bool mine = tid < n;
  const int bi = mine ? tid : 0;
#pragma unroll 1
  for (int step = 0; step <= n_total; ++step) {
    if (want_energy && (step == 0 || step == a.n_steps)) {
      __syncthreads();
      if (tid == 0) {
        double hloc[HM_MAX];
        const double es = hm_eps_target(sh.x, sh.y, sh.m, n, eps, sh.P, hloc);
        const double E = hm_energy(sh, n, eps, pi, is_finite(es) ? es : sh.P.s0, G);
        double Lz = 0.0;
        for (int i = 0; i < n; ++i) Lz += sh.m[i] * (sh.x[i] * sh.vy[i] - sh.y[i] * sh.vx[i]);
        if (step == 0) { sh.acc[HA_E0] = E; sh.acc[HA_L0] = Lz; }
        if (step == a.n_steps) { sh.acc[HA_E1] = E; sh.acc[HA_L1] = Lz; }
      }
      __syncthreads();
    }
    if (step == a.n_steps && a.n_megno > 0) {     // evolution_features.py:37-44
      __syncthreads();
      if (tid == 0) {
        double M = 0.0, cx = 0.0, cy = 0.0, ux = 0.0, uy = 0.0;
        for (int i = 0; i < n; ++i) {
          const double rx = a.raw_dr[((size_t)sys * n + i) * 2 + 0], ry = a.raw_dr[((size_t)sys * n + i) * 2 + 1];
          const double wx = a.raw_dv[((size_t)sys * n + i) * 2 + 0], wy = a.raw_dv[((size_t)sys * n + i) * 2 + 1];
          sh.drx[i] = rx; sh.dry[i] = ry; sh.dvx[i] = wx; sh.dvy[i] = wy;
          M += sh.m[i];
          cx += sh.m[i] * rx; cy += sh.m[i] * ry; ux += sh.m[i] * wx; uy += sh.m[i] * wy;
        }
        cx /= M; cy /= M; ux /= M; uy /= M;
        double nr = 0.0, nv = 0.0;
        for (int i = 0; i < n; ++i) {
          sh.drx[i] -= cx; sh.dry[i] -= cy; sh.dvx[i] -= ux; sh.dvy[i] -= uy;
          nr += sh.drx[i] * sh.drx[i] + sh.dry[i] * sh.dry[i]; nv += sh.dvx[i] * sh.dvx[i] + sh.dvy[i] * sh.dvy[i];
        }
        nr = sqrt(nr); nv = sqrt(nv);
        for (int i = 0; i < n; ++i) { sh.drx[i] /= nr; sh.dry[i] /= nr; sh.dvx[i] /= nv; sh.dvy[i] /= nv; }
      }
      __syncthreads();
    }
    if (step == n_total) break;
#pragma unroll 1
    for (int k = 0; k < n_sub; ++k) hm_strang(sh, n, eps, pi, G, h, sweeps);
    __syncthreads();
    if (step < a.n_steps) {
      if (a.sample_interval > 0 && step == next_sample) {   // diagnostics.py:241-285
        next_sample += a.sample_interval;
        if (tid == 0) {
          double* A = sh.acc;
          double cx = 0.0, cy = 0.0, Lt = 0.0;
          for (int i = 0; i < n; ++i) {
            cx += sh.m[i] * sh.x[i]; cy += sh.m[i] * sh.y[i];
            Lt += sh.m[i] * (sh.x[i] * sh.vy[i] - sh.y[i] * sh.vx[i]);
          }
          const double mean = Lt / n;
          double var = 0.0;
          for (int i = 0; i < n; ++i) {
            const double Li = sh.m[i] * (sh.x[i] * sh.vy[i] - sh.y[i] * sh.vx[i]);
            var += (Li - mean) * (Li - mean);
          }
          var /= n;
          hs_sample_scalars(A, sqrt(cx * cx + cy * cy), var, Lt, eps, pi, sh.P.mu);
        }
      }
    } else {
      // tangent map, body i on thread i: dr += dv dt; dv += da(dr) dt  (tangent_map.py:21-59 at the post-step epsilon)
      double drx = sh.drx[bi], dry = sh.dry[bi];
      drx = fma(sh.dvx[bi], dt, drx); dry = fma(sh.dvy[bi], dt, dry);
      __syncthreads();
      if (mine) { sh.drx[bi] = drx; sh.dry[bi] = dry; }
      __syncthreads();
      const double xi = sh.x[bi], yi = sh.y[bi], e2 = eps * eps;
      double dax = 0.0, day = 0.0;
      for (int j = 0; j < n; ++j) {
        if (j == bi) continue;
        const double dx = xi - sh.x[j], dy = yi - sh.y[j];
        const double w = rsqrt_f64<true>(fma(dx, dx, fma(dy, dy, e2)));
        const double w2 = w * w, w3 = w2 * w;
        const double ex = sh.drx[j] - drx, ey = sh.dry[j] - dry;
        const double dot = -fma(dx, ex, dy * ey);
        const double c5 = 3.0 * dot * w2 * w3;
        const double gmj = G * sh.m[j];
        dax = fma(gmj, fma(ex, w3, c5 * dx), dax);
        day = fma(gmj, fma(ey, w3, c5 * dy), day);
      }
      double dvx = fma(dax, dt, sh.dvx[bi]), dvy = fma(day, dt, sh.dvy[bi]);
      double nr = sqrt(hm_sum(sh, mine ? drx * drx + dry * dry : 0.0));
      tt += dt;
      if (nr < 1e-12) {
        drx /= nr; dry /= nr; dvx /= nr; dvy /= nr;
        nr = 1.0;
      }
      const double nv = sqrt(hm_sum(sh, mine ? dvx * dvx + dvy * dvy : 0.0));
      __syncthreads();
      if (mine) { sh.drx[bi] = drx; sh.dry[bi] = dry; sh.dvx[bi] = dvx; sh.dvy[bi] = dvy; }
      accum += (nv / nr) * tt * dt;
    }
  }
  double megno = 2.0, lyap = inf, t_end = 0.0;
  if (a.n_megno > 0) {
    megno = 2.0 * accum / tt;
    lyap = (megno == 0.0) ? inf : tt / fabs(megno);
    t_end = tt;
  }
  const double sweeps_tot = hm_sum(sh, (double)sweeps);
  __syncthreads();
  stamp_end(a.tstamp);
  if (a.flags & NB_RUN_WRITE_STATE) {
    if (mine) {
      a.q[((size_t)sys * n + tid) * 2 + 0] = sh.x[tid]; a.q[((size_t)sys * n + tid) * 2 + 1] = sh.y[tid];
      a.v[((size_t)sys * n + tid) * 2 + 0] = sh.vx[tid]; a.v[((size_t)sys * n + tid) * 2 + 1] = sh.vy[tid];
    }
    if (tid == 0) { a.eps_pi[2 * (size_t)sys] = eps; a.eps_pi[2 * (size_t)sys + 1] = pi; }
  }
  if (tid != 0) return;
  bool finite = is_finite(eps) && is_finite(pi);
  for (int i = 0; i < n; ++i)
    finite = finite && is_finite(sh.x[i]) && is_finite(sh.y[i]) && is_finite(sh.vx[i]) && is_finite(sh.vy[i]);
  int st = finite ? 0 : NB_STATUS_NONFINITE;
  {
    const double R = sh.P.eps_max - sh.P.eps_min;
    if (finite && (eps < sh.P.eps_min - R || eps > sh.P.eps_max + R)) st |= NB_STATUS_EPS_OOB;
  }
  if (a.status) a.status[sys] = st;
  if (a.work) {
    a.work[2 * (size_t)sys] = sweeps_tot;
    a.work[2 * (size_t)sys + 1] = 2.0 * (double)n_sub * (double)n_total;
  }
  if (a.dyn) hs_write_dyn(a.dyn + (size_t)sys * NB_N_DYN, sh.acc, want_energy, megno, lyap, t_end);
}

// ---------------------------------------------------------------------------------------------
// setup kernel (one CTA per system, thread 0 works): constructor calibration and the frozen sub-step schedule
//   flags bit0: calibrate (hamsoft_eps_model.py:645-729 + hamiltonian_softening_integrator.py:251-296)
//   flags bit1: freeze the production schedule for step size dt (:986-1221) -> n_sub
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) hamsoft_mid_setup_kernel(const double* m_, const double* q_, double G, int n,
                                                               unsigned flags, double dt, double* hs, double* eps_pi,
                                                               int32_t* n_sub) {
  __shared__ double m[HM_MAX], x[HM_MAX], y[HM_MAX];
  const int sys = blockIdx.x, tid = threadIdx.x;
  if (tid < n) {
    m[tid] = m_[(size_t)sys * n + tid];
    x[tid] = q_[((size_t)sys * n + tid) * 2 + 0];
    y[tid] = q_[((size_t)sys * n + tid) * 2 + 1];
  }
  __syncthreads();
  if (tid != 0) return;
  double* hp = hs + (size_t)sys * NB_HS_NPARAM;
  HsPar P = hs_load(hp);
  double eps = eps_pi[2 * (size_t)sys];
  const double pinf = __longlong_as_double(0x7ff0000000000000LL);
  auto tau_grav = [&](double fallback) {
    double tau = pinf;
    if (G != 0.0) {
      for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
          const double dx = x[j] - x[i], dy = y[j] - y[i];
          const double r2 = dx * dx + dy * dy + eps * eps;
          if (r2 > 0.0 && is_finite(r2)) {
            const double r = sqrt(r2);
            const double om = sqrt(G * (m[i] + m[j]) / (r2 * r));
            if (is_finite(om) && om > 0.0) tau = fmin(tau, 1.0 / om);
          }
        }
    }
    if (!is_finite(tau) || tau <= 0.0) tau = fallback;
    return tau;
  };
  double h[HM_MAX];
  if (flags & 1u) {
    hm_solve(x, y, m, n, -1, false, 0.0, eps, P, h);
    double lo_v = 0.0, hi_v = 0.0;             // median of h
    const int klo = (n - 1) / 2, khi = n / 2;
    for (int p = 0; p < n; ++p) {
      int less = 0, eq = 0;
      for (int r = 0; r < n; ++r) { less += h[r] < h[p]; eq += h[r] == h[p]; }
      if (less <= klo && klo < less + eq) lo_v = h[p];
      if (less <= khi && khi < less + eq) hi_v = h[p];
    }
    double med = 0.5 * (lo_v + hi_v);
    const double a_seed = P.alpha > 0.0 ? P.alpha : fmax(eps, 1e-12);   // on entry ALPHA_RUN holds cfg.alpha
    if (!is_finite(med) || med <= 0.0) med = a_seed;
    double arun = 0.3 * med;
    if (!is_finite(arun) || arun <= 0.0) arun = a_seed;
    double cand = 0.25 * med;
    const double emin0 = (is_finite(P.eps_min) && P.eps_min >= 0.0) ? P.eps_min : 0.0;
    const double emax = (is_finite(P.eps_max) && P.eps_max > 0.0) ? P.eps_max : 10.0 * P.s0;
    if (!is_finite(cand)) cand = emin0;
    if (cand > emax) cand = emax;
    double enew = emin0 >= cand ? emin0 : cand;
    if (enew > emax) enew = emax;
    P.alpha = arun;
    P.eps_min = enew;
    if (eps < enew) eps = enew;
    if (!(is_finite(P.k) && P.k > 0.0)) {
      double M = 0.0;
      for (int i = 0; i < n; ++i) M += m[i];
      const double em = (is_finite(P.eps_min) && P.eps_min > 0.0) ? P.eps_min : fmax(P.s0 * 0.1, 1e-12);
      P.k = 8.0 * G * M * M / (em * em * em);
    }
    const double tau = tau_grav(1.0);
    const double om = tau > 0.0 ? 8.0 / tau : 0.0;
    double mu = om > 0.0 ? (P.k > 0.0 ? P.k / (om * om) : 1.0) : 1.0;
    if (!is_finite(mu) || mu <= 0.0) mu = 1.0;
    P.mu = mu;
    P.omega0 = om;
    hp[NB_HS_ALPHA_RUN] = P.alpha; hp[NB_HS_EPS_MIN] = P.eps_min; hp[NB_HS_K_SOFT] = P.k; hp[NB_HS_MU_SOFT] = P.mu;
    hp[NB_HS_OMEGA_SPR0] = P.omega0;
    eps_pi[2 * (size_t)sys] = eps;
  }
  if (flags & 2u) {
    double dt_abs = fabs(dt);
    if (!is_finite(dt_abs) || dt_abs <= 0.0) dt_abs = 1.0e-2;
    const double tau = tau_grav(dt_abs);
    double om = P.omega0;
    if (!is_finite(om) || om <= 0.0) { om = tau > 0.0 ? 8.0 / tau : 0.0; hp[NB_HS_OMEGA_SPR0] = om; }
    const double theta_cap = (is_finite(P.theta_cap) && P.theta_cap > 0.0) ? P.theta_cap : 0.1;
    const double h_g = 0.9 * tau;
    const double h_o = om > 0.0 ? theta_cap / om : pinf;
    const double h_theta = (is_finite(h_o) && h_o > 0.0) ? fmin(h_g, h_o) : h_g;
    // pi budget (hamiltonian_softening_integrator.py:1125-1221)
    double h_pi = dt_abs;
    if (is_finite(P.k) && P.k > 0.0) {
      const double es = hm_eps_target(x, y, m, n, eps, P, h);
      const double s0 = (is_finite(P.s0) && P.s0 > 0.0) ? P.s0 : 1.0;
      const double d_eff = fmax(fabs(eps - (is_finite(es) ? es : P.s0)), 1.0e-4 * s0);
      double S3 = 0.0;
      const double e2 = eps * eps;
      for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
          const double dx = x[i] - x[j], dy = y[i] - y[j];
          const double w = rsqrt_f64<true>(fma(dx, dx, fma(dy, dy, e2)));
          S3 = fma((G * m[i]) * m[j], w * w * w, S3);
        }
      const double dV = (eps == 0.0 || G == 0.0) ? 0.0 : eps * S3;
      const double dB = (P.policy == 0) ? -hs_barrier_force(eps, P) : 0.0;
      const double tot = fmax(fabs(dV + dB), 1.0e-16);
      h_pi = (2.0 * P.chi_pi * sqrt(P.k) * d_eff) / tot;
      if (!is_finite(h_pi) || h_pi < 0.0) h_pi = dt_abs;
    }
    if (!is_finite(h_pi) || h_pi <= 0.0) h_pi = dt_abs;
    double h_sub = fmin(h_theta, h_pi);
    if (!is_finite(h_sub) || h_sub <= 0.0) h_sub = dt_abs;
    double ns = ceil(dt_abs / h_sub);
    if (!(ns >= 1.0)) ns = 1.0;
    if (ns > 2.0e9) ns = 2.0e9;
    n_sub[sys] = (int)ns;
  }
}

// eps*(q), its gradient and H_ext for B systems (parity taps), one CTA per system
__global__ void __launch_bounds__(HM_THREADS) hamsoft_mid_probe_kernel(const double* m_, const double* q_,
                                                                       const double* v_, double G, int n,
                                                                       const double* eps_pi, const double* hs,
                                                                       double* out /*[B][3+2N]: eps*, H, fallback, grad*/) {
  __shared__ HmSh sh;
  const int sys = blockIdx.x, tid = threadIdx.x;
  hm_load_system(sh, n, m_, q_, v_, hs, sys);
  __syncthreads();
  const double eps = eps_pi[2 * (size_t)sys], pi = eps_pi[2 * (size_t)sys + 1];
  bool fb;
  int sweeps = 0;
  const double es = hm_eps_star_and_grad(sh, n, eps, fb, sweeps);
  double* o = out + (size_t)sys * (3 + 2 * n);
  if (tid == 0) {
    o[0] = es;
    o[1] = hm_energy(sh, n, eps, pi, es, G);
    o[2] = fb ? 1.0 : 0.0;
  }
  if (tid < n) { o[3 + 2 * tid] = sh.gx[tid]; o[4 + 2 * tid] = sh.gy[tid]; }
}

int hamsoft_mid_run(const HsArgs& a, int N, cudaStream_t st) {
  if (a.B == 0) return NB_OK;
  hamsoft_mid_run_kernel<<<a.B, HM_THREADS, 0, st>>>(a, N);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int hamsoft_mid_setup(const double* m, const double* q, double G, int B, int N, unsigned flags, double dt, double* hs,
                      double* eps_pi, int32_t* n_sub, cudaStream_t st) {
  if (B == 0) return NB_OK;
  hamsoft_mid_setup_kernel<<<B, 64, 0, st>>>(m, q, G, N, flags, dt, hs, eps_pi, n_sub);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int hamsoft_mid_probe(const double* m, const double* q, const double* v, double G, int B, int N, const double* eps_pi,
                      const double* hs, double* out, cudaStream_t st) {
  if (B == 0) return NB_OK;
  hamsoft_mid_probe_kernel<<<B, HM_THREADS, 0, st>>>(m, q, v, G, N, eps_pi, hs, out);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
