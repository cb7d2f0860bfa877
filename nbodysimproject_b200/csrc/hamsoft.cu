// hamsoft.cu -- ham_soft (Strang split with a dynamical softening length) for ensembles of small systems.
//
// Mapping: ONE WARP PER SYSTEM.  79 % of the reference's ham_soft time is the eps* model: every S half-flow
// needs eps*(q) and its gradient by central finite differences over all 2N coordinates, i.e. 4N+1 independent
// fixed-point solves (hamsoft_eps_model.py:94-234, 316-400).  Those 4N+1 <= 33 evaluations are exactly one
// per lane (N = 8 needs one extra pass); every lane keeps a full replica of the tiny system state in registers,
// so the V and T flows and the spring rotation are executed redundantly (free in SIMT) and the only
// communication is the gather of the 4N+1 scalars with full-mask shuffles.
//
// Restates (not ports): hamsoft_stepper.py:47-308, 543-663 (strang_step / s_half / v_half_kick / t_drift),
// hamsoft_flows.py:427-762 (live spring_oscillation incl. the J cap), :1102-1132 (pi_half_kick),
// hamsoft_eps_model.py (eps_target_production, _solve_hi, eps_star_and_grad, _production_grad,
// calibrate_from_initial_conditions), softening.py:86-131 (legacy gradient, sign reference),
// barrier.py:35-113, hamiltonian_softening_integrator.py:145-296, 986-1221 (mu calibration, frozen schedule),
// diagnostics.py:241-285, 457-549 (step_metrics, extended Hamiltonian).
// Barrier policies: 0 = soft (shipped default), 1 = reflection (fold eps into [eps_min, eps_max], flip pi;
// hamsoft_utils.py:150-176), 2 = none (disable_barrier).
#include "pair_small.cuh"
#include "args.cuh"

namespace nb {

#define NB_PI 3.14159265358979323846

struct HsPar {
  double k, mu, eps_min, eps_max, alpha, k_wall, eta, jcap, lam, theta_imp, theta_cap, chi_pi, omega0, s0;
  int n_exp, policy;
};

__device__ __forceinline__ HsPar hs_load(const double* p) {
  HsPar h;
  h.k = p[NB_HS_K_SOFT]; h.mu = p[NB_HS_MU_SOFT]; h.eps_min = p[NB_HS_EPS_MIN]; h.eps_max = p[NB_HS_EPS_MAX];
  h.alpha = p[NB_HS_ALPHA_RUN]; h.k_wall = p[NB_HS_K_WALL]; h.n_exp = (int)p[NB_HS_BARRIER_N]; h.eta = p[NB_HS_ETA];
  h.jcap = p[NB_HS_J_MAX_CAP]; h.lam = p[NB_HS_LAMBDA]; h.policy = (int)p[NB_HS_POLICY];
  h.theta_imp = p[NB_HS_THETA_IMP]; h.theta_cap = p[NB_HS_THETA_CAP]; h.chi_pi = p[NB_HS_CHI_PI];
  h.omega0 = p[NB_HS_OMEGA_SPR0]; h.s0 = p[NB_HS_S0];
  return h;
}

__device__ __forceinline__ double hs_ipow(double x, int e) {   // x ** e for small non-negative integer e
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= x;
  return r;
}
// barrier.py:66-113
__device__ __forceinline__ double hs_barrier_force(double eps, const HsPar& P) {
  if (P.policy != 0) return 0.0;
  if (!(is_finite(P.k_wall) && P.k_wall > 0.0)) return 0.0;
  const int n = max(2, P.n_exp);
  const double la = fmax(0.0, P.eps_min - eps), rb = fmax(0.0, eps - P.eps_max);
  const int e = n - 2;
  const double left = la > 0.0 ? (e == 0 ? 1.0 : hs_ipow(la, e)) : 0.0;
  const double right = rb > 0.0 ? (e == 0 ? 1.0 : hs_ipow(rb, e)) : 0.0;
  return P.k_wall * (left - right);
}
// barrier.py:35-63
__device__ __forceinline__ double hs_barrier_energy(double eps, const HsPar& P) {
  if (P.policy != 0) return 0.0;
  if (!(is_finite(P.k_wall) && P.k_wall > 0.0) || P.n_exp < 2) return 0.0;
  double a = P.eps_min, b = P.eps_max;
  if (b < a) { const double t = a; a = b; b = t; }
  const double left = fmax(0.0, a - eps), right = fmax(0.0, eps - b);
  const int p = P.n_exp - 1;
  return (P.k_wall / (double)p) * (hs_ipow(left, p) + hs_ipow(right, p));
}

// hamsoft_eps_model.py:316-400 -- Jacobi sweeps for the SPH-like smoothing lengths h_i, started from the
// current epsilon, at most 8 sweeps, relative tolerance 1e-6, clamped to [max(eps_min,1e-12), eps_max].
template <int N>
__device__ __forceinline__ void hs_solve_hi(const double* qx, const double* qy, const double* m, double eps_cur,
                                            const HsPar& P, double* h) {
  double lo = P.eps_min, hi = P.eps_max;
  if (hi < lo) { const double t = lo; lo = hi; hi = t; }
  const double flo = fmax(lo, 1.0e-12), cap = fmax(flo, hi);
  double h0 = eps_cur;
  if (!is_finite(h0) || h0 <= 0.0) h0 = 1.0;
  h0 = fmin(fmax(h0, flo), cap);
#pragma unroll
  for (int i = 0; i < N; ++i) h[i] = h0;
  double r2[N * (N - 1) / 2 > 0 ? N * (N - 1) / 2 : 1];
  {
    int p = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i + 1; j < N; ++j) {
        const double dx = qx[i] - qx[j], dy = qy[i] - qy[j];
        r2[p++] = dx * dx + dy * dy;
      }
  }
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    double changed = 0.0;
    double hn[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double hj = fmax(h[i], 1.0e-12);
      const double h2 = hj * hj;
      const double c = 1.0 / (NB_PI * h2);
      const double nih2 = -1.0 / h2;            // one division per body and sweep instead of one per pair
      double S = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        const int a = i < j ? i : j, b = i < j ? j : i;
        const double rr = r2[a * N - a * (a + 1) / 2 + (b - a - 1)];
        S += m[j] * (c * exp(rr * nih2));
      }
      const double Si = fmax(S, 1.0e-30);
      double v = P.eta * sqrt(m[i] / Si);
      if (!is_finite(v) || v <= 0.0) v = h[i];
      if (v < flo) v = flo;
      else if (v > cap) v = cap;
      const double rel = fabs(v - h[i]) / fmax(h[i], 1.0e-12);
      changed = fmax(changed, rel);
      hn[i] = v;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) h[i] = hn[i];
    if (changed < 1.0e-6) break;
  }
}

// hamsoft_eps_model.py:240-289 eps_target_production: soft-min of the h_i with temperature alpha_run
template <int N>
__device__ __forceinline__ double hs_eps_target(const double* qx, const double* qy, const double* m, double eps_cur,
                                                const HsPar& P) {
  double h[N];
  hs_solve_hi<N>(qx, qy, m, eps_cur, P, h);
  double tmax = -h[0] / P.alpha;
#pragma unroll
  for (int i = 1; i < N; ++i) tmax = fmax(tmax, -h[i] / P.alpha);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) s += exp(-h[i] / P.alpha - tmax);
  double es;
  if (s <= 0.0 || !is_finite(s)) es = P.s0;
  else es = -P.alpha * (tmax + log(s));
  if (P.policy == 0) {
    double lo = P.eps_min, hi = P.eps_max;
    if (hi < lo) { const double t = lo; lo = hi; hi = t; }
    if (es < lo) es = lo;
    else if (es > hi) es = hi;
  }
  return es;
}

// hamsoft_eps_model.py:451-556 analytic SPH gradient
template <int N>
__device__ __noinline__ void hs_production_grad(const double* qx, const double* qy, const double* m, double eps_cur,
                                                const HsPar& P, double* gx, double* gy) {
  double h[N];
  hs_solve_hi<N>(qx, qy, m, eps_cur, P, h);
  const double flo = fmax(P.eps_min, 1.0e-12);
  const double hmin = fmax(1.0e-12, 0.1 * flo);
  double tmax = -h[0] / P.alpha;
  for (int i = 1; i < N; ++i) tmax = fmax(tmax, -h[i] / P.alpha);
  double den = 0.0;
  for (int i = 0; i < N; ++i) den += exp(-h[i] / P.alpha - tmax);
  for (int i = 0; i < N; ++i) { gx[i] = 0.0; gy[i] = 0.0; }
  if (den <= 0.0 || !is_finite(den)) return;
  double Pi[N], w[N];
  for (int i = 0; i < N; ++i) {
    w[i] = exp(-h[i] / P.alpha - tmax) / den;
    const double hj = fmax(h[i], hmin);
    double S = 0.0, Sd = 0.0;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const double dx = qx[i] - qx[j], dy = qy[i] - qy[j];
      const double rr = dx * dx + dy * dy;
      const double c = 1.0 / (NB_PI * hj * hj);
      const double W = c * exp(-rr / (hj * hj));
      S += m[j] * W;
      Sd += m[j] * (W * (-2.0 / hj + 2.0 * rr / (hj * hj * hj)));
    }
    const double Si = fmax(S, 1.0e-30);
    double Om = 1.0 + hj * Sd / (2.0 * Si);
    if (!is_finite(Om) || Om == 0.0) Om = 1.0;
    Pi[i] = -hj / (2.0 * Si * Om);
  }
  for (int i = 0; i < N; ++i) {
    const double hj = fmax(h[i], hmin);
    const double s_i = -w[i] * Pi[i];
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const double rx = qx[i] - qx[j], ry = qy[i] - qy[j];
      const double c = 1.0 / (NB_PI * hj * hj);
      const double W = c * exp(-(rx * rx + ry * ry) / (hj * hj));
      const double coef = -2.0 * W / (hj * hj);
      gx[i] += s_i * m[j] * (coef * rx);
      gy[i] += s_i * m[j] * (coef * ry);
      gx[j] -= s_i * m[j] * (coef * rx);
      gy[j] -= s_i * m[j] * (coef * ry);
    }
  }
  for (int i = 0; i < N; ++i) {
    if (!is_finite(gx[i])) gx[i] = 0.0;
    if (!is_finite(gy[i])) gy[i] = 0.0;
  }
}

// softening.py:86-131 legacy gradient; only sum(g_use . g_legacy) is needed (sign alignment)
template <int N>
__device__ __noinline__ double hs_legacy_dot(const double* qx, const double* qy, const double* gx, const double* gy,
                                             double lam) {
  double D = 0.0;
  for (int i = 0; i < N; ++i)
    for (int j = i + 1; j < N; ++j) {
      const double dx = qx[i] - qx[j], dy = qy[i] - qy[j];
      const double r = fmax(sqrt(dx * dx + dy * dy), 1.0e-15);
      D += 1.0 / (r + 1.0e-12);
    }
  if (!is_finite(D) || D <= 0.0) return 0.0;
  const double cp = lam * ((double)N / (D * D));
  double dot = 0.0;
  bool ok = true;
  for (int i = 0; i < N; ++i) {
    double sx = 0.0, sy = 0.0;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const double dx = qx[i] - qx[j], dy = qy[i] - qy[j];
      const double r = fmax(sqrt(dx * dx + dy * dy), 1.0e-15);
      const double den = r + 1.0e-12;
      const double A = 1.0 / (r * den * den);
      sx += A * dx;
      sy += A * dy;
    }
    const double lx = -cp * sx, ly = -cp * sy;
    ok = ok && is_finite(lx) && is_finite(ly);
    dot += gx[i] * lx + gy[i] * ly;
  }
  return ok ? dot : 0.0;
}

// FD step of coordinate value x (hamsoft_eps_model.py:136-144)
__device__ __forceinline__ double hs_fd_step(double x) {
  double h = 1.0e-5 * fmax(fabs(x), 1.0);
  if (h < 1.0e-10) h = 1.0e-10;
  return h;
}

// eps*(q) and its gradient, cooperatively over the warp: lane 0 evaluates the unperturbed configuration,
// lane 1+2c+s the coordinate c = 2 i + a perturbed by +h (s = 0) or -h (s = 1).  Every lane returns the full
// gradient.  hamsoft_eps_model.py:94-234.
// Lanes per system.  4 N + 1 evaluations fit a half warp for N <= 3, so two systems share a warp there.  For N = 4
// (17) and N = 8 (33) the odd one out -- the UNPERTURBED evaluation -- is computed cooperatively instead (one body per
// lane, Jacobi sweeps exchanged by shuffles; same arithmetic per body, so the same bits as the serial solve at 1/N of
// its cost): the 4 N perturbed evaluations then fit a half warp (N = 4: two systems per warp) or one pass (N = 8).
template <int N>
struct HsLanes {
  static constexpr bool COOP = (N == 4 || N == 8);
  static constexpr int NE = COOP ? 4 * N : 4 * N + 1;        // evaluations spread one per lane
  static constexpr int LPS = (NE <= 16) ? 16 : 32;
};

// eps_target of the unperturbed configuration, cooperatively over the lanes [base, base + LPS) of one system
// (hamsoft_eps_model.py:316-400 + :240-289).  All lanes of the group hold the same x, y, m.  Warp-uniform control flow:
// the sweep loop always runs 8 times and a converged group simply stops updating (the two systems of a warp may
// converge at different sweeps, and the shuffles need the whole warp).
template <int N>
__device__ __forceinline__ double hs_eps_target_coop(const double* x, const double* y, const double* m, double eps_cur,
                                                     const HsPar& P, int lane, int base) {
  constexpr int LPS = HsLanes<N>::LPS;
  const int i = lane < N ? lane : 0;                        // spare lanes shadow body 0
  double xi = x[0], yi = y[0], mi = m[0];
#pragma unroll
  for (int k = 1; k < N; ++k) if (k == i) { xi = x[k]; yi = y[k]; mi = m[k]; }
  double lo = P.eps_min, hi = P.eps_max;
  if (hi < lo) { const double t = lo; lo = hi; hi = t; }
  const double flo = fmax(lo, 1.0e-12), cap = fmax(flo, hi);
  double h0 = eps_cur;
  if (!is_finite(h0) || h0 <= 0.0) h0 = 1.0;
  h0 = fmin(fmax(h0, flo), cap);
  double r2[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    // the serial solver stores r^2 of the pair (min, max): same value, the squares do not see the sign
    const double dx = (i < j) ? xi - x[j] : x[j] - xi, dy = (i < j) ? yi - y[j] : y[j] - yi;
    r2[j] = dx * dx + dy * dy;
  }
  double hcur = h0;
  bool done = false;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const double hj = fmax(hcur, 1.0e-12);
    const double h2 = hj * hj;
    const double c = 1.0 / (NB_PI * h2);
    const double nih2 = -1.0 / h2;
    double S = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double term = m[j] * (c * exp(r2[j] * nih2));
      if (j != i) S += term;
    }
    const double Si = fmax(S, 1.0e-30);
    double v = P.eta * sqrt(mi / Si);
    if (!is_finite(v) || v <= 0.0) v = hcur;
    if (v < flo) v = flo;
    else if (v > cap) v = cap;
    double rel = fabs(v - hcur) / fmax(hcur, 1.0e-12);
#pragma unroll
    for (int off = LPS / 2; off > 0; off >>= 1) rel = fmax(rel, __shfl_xor_sync(0xffffffffu, rel, off));
    if (!done) hcur = v;
    if (rel < 1.0e-6) done = true;                          // group-uniform
  }
  double h[N];
#pragma unroll
  for (int k = 0; k < N; ++k) h[k] = __shfl_sync(0xffffffffu, hcur, base + k);
  double tmax = -h[0] / P.alpha;
#pragma unroll
  for (int k = 1; k < N; ++k) tmax = fmax(tmax, -h[k] / P.alpha);
  double sum = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) sum += exp(-h[k] / P.alpha - tmax);
  double es;
  if (sum <= 0.0 || !is_finite(sum)) es = P.s0;
  else es = -P.alpha * (tmax + log(sum));
  if (P.policy == 0) {
    double l2 = P.eps_min, h2b = P.eps_max;
    if (h2b < l2) { const double t = l2; l2 = h2b; h2b = t; }
    if (es < l2) es = l2;
    else if (es > h2b) es = h2b;
  }
  return es;
}

template <int N>
__device__ __forceinline__ double hs_eps_star_and_grad(const double* x, const double* y, const double* m,
                                                       double eps_cur, const HsPar& P, int lane_full, double* gx,
                                                       double* gy, bool& used_fallback) {
  constexpr bool COOP = HsLanes<N>::COOP;
  constexpr int NE = HsLanes<N>::NE;
  constexpr int LPS = HsLanes<N>::LPS;
  constexpr int OFF = COOP ? 0 : 1;           // evaluation index of the first perturbed configuration
  const int lane = lane_full & (LPS - 1);     // lane within this system's group
  const int base = lane_full - lane;          // first lane of the group
  double f[2] = {0.0, 0.0};
#pragma unroll
  for (int pass = 0; pass < (NE + LPS - 1) / LPS; ++pass) {
    const int e = lane + LPS * pass;          // evaluation index
    const int ee = e < NE ? e : 0;            // idle lanes redo evaluation 0
    const bool pert = COOP || ee >= 1;        // without COOP evaluation 0 is the unperturbed configuration
    const int c = (ee - OFF) >> 1;            // perturbed coordinate
    const double sgn = ((ee - OFF) & 1) ? -1.0 : 1.0;
    double px[N], py[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      px[i] = x[i];
      py[i] = y[i];
      if (pert && c == 2 * i) px[i] = x[i] + sgn * hs_fd_step(x[i]);
      if (pert && c == 2 * i + 1) py[i] = y[i] + sgn * hs_fd_step(y[i]);
    }
    f[pass] = hs_eps_target<N>(px, py, m, eps_cur, P);
  }
  const double es = COOP ? hs_eps_target_coop<N>(x, y, m, eps_cur, P, lane, base) : __shfl_sync(0xffffffffu, f[0], base);
  double gmax2 = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int c = 2 * i + a;
      const int ep = OFF + 2 * c, em = OFF + 1 + 2 * c;
      const double fp = ep < LPS ? __shfl_sync(0xffffffffu, f[0], base + (ep & (LPS - 1)))
                                 : __shfl_sync(0xffffffffu, f[1], base + (ep & (LPS - 1)));
      const double fm = em < LPS ? __shfl_sync(0xffffffffu, f[0], base + (em & (LPS - 1)))
                                 : __shfl_sync(0xffffffffu, f[1], base + (em & (LPS - 1)));
      const double h = hs_fd_step(a == 0 ? x[i] : y[i]);
      double g = (fp - fm) / (2.0 * h);
      if (!is_finite(g)) g = 0.0;
      if (a == 0) gx[i] = g; else gy[i] = g;
    }
    gmax2 = fmax(gmax2, gx[i] * gx[i] + gy[i] * gy[i]);
  }
  const double gmax = sqrt(gmax2);
  // median pair separation
  constexpr int NP = N * (N - 1) / 2;
  double rs[NP > 0 ? NP : 1];
  {
    int p = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i + 1; j < N; ++j) {
        const double dx = x[i] - x[j], dy = y[i] - y[j];
        rs[p++] = sqrt(dx * dx + dy * dy);
      }
  }
  double rmed = 0.0;
  if (NP > 0) {
    // rank selection without dynamic indexing of a sorted copy
    double lo_v = 0.0, hi_v = 0.0;
    const int klo = (NP - 1) / 2, khi = NP / 2;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      int less = 0, eq = 0;
#pragma unroll
      for (int r = 0; r < NP; ++r) { less += rs[r] < rs[p]; eq += rs[r] == rs[p]; }
      if (less <= klo && klo < less + eq) lo_v = rs[p];
      if (less <= khi && khi < less + eq) hi_v = rs[p];
    }
    rmed = 0.5 * (lo_v + hi_v);
  }
  used_fallback = (gmax <= 1.0e-12) || (gmax <= 1.0e-9 * rmed);
  if (used_fallback) {
    double ax[N], ay[N], cx[N], cy[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { cx[i] = x[i]; cy[i] = y[i]; }
    hs_production_grad<N>(cx, cy, m, eps_cur, P, ax, ay);
    const double dot = hs_legacy_dot<N>(cx, cy, ax, ay, P.lam);
    const double sg = (is_finite(dot) && dot < 0.0) ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < N; ++i) { gx[i] = sg * ax[i]; gy[i] = sg * ay[i]; }
  }
  return es;
}

template <int N>
struct HsState {
  double m[N], x[N], y[N], vx[N], vy[N];
  double eps, pi;
};

// reflect_and_bounce(eps, pi, h = 0) = reflect_if_needed (hamsoft_barrier_controller.py:27-69, hamsoft_utils.py:105-176)
__device__ __forceinline__ void hs_fold(double& eps, double& pi, const HsPar& P) {
  if (P.policy != 1) return;
  const double a = P.eps_min, b = P.eps_max;
  const double R = b - a;
  if (!is_finite(R) || R <= 0.0) { eps = a; pi = -pi; return; }
  const double per = 2.0 * R;
  double y = fmod(eps - a, per);                 // Python float modulo: result takes the sign of the divisor
  if (y != 0.0) { if (y < 0.0) y += per; } else y = 0.0;
  if (y <= R) { eps = a + y; }
  else { eps = b - (y - R); pi = -pi; }
}

// S half-flow: hamsoft_stepper.py:47-88 -> spring_oscillation (live definition) hamsoft_flows.py:427-762
template <int N>
__device__ __forceinline__ void hs_s_half(HsState<N>& s, const HsPar& P, double h, int lane) {
  const double dt = 0.5 * h;
  double gx[N], gy[N];
  bool fb;
  hs_fold(s.eps, s.pi, P);                       // hamsoft_stepper.py:107-113
  const double es = hs_eps_star_and_grad<N>(s.x, s.y, s.m, s.eps, P, lane, gx, gy, fb);
  const double k = P.k, mu = P.mu;
  const double om = (k > 0.0 && mu > 0.0) ? sqrt(k / mu) : 0.0;
  const double th = om * dt;
  double sn, cs;
  if (fabs(th) < 1.0e-8) {
    const double t2 = th * th;
    sn = th - t2 * th / 6.0 + t2 * t2 * th / 120.0;
    cs = 1.0 - t2 / 2.0 + t2 * t2 / 24.0;
  } else {
    sincos(th, &sn, &cs);
  }
  const double eps0 = s.eps, pi0 = s.pi;
  const double kick1 = (P.policy == 0) ? 0.5 * dt * hs_barrier_force(eps0, P) : 0.0;
  const double D0 = eps0 - es;
  const double pin = pi0 + kick1;
  double dlt, eta_t, I;
  if (om != 0.0 && mu != 0.0) {
    const double mo = sqrt(mu * fmax(k, 0.0));
    dlt = D0 * cs + (pin / (mu * om)) * sn;
    eta_t = pin * cs - mo * D0 * sn;
    const double den = mu * om * om;
    I = den != 0.0 ? (D0 / om) * sn + (pin / den) * (1.0 - cs) : 0.0;
  } else {
    dlt = D0; eta_t = pin; I = 0.0;
  }
  const double eps_rot = es + dlt;
  const double kick2 = (P.policy == 0) ? 0.5 * dt * hs_barrier_force(eps_rot, P) : 0.0;
  const double J = k * I;
  double pmax2 = 0.0, dmax2 = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double px = s.m[i] * s.vx[i], py = s.m[i] * s.vy[i];
    pmax2 = fmax(pmax2, px * px + py * py);
    const double dx = J * gx[i], dy = J * gy[i];
    dmax2 = fmax(dmax2, dx * dx + dy * dy);
  }
  const double p_scale = fmax(sqrt(pmax2), 1.0e-12);
  const double dp_inf = sqrt(dmax2);
  const double thr = P.jcap * p_scale;
  const double Ja = (dp_inf > thr && dp_inf > 0.0) ? J * (thr / dp_inf) : J;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.vx[i] = (s.m[i] * s.vx[i] + Ja * gx[i]) / s.m[i];
    s.vy[i] = (s.m[i] * s.vy[i] + Ja * gy[i]) / s.m[i];
  }
  s.eps = eps_rot;
  s.pi = eta_t + kick2;
  hs_fold(s.eps, s.pi, P);                       // hamsoft_stepper.py:72-80
}

// V half-kick: hamsoft_stepper.py:543-663 + pi_half_kick hamsoft_flows.py:1102-1132
template <int N>
__device__ __forceinline__ void hs_v_half(HsState<N>& s, const HsPar& P, double G, double h) {
  const double hh = 0.5 * h;
  const double e = s.eps, e2 = e * e;
  double fx[N], fy[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { fx[i] = 0.0; fy[i] = 0.0; }
  double s3 = 0.0;
  if (G != 0.0) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i + 1; j < N; ++j) {
        const double dx = s.x[i] - s.x[j], dy = s.y[i] - s.y[j];
        const double w = rsqrt_f64<true>(fma(dx, dx, fma(dy, dy, e2)));
        const double w3 = w * w * w;
        const double mm = G * s.m[i] * s.m[j] * w3;          // force magnitude factor
        fx[i] -= mm * dx; fy[i] -= mm * dy;
        fx[j] += mm * dx; fy[j] += mm * dy;
        s3 += s.m[i] * s.m[j] * w3;
      }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.vx[i] = (s.m[i] * s.vx[i] + hh * fx[i]) / s.m[i];
    s.vy[i] = (s.m[i] * s.vy[i] + hh * fy[i]) / s.m[i];
  }
  const double dU = (e == 0.0 || G == 0.0) ? 0.0 : G * e * s3;
  const double dB = (P.policy == 0) ? -hs_barrier_force(e, P) : 0.0;
  s.pi = s.pi - (dU + dB) * hh;
}

template <int N>
__device__ __forceinline__ void hs_strang(HsState<N>& s, const HsPar& P, double G, double h, int lane) {
  // S(h/2) V(h/2) T(h) V(h/2) S(h/2).  Written as a two-trip loop so that the S half-flow (the eps* model: by far the
  // largest piece of code) is instantiated ONCE; the unrolled form made the kernel stall on instruction fetch
  // (ncu: no_instruction 1.8 warps per issue slot at N = 3).
  hs_fold(s.eps, s.pi, P);                       // hamsoft_stepper.py:261-264
  if constexpr (N >= 8) {                        // N = 8 is register-bound: the rolled form spills more than it saves
    hs_s_half<N>(s, P, h, lane);
    hs_v_half<N>(s, P, G, h);
#pragma unroll
    for (int i = 0; i < N; ++i) { s.x[i] = fma(h, s.vx[i], s.x[i]); s.y[i] = fma(h, s.vy[i], s.y[i]); }
    hs_v_half<N>(s, P, G, h);
    hs_s_half<N>(s, P, h, lane);
    hs_fold(s.eps, s.pi, P);
    return;
  }
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    hs_s_half<N>(s, P, h, lane);
    if (half == 0) {
#pragma unroll 1
      for (int kick = 0; kick < 2; ++kick) {
        hs_v_half<N>(s, P, G, h);
        if (kick == 0) {
#pragma unroll
          for (int i = 0; i < N; ++i) { s.x[i] = fma(h, s.vx[i], s.x[i]); s.y[i] = fma(h, s.vy[i], s.y[i]); }
        }
      }
    }
  }
  hs_fold(s.eps, s.pi, P);                       // hamsoft_stepper.py:300-303
}

// diagnostics.py:457-549: T + V (double-double, each rounded to fp64) + pi^2/2mu + k/2 (eps-eps*)^2 + S_bar
template <int N>
__device__ __noinline__ double hs_energy(const double* m, const double* x, const double* y, const double* vx,
                                         const double* vy, double eps, double pi, double es, const HsPar& P, double G) {
  dd T = dd_make(0.0);
  for (int i = 0; i < N; ++i) {
    dd v2 = dd_add(two_prod(vx[i], vx[i]), two_prod(vy[i], vy[i]));
    T = dd_add(T, dd_mul_d(dd_mul_d(v2, m[i]), 0.5));
  }
  dd S = dd_make(0.0);
  const dd e2 = two_prod(eps, eps);
  if (G != 0.0)
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) {
        dd dx = two_sum(x[i], -x[j]);
        dd dy = two_sum(y[i], -y[j]);
        dd r2 = dd_add(dd_add(dd_mul(dx, dx), dd_mul(dy, dy)), e2);
        if (!(r2.hi > 0.0)) r2 = dd_make(1e-300);
        S = dd_add(S, dd_mul(two_prod(m[i], m[j]), dd_div(dd_make(1.0), dd_sqrt(r2))));
      }
  const double Tf = dd_to_double(T), Vf = dd_to_double(dd_mul_d(S, -G));
  const double K = 0.5 * (pi * pi) / P.mu;
  const double d = eps - es;
  const double Sp = 0.5 * (P.k * (d * d));
  return Tf + Vf + K + Sp + hs_barrier_energy(eps, P);
}

// ---------------------------------------------------------------------------------------------
// run kernel: one warp per system
// ---------------------------------------------------------------------------------------------
struct HsArgs {
  const double* m; double* q; double* v; double G; int B; unsigned flags; double dt; int n_steps; int sample_interval;
  int n_megno; const int32_t* n_sub; const double* raw_dr; const double* raw_dv; double* eps_pi; const double* hs;
  double* dyn; int32_t* status;
};

struct Welford {
  double mean, m2; int n;
  __device__ __forceinline__ void add(double x) { ++n; const double d = x - mean; mean += d / n; m2 += d * (x - mean); }
};

template <int N>
static inline int hs_run_blocks(int B) {
  const int spw = 32 / HsLanes<N>::LPS;
  const int warps = (B + spw - 1) / spw;
  return (warps + 3) / 4;
}

template <int N>
__global__ void __launch_bounds__(128, (N <= 4 ? 4 : 2)) hamsoft_run_kernel(HsArgs a) {
  constexpr int LPS = HsLanes<N>::LPS, SPW = 32 / LPS;   // lanes per system, systems per warp
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp * SPW >= a.B) return;                        // warp-uniform
  // an odd tail slot shadows the warp's first system (same arithmetic, no writes) so the warp stays converged
  const bool live = warp * SPW + lane / LPS < a.B;
  const int sys = live ? warp * SPW + lane / LPS : warp * SPW;
  HsState<N> s;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.m[i] = a.m[(size_t)sys * N + i];
    s.x[i] = a.q[((size_t)sys * N + i) * 2 + 0];
    s.y[i] = a.q[((size_t)sys * N + i) * 2 + 1];
    s.vx[i] = a.v[((size_t)sys * N + i) * 2 + 0];
    s.vy[i] = a.v[((size_t)sys * N + i) * 2 + 1];
  }
  s.eps = a.eps_pi[2 * (size_t)sys];
  s.pi = a.eps_pi[2 * (size_t)sys + 1];
  HsPar P = hs_load(a.hs + (size_t)sys * NB_HS_NPARAM);
  const double G = a.G;
  const int n_sub = max(1, a.n_sub ? a.n_sub[sys] : 1);
  const int n_sub_warp = SPW > 1 ? __reduce_max_sync(0xffffffffu, n_sub) : n_sub;
  const double h = a.dt / (double)n_sub;
  // one macro step = n_sub Strang sub-steps; systems sharing a warp run to the larger count and discard the excess
  auto macro_step = [&]() {
#pragma unroll 1
    for (int k = 0; k < n_sub_warp; ++k) {
      if (SPW > 1) {
        HsState<N> t = s;
        hs_strang<N>(t, P, G, h, lane);
        if (k < n_sub) s = t;
      } else {
        hs_strang<N>(s, P, G, h, lane);
      }
    }
  };
  // hamiltonian_softening_integrator.py:232-242: mu is raised to k (dt/theta_imp)^2 on the first step
  if (a.n_steps + a.n_megno > 0 && is_finite(P.k) && P.k > 0.0) {
    const double mu_macro = P.k * (fabs(a.dt) / P.theta_imp) * (fabs(a.dt) / P.theta_imp);
    if (P.mu < mu_macro) P.mu = mu_macro;
  }
  const bool want_energy = (a.flags & NB_RUN_ENERGY) != 0;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);

  auto energy = [&](double& E, double& L) {
    double cx[N], cy[N], cu[N], cw[N], cm[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { cx[i] = s.x[i]; cy[i] = s.y[i]; cu[i] = s.vx[i]; cw[i] = s.vy[i]; cm[i] = s.m[i]; }
    const double es = hs_eps_target<N>(cx, cy, cm, s.eps, P);
    E = hs_energy<N>(cm, cx, cy, cu, cw, s.eps, s.pi, is_finite(es) ? es : P.s0, P, G);
    L = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) L += s.m[i] * (s.x[i] * s.vy[i] - s.y[i] * s.vx[i]);
  };
  double E0 = nan, L0 = nan, E1 = nan, L1 = nan;

  double com_sum = 0.0, com_max = -1.0, var_sum = 0.0, var_max = -1.0, cos_sum = 0.0, cos_min = 2.0;
  Welford wj{0.0, 0.0, 0}, wt{0.0, 0.0, 0};
  double Lfirst = 0.0;
  bool have_first = false, cos_nan = false, th_nan = false;
  int n_samp = 0, next_sample = 0;
  // MEGNO (evolution_features.py:34-66); the tangent map uses the post-step epsilon^2 (softening_manager.py:359-366)
  double megno = 2.0, lyap = inf, t_end = 0.0;
  double drx[N], dry[N], dvx[N], dvy[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { drx[i] = 0.0; dry[i] = 0.0; dvx[i] = 0.0; dvy[i] = 0.0; }
  double tt = 0.0, accum = 0.0;
  const double dt = a.dt;
  const int n_total = a.n_steps + a.n_megno;
  // ONE loop over the main steps and the MEGNO steps, so that the macro step (and the energy evaluation) is
  // instantiated once: step < n_steps samples step_metrics, step >= n_steps advances the tangent vectors
  for (int step = 0; step <= n_total; ++step) {
    if (want_energy && (step == 0 || step == a.n_steps)) {
      double E, Lz;
      energy(E, Lz);
      if (step == 0) { E0 = E; L0 = Lz; }
      if (step == a.n_steps) { E1 = E; L1 = Lz; }
    }
    if (step == a.n_steps && a.n_megno > 0) {
      double M = 0.0, cx = 0.0, cy = 0.0, ux = 0.0, uy = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        drx[i] = a.raw_dr[((size_t)sys * N + i) * 2 + 0]; dry[i] = a.raw_dr[((size_t)sys * N + i) * 2 + 1];
        dvx[i] = a.raw_dv[((size_t)sys * N + i) * 2 + 0]; dvy[i] = a.raw_dv[((size_t)sys * N + i) * 2 + 1];
        M += s.m[i];
        cx += s.m[i] * drx[i]; cy += s.m[i] * dry[i]; ux += s.m[i] * dvx[i]; uy += s.m[i] * dvy[i];
      }
      cx /= M; cy /= M; ux /= M; uy /= M;
      double nr = 0.0, nv = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        drx[i] -= cx; dry[i] -= cy; dvx[i] -= ux; dvy[i] -= uy;
        nr += drx[i] * drx[i] + dry[i] * dry[i]; nv += dvx[i] * dvx[i] + dvy[i] * dvy[i];
      }
      nr = sqrt(nr); nv = sqrt(nv);
#pragma unroll
      for (int i = 0; i < N; ++i) { drx[i] /= nr; dry[i] /= nr; dvx[i] /= nv; dvy[i] /= nv; }
    }
    if (step == n_total) break;
    macro_step();
    if (step < a.n_steps) {
      if (a.sample_interval > 0 && step == next_sample) {   // diagnostics.py:241-285
        next_sample += a.sample_interval;
        double cx = 0.0, cy = 0.0, Lt = 0.0, Li[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          cx += s.m[i] * s.x[i]; cy += s.m[i] * s.y[i];
          Li[i] = s.m[i] * (s.x[i] * s.vy[i] - s.y[i] * s.vx[i]);
          Lt += Li[i];
        }
        const double mean = Lt / N;
        double var = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) var += (Li[i] - mean) * (Li[i] - mean);
        var /= N;
        const double com = sqrt(cx * cx + cy * cy);
        if (!have_first) { Lfirst = Lt; have_first = true; }
        double c;
        if (Lfirst != 0.0 && Lt != 0.0) c = (Lt * Lfirst) / (fabs(Lt) * fabs(Lfirst));
        else { c = 0.0; cos_nan = true; }
        com_sum += com; com_max = fmax(com_max, com);
        var_sum += var; var_max = fmax(var_max, var);
        cos_sum += c; cos_min = fmin(cos_min, c);
        wj.add(s.eps * s.pi / P.mu);
        if (P.mu * s.eps != 0.0 || s.pi != 0.0) wt.add(atan2(s.pi, P.mu * s.eps));
        else th_nan = true;
        ++n_samp;
      }
    } else {
      SysState<N> t;
      double dax[N], day[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        drx[i] = fma(dvx[i], dt, drx[i]); dry[i] = fma(dvy[i], dt, dry[i]);
        t.gm[i] = G * s.m[i]; t.x[i] = s.x[i]; t.y[i] = s.y[i];
      }
      t.eps2 = s.eps * s.eps;
      pair_pass<N, true, true>(t, drx, dry, dax, day);
      double nr = 0.0, nv = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        dvx[i] = fma(dax[i], dt, dvx[i]); dvy[i] = fma(day[i], dt, dvy[i]);
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
  }
  if (a.n_megno > 0) {
    megno = 2.0 * accum / tt;
    lyap = (megno == 0.0) ? inf : tt / fabs(megno);
    t_end = tt;
  }

  if ((lane & (LPS - 1)) != 0 || !live) return;
  bool finite = is_finite(s.eps) && is_finite(s.pi);
#pragma unroll
  for (int i = 0; i < N; ++i)
    finite = finite && is_finite(s.x[i]) && is_finite(s.y[i]) && is_finite(s.vx[i]) && is_finite(s.vy[i]);
  int st = finite ? 0 : NB_STATUS_NONFINITE;
  {
    const double R = P.eps_max - P.eps_min;
    if (finite && (s.eps < P.eps_min - R || s.eps > P.eps_max + R)) st |= NB_STATUS_EPS_OOB;
  }
  if (a.status) a.status[sys] = st;
  if (a.flags & NB_RUN_WRITE_STATE) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      a.q[((size_t)sys * N + i) * 2 + 0] = s.x[i]; a.q[((size_t)sys * N + i) * 2 + 1] = s.y[i];
      a.v[((size_t)sys * N + i) * 2 + 0] = s.vx[i]; a.v[((size_t)sys * N + i) * 2 + 1] = s.vy[i];
    }
    a.eps_pi[2 * (size_t)sys] = s.eps;
    a.eps_pi[2 * (size_t)sys + 1] = s.pi;
  }
  if (a.dyn) {
    double* f = a.dyn + (size_t)sys * NB_N_DYN;
    auto drift_of = [&](double a0, double a1) {
      if (is_finite(a0) && fabs(a0) > 0.0 && is_finite(a1)) return fabs((a1 - a0) / a0);
      if (is_finite(a0) && is_finite(a1)) return fabs(a1 - a0);
      return inf;
    };
    const double ed = want_energy ? drift_of(E0, E1) : nan, ld = want_energy ? drift_of(L0, L1) : nan;
    const double inv = n_samp > 0 ? 1.0 / (double)n_samp : nan;
    const double com_mean = n_samp > 0 ? com_sum * inv : nan;
    f[NB_F_ENERGY_DRIFT] = ed; f[NB_F_ANGMOM_DRIFT] = ld;
    f[NB_F_COM_MEAN] = com_mean; f[NB_F_COM_MAX] = n_samp > 0 ? com_max : nan;
    f[NB_F_JEPS_MEAN] = n_samp > 0 ? wj.mean : nan;
    f[NB_F_JEPS_STD] = n_samp > 0 ? sqrt(wj.m2 / wj.n) : nan;
    f[NB_F_THETA_MEAN] = (n_samp > 0 && !th_nan) ? wt.mean : nan;
    f[NB_F_THETA_STD] = (n_samp > 0 && !th_nan) ? sqrt(wt.m2 / wt.n) : nan;
    f[NB_F_COS_MEAN] = (n_samp > 0 && !cos_nan) ? cos_sum * inv : nan;
    f[NB_F_COS_MIN] = (n_samp > 0 && !cos_nan) ? cos_min : nan;
    f[NB_F_VARL_MEAN] = n_samp > 0 ? var_sum * inv : nan; f[NB_F_VARL_MAX] = n_samp > 0 ? var_max : nan;
    f[NB_F_TIDAL_MEAN] = n_samp > 0 ? 0.0 : nan; f[NB_F_TIDAL_MAX] = n_samp > 0 ? 0.0 : nan;
    f[NB_F_MEGNO] = megno; f[NB_F_LYAP_TIME] = lyap;
    f[NB_F_IS_STABLE] = ((ed < 0.01) && (ld < 0.01) && (com_mean < 1.0) && (megno < 10.0)) ? 1.0 : 0.0;
    f[NB_F_E0] = E0; f[NB_F_E1] = E1; f[NB_F_L0] = L0; f[NB_F_L1] = L1; f[NB_F_T_END] = t_end;
  }
}

// ---------------------------------------------------------------------------------------------
// setup kernel (one thread per system): constructor calibration and the frozen sub-step schedule
//   flags bit0: calibrate (hamsoft_eps_model.py:645-729 + hamiltonian_softening_integrator.py:251-296)
//   flags bit1: freeze the production schedule for step size dt (:986-1221) -> n_sub
// ---------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(64) hamsoft_setup_kernel(const double* m_, const double* q_, double G, int B,
                                                           unsigned flags, double dt, double* hs, double* eps_pi,
                                                           int32_t* n_sub) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= B) return;
  double m[N], x[N], y[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    m[i] = m_[(size_t)sys * N + i];
    x[i] = q_[((size_t)sys * N + i) * 2 + 0];
    y[i] = q_[((size_t)sys * N + i) * 2 + 1];
  }
  double* hp = hs + (size_t)sys * NB_HS_NPARAM;
  HsPar P = hs_load(hp);
  double eps = eps_pi[2 * (size_t)sys];
  auto tau_grav = [&](double fallback) {
    double tau = __longlong_as_double(0x7ff0000000000000LL);
    if (G != 0.0) {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j) {
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
  if (flags & 1u) {
    double h[N];
    hs_solve_hi<N>(x, y, m, eps, P, h);
    // median of h
    double lo_v = 0.0, hi_v = 0.0;
    const int klo = (N - 1) / 2, khi = N / 2;
#pragma unroll
    for (int p = 0; p < N; ++p) {
      int less = 0, eq = 0;
#pragma unroll
      for (int r = 0; r < N; ++r) { less += h[r] < h[p]; eq += h[r] == h[p]; }
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
#pragma unroll
      for (int i = 0; i < N; ++i) M += m[i];
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
    const double h_o = om > 0.0 ? theta_cap / om : __longlong_as_double(0x7ff0000000000000LL);
    const double h_theta = (is_finite(h_o) && h_o > 0.0) ? fmin(h_g, h_o) : h_g;
    // pi budget (hamiltonian_softening_integrator.py:1125-1221)
    double h_pi = dt_abs;
    if (is_finite(P.k) && P.k > 0.0) {
      const double es = hs_eps_target<N>(x, y, m, eps, P);
      const double s0 = (is_finite(P.s0) && P.s0 > 0.0) ? P.s0 : 1.0;
      const double d_eff = fmax(fabs(eps - (is_finite(es) ? es : P.s0)), 1.0e-4 * s0);
      double gmv[N], U, S3;
#pragma unroll
      for (int i = 0; i < N; ++i) gmv[i] = G * m[i];
      pair_scalars<N, true>(gmv, m, x, y, eps * eps, U, S3);
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

// eps*(q), its gradient and H_ext for B systems (parity taps; one warp per system)
template <int N>
__global__ void __launch_bounds__(128) hamsoft_probe_kernel(const double* m_, const double* q_, const double* v_,
                                                            double G, int B, const double* eps_pi, const double* hs,
                                                            double* out /*[B][2+2N]: eps*, H, grad*/) {
  const int lane = threadIdx.x & 31;
  const int sys = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (sys >= B) return;
  double m[N], x[N], y[N], vx[N], vy[N], gx[N], gy[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    m[i] = m_[(size_t)sys * N + i];
    x[i] = q_[((size_t)sys * N + i) * 2 + 0]; y[i] = q_[((size_t)sys * N + i) * 2 + 1];
    vx[i] = v_[((size_t)sys * N + i) * 2 + 0]; vy[i] = v_[((size_t)sys * N + i) * 2 + 1];
  }
  const HsPar P = hs_load(hs + (size_t)sys * NB_HS_NPARAM);
  const double eps = eps_pi[2 * (size_t)sys], pi = eps_pi[2 * (size_t)sys + 1];
  bool fb;
  const double es = hs_eps_star_and_grad<N>(x, y, m, eps, P, lane, gx, gy, fb);
  const double H = hs_energy<N>(m, x, y, vx, vy, eps, pi, es, P, G);
  if (lane == 0) {
    double* o = out + (size_t)sys * (3 + 2 * N);
    o[0] = es; o[1] = H; o[2] = fb ? 1.0 : 0.0;
    for (int i = 0; i < N; ++i) { o[3 + 2 * i] = gx[i]; o[4 + 2 * i] = gy[i]; }
  }
}

#define NB_HS_DISPATCH(N_, CALL)              \
  switch (N_) {                               \
    case 2: { constexpr int NN = 2; CALL; } break; \
    case 3: { constexpr int NN = 3; CALL; } break; \
    case 4: { constexpr int NN = 4; CALL; } break; \
    case 5: { constexpr int NN = 5; CALL; } break; \
    case 6: { constexpr int NN = 6; CALL; } break; \
    case 7: { constexpr int NN = 7; CALL; } break; \
    case 8: { constexpr int NN = 8; CALL; } break; \
    default: set_error("N must be in 2..8"); return NB_ERR_ARG; \
  }

int hamsoft_run(const double* m, double* q, double* v, double G, int B, int N, unsigned flags, double dt, int n_steps,
                int sample_interval, int n_megno, const int32_t* n_sub, const int32_t* perm, const double* raw_dr,
                const double* raw_dv, double* eps_pi, const double* hs, double* dyn, int32_t* status, cudaStream_t st) {
  (void)perm;
  HsArgs a{m, q, v, G, B, flags, dt, n_steps, sample_interval, n_megno, n_sub, raw_dr, raw_dv, eps_pi, hs, dyn, status};
  const int blocks = (B + 3) / 4;
  NB_HS_DISPATCH(N, (hamsoft_run_kernel<NN><<<hs_run_blocks<NN>(a.B), 128, 0, st>>>(a)));
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int hamsoft_setup(const double* m, const double* q, double G, int B, int N, unsigned flags, double dt, double* hs,
                  double* eps_pi, int32_t* n_sub, cudaStream_t st) {
  const int blocks = (B + 63) / 64;
  NB_HS_DISPATCH(N, (hamsoft_setup_kernel<NN><<<blocks, 64, 0, st>>>(m, q, G, B, flags, dt, hs, eps_pi, n_sub)));
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int hamsoft_probe(const double* m, const double* q, const double* v, double G, int B, int N, const double* eps_pi,
                  const double* hs, double* out, cudaStream_t st) {
  const int blocks = (B + 3) / 4;
  NB_HS_DISPATCH(N, (hamsoft_probe_kernel<NN><<<blocks, 128, 0, st>>>(m, q, v, G, B, eps_pi, hs, out)));
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
