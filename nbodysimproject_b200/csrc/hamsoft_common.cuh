// hamsoft_common.cuh -- parameter block, barrier, fold, spring-rotation constants and the small numeric helpers shared by
// the ham_soft kernels (hamsoft.cu: N <= 8, one group of lanes per system; hamsoft_mid.cu: 9..64 bodies, one CTA per system).
#pragma once
#include "pair_small.cuh"
#include "args.cuh"

namespace nb {

#define NB_PI 3.14159265358979323846

struct HsPar {
  double k, mu, eps_min, eps_max, alpha, k_wall, eta, jcap, lam, theta_imp, theta_cap, chi_pi, omega0, s0;
  int n_exp, policy, flags;
};

__device__ __forceinline__ HsPar hs_load(const double* p) {
  HsPar h;
  h.k = p[NB_HS_K_SOFT]; h.mu = p[NB_HS_MU_SOFT]; h.eps_min = p[NB_HS_EPS_MIN]; h.eps_max = p[NB_HS_EPS_MAX];
  h.alpha = p[NB_HS_ALPHA_RUN]; h.k_wall = p[NB_HS_K_WALL]; h.n_exp = (int)p[NB_HS_BARRIER_N]; h.eta = p[NB_HS_ETA];
  h.jcap = p[NB_HS_J_MAX_CAP]; h.lam = p[NB_HS_LAMBDA]; h.policy = (int)p[NB_HS_POLICY];
  h.theta_imp = p[NB_HS_THETA_IMP]; h.theta_cap = p[NB_HS_THETA_CAP]; h.chi_pi = p[NB_HS_CHI_PI];
  h.omega0 = p[NB_HS_OMEGA_SPR0]; h.s0 = p[NB_HS_S0]; h.flags = (int)p[NB_HS_FLAGS];
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
  if (eps >= P.eps_min && eps <= P.eps_max) return 0.0;          // inside the admissible interval: both terms vanish
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

#define HS_NACC 22
#define NB_INV_PI 0.31830988618379067154

// exp / log out of line: inlined at every call site (N (N-1) in the Jacobi sweep alone) they made the straight-line code
// of a sub-step ~80 KB per warp and the kernel stalled on instruction fetch (ncu r2: no_instruction 2.8 warps per issue)
static __device__ __noinline__ double hs_exp(double x) { return exp(x); }
static __device__ __noinline__ double hs_log(double x) { return log(x); }

// a / b for a divisor that is finite, positive and far from the denormal range (smoothing lengths, masses, alpha ...):
// MUFU.RCP64H seed, two Newton steps, one residual correction = 8 FP64-pipe instructions and no slow-path call.
// The compiler's generic division is ~15 instructions PLUS a ~60-instruction subroutine whenever the dividend is zero
// or tiny -- which is the common case here (a converged Jacobi sweep divides |h_new - h| = 0, a clamped eps* has a zero
// central difference): ncu attributed 15-22 % of all executed instructions to that subroutine.  Faithfully rounded
// (<= 1 ulp), like the compiler's fast path.
__device__ __forceinline__ double hs_div(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
__device__ __forceinline__ double hs_rcp(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  return fma(r, e, r);
}

// loop invariants of the spring rotation (hamsoft_flows.py:427-762): k, mu and the sub-step h are frozen for a launch
struct HsSpring {
  double om, sn, cs, mo, inv_mu_om, inv_om, inv_den;
  int rot;      // om != 0 && mu != 0
  int den_ok;   // mu om^2 != 0
};
enum { HA_COM_SUM = 0, HA_COM_MAX, HA_VAR_SUM, HA_VAR_MAX, HA_COS_SUM, HA_COS_MIN, HA_WJ_MEAN, HA_WJ_M2, HA_WT_MEAN,
       HA_WT_M2, HA_LFIRST, HA_NSAMP, HA_WJ_N, HA_WT_N, HA_HAVE_FIRST, HA_COS_NAN, HA_TH_NAN, HA_E0, HA_L0, HA_E1, HA_L1 };

// FD step of coordinate value x (hamsoft_eps_model.py:136-144)
__device__ __forceinline__ double hs_fd_step(double x) {
  double h = 1.0e-5 * fmax(fabs(x), 1.0);
  if (h < 1.0e-10) h = 1.0e-10;
  return h;
}

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

// rotation constants of the S half-flow for sub-step h (theta = omega h / 2; small-angle series below 1e-8)
__device__ __forceinline__ void hs_spring_setup(HsSpring& R, const HsPar& P, double h) {
  const double k = P.k, mu = P.mu;
  const double om = (k > 0.0 && mu > 0.0) ? sqrt(k / mu) : 0.0;
  const double th = om * (0.5 * h);
  if (fabs(th) < 1.0e-8) {
    const double t2 = th * th;
    R.sn = th - t2 * th / 6.0 + t2 * t2 * th / 120.0;
    R.cs = 1.0 - t2 / 2.0 + t2 * t2 / 24.0;
  } else {
    sincos(th, &R.sn, &R.cs);
  }
  R.om = om;
  R.rot = (om != 0.0 && mu != 0.0) ? 1 : 0;
  R.mo = sqrt(mu * fmax(k, 0.0));
  const double den = mu * om * om;
  R.den_ok = (den != 0.0) ? 1 : 0;
  R.inv_mu_om = R.rot ? 1.0 / (mu * om) : 0.0;
  R.inv_om = R.rot ? 1.0 / om : 0.0;
  R.inv_den = R.den_ok ? 1.0 / den : 0.0;
}

// launch arguments of the run kernels (hamsoft.cu: N <= 8; hamsoft_mid.cu: 9..64 bodies)
struct HsArgs {
  const double* m; double* q; double* v; double G; int B; unsigned flags; double dt; int n_steps; int sample_interval;
  int n_megno; const int32_t* n_sub; const int32_t* perm; const double* raw_dr; const double* raw_dv; double* eps_pi;
  const double* hs; double* dyn; int32_t* status; double* work; unsigned long long* tstamp;
};

// one step_metrics sample (diagnostics.py:241-285) from the scalars of the system: |COM|, var(L_i), L_total, and the
// Welford accumulators of J_eps = eps pi / mu and theta = atan2(pi, mu eps)
__device__ __forceinline__ void hs_sample_scalars(double* A, double com, double var, double Lt, double eps, double pi,
                                                  double mu) {
  if (A[HA_HAVE_FIRST] == 0.0) { A[HA_LFIRST] = Lt; A[HA_HAVE_FIRST] = 1.0; }
  const double Lfirst = A[HA_LFIRST];
  double c;
  if (Lfirst != 0.0 && Lt != 0.0) c = (Lt * Lfirst) / (fabs(Lt) * fabs(Lfirst));
  else { c = 0.0; A[HA_COS_NAN] = 1.0; }
  A[HA_COM_SUM] += com; A[HA_COM_MAX] = fmax(A[HA_COM_MAX], com);
  A[HA_VAR_SUM] += var; A[HA_VAR_MAX] = fmax(A[HA_VAR_MAX], var);
  A[HA_COS_SUM] += c; A[HA_COS_MIN] = fmin(A[HA_COS_MIN], c);
  {
    const double xj = eps * pi / mu;
    const double n = (A[HA_WJ_N] += 1.0);
    const double d = xj - A[HA_WJ_MEAN];
    A[HA_WJ_MEAN] += d / n;
    A[HA_WJ_M2] += d * (xj - A[HA_WJ_MEAN]);
  }
  if (mu * eps != 0.0 || pi != 0.0) {
    const double xt = atan2(pi, mu * eps);
    const double n = (A[HA_WT_N] += 1.0);
    const double d = xt - A[HA_WT_MEAN];
    A[HA_WT_MEAN] += d / n;
    A[HA_WT_M2] += d * (xt - A[HA_WT_MEAN]);
  } else {
    A[HA_TH_NAN] = 1.0;
  }
  A[HA_NSAMP] += 1.0;
}

// the dynamic feature row of one system from its accumulators (stability_analyzer.py:69-259)
__device__ __forceinline__ void hs_write_dyn(double* f, const double* A, bool want_energy, double megno, double lyap,
                                             double t_end) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  auto drift_of = [&](double a0, double a1) {
    if (is_finite(a0) && fabs(a0) > 0.0 && is_finite(a1)) return fabs((a1 - a0) / a0);
    if (is_finite(a0) && is_finite(a1)) return fabs(a1 - a0);
    return inf;
  };
  const int n_samp = (int)A[HA_NSAMP];
  const double E0 = A[HA_E0], L0 = A[HA_L0], E1 = A[HA_E1], L1 = A[HA_L1];
  const bool th_nan = A[HA_TH_NAN] != 0.0, cos_nan = A[HA_COS_NAN] != 0.0;
  const double ed = want_energy ? drift_of(E0, E1) : nan, ld = want_energy ? drift_of(L0, L1) : nan;
  const double inv = n_samp > 0 ? 1.0 / (double)n_samp : nan;
  const double com_mean = n_samp > 0 ? A[HA_COM_SUM] * inv : nan;
  f[NB_F_ENERGY_DRIFT] = ed; f[NB_F_ANGMOM_DRIFT] = ld;
  f[NB_F_COM_MEAN] = com_mean; f[NB_F_COM_MAX] = n_samp > 0 ? A[HA_COM_MAX] : nan;
  f[NB_F_JEPS_MEAN] = n_samp > 0 ? A[HA_WJ_MEAN] : nan;
  f[NB_F_JEPS_STD] = n_samp > 0 ? sqrt(A[HA_WJ_M2] / A[HA_WJ_N]) : nan;
  f[NB_F_THETA_MEAN] = (n_samp > 0 && !th_nan) ? A[HA_WT_MEAN] : nan;
  f[NB_F_THETA_STD] = (n_samp > 0 && !th_nan) ? sqrt(A[HA_WT_M2] / A[HA_WT_N]) : nan;
  f[NB_F_COS_MEAN] = (n_samp > 0 && !cos_nan) ? A[HA_COS_SUM] * inv : nan;
  f[NB_F_COS_MIN] = (n_samp > 0 && !cos_nan) ? A[HA_COS_MIN] : nan;
  f[NB_F_VARL_MEAN] = n_samp > 0 ? A[HA_VAR_SUM] * inv : nan; f[NB_F_VARL_MAX] = n_samp > 0 ? A[HA_VAR_MAX] : nan;
  f[NB_F_TIDAL_MEAN] = n_samp > 0 ? 0.0 : nan; f[NB_F_TIDAL_MAX] = n_samp > 0 ? 0.0 : nan;
  f[NB_F_MEGNO] = megno; f[NB_F_LYAP_TIME] = lyap;
  f[NB_F_IS_STABLE] = ((ed < 0.01) && (ld < 0.01) && (com_mean < 1.0) && (megno < 10.0)) ? 1.0 : 0.0;
  f[NB_F_E0] = E0; f[NB_F_E1] = E1; f[NB_F_L0] = L0; f[NB_F_L1] = L1; f[NB_F_T_END] = t_end;
}

// 9..64 bodies, one CTA per system (hamsoft_mid.cu)
int hamsoft_mid_run(const HsArgs& a, int N, cudaStream_t st);
int hamsoft_mid_setup(const double* m, const double* q, double G, int B, int N, unsigned flags, double dt, double* hs,
                      double* eps_pi, int32_t* n_sub, cudaStream_t st);
int hamsoft_mid_probe(const double* m, const double* q, const double* v, double G, int B, int N, const double* eps_pi,
                      const double* hs, double* out, cudaStream_t st);

}  // namespace nb
