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
#include "hamsoft_common.cuh"

namespace nb {

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

// =================================================================================================
// Group-cooperative run path.  A "group" = the LPS (16 or 32) lanes that serve one system.  The system's state lives
// ONCE in shared memory (HsSh); a lane keeps in registers only what its current phase needs:
//   * eps* model  : lane e holds ONE finite-difference evaluation (pair distances r2[NP] + smoothing lengths h[N]),
//                   nothing else is live across the Jacobi sweeps -> no local-memory spill
//   * every other flow (gradient assembly, J cap, V half kick, T drift, tangent map): lane i < N owns BODY i,
//                   sums / maxima over bodies are xor-butterflies inside the group
// Round 1 kept a full replica of the state, the parameters, the sampling accumulators and the tangent vectors in the
// registers of every lane (61 live doubles at N = 3 against a 128-register cap): 0.8-1.7 KB of spill per thread,
// 410 MB of DRAM writes per launch, FP64 pipe 34 % busy.
// Two systems may share a warp (LPS = 16): every shuffle names its source lane inside the group and the butterflies use
// offsets < LPS, so full-mask intrinsics stay legal; a system that has fewer sub-steps than its neighbour executes the
// excess with its writes predicated off (`act`).
// =================================================================================================
template <int N>
struct HsSh {
  static constexpr int NP = N * (N - 1) / 2;
  double m[N], x[N], y[N], vx[N], vy[N];
  double gx[N], gy[N];          // grad eps* of the current S half-flow
  double h[N];                  // smoothing lengths of the UNPERTURBED configuration (reused by the analytic fallback)
  double bx[N], by[N];          // per-body exchange buffer of the analytic fallback
  double drx[N], dry[N], dvx[N], dvy[N];   // tangent vectors (MEGNO)
  double rs[NP > 0 ? NP : 1];   // pair separations (median test, legacy gradient)
  double acc[HS_NACC];          // step_metrics accumulators, touched by the group's lane 0 only
  double inv_alpha;             // 1 / alpha_run
  HsSpring spr;
  HsPar P;
};

template <int LPS>
__device__ __forceinline__ double grp_max(double v) {
#pragma unroll
  for (int off = LPS / 2; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
template <int LPS>
__device__ __forceinline__ double grp_sum(double v) {
#pragma unroll
  for (int off = LPS / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
// index of the unordered pair (a < b) in the row-major upper triangle
template <int N>
__device__ __forceinline__ int hs_pair_index(int a, int b) { return a * N - a * (a + 1) / 2 + (b - a - 1); }

// Jacobi sweeps of hamsoft_eps_model.py:316-400 on register-resident pair distances; returns the sweeps executed.
// exp(-r^2/h^2) underflows to exactly 0 below -745.13: those pairs (most of them once bodies are a few h apart) skip
// the exponential altogether -- same bits, and when a whole warp agrees, none of its instructions.
template <int N>
__device__ __forceinline__ int hs_solve_regs(const double (&r2)[N * (N - 1) / 2 > 0 ? N * (N - 1) / 2 : 1],
                                             const double* m, double eps_cur, const HsPar& P, double (&h)[N]) {
  double lo = P.eps_min, hi = P.eps_max;
  if (hi < lo) { const double t = lo; lo = hi; hi = t; }
  const double flo = fmax(lo, 1.0e-12), cap = fmax(flo, hi);
  double h0 = eps_cur;
  if (!is_finite(h0) || h0 <= 0.0) h0 = 1.0;
  h0 = fmin(fmax(h0, flo), cap);
#pragma unroll
  for (int i = 0; i < N; ++i) h[i] = h0;
  const double eta = P.eta;
  int it = 0;
#pragma unroll 1
  while (it < 8) {
    bool conv = true;
    double hn[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double hj = fmax(h[i], 1.0e-12);
      const double inv = hs_rcp(hj * hj);         // one reciprocal per body and sweep: c = 1/(pi h^2), -1/h^2
      const double c = inv * NB_INV_PI;
      const double nih2 = -inv;
      double S = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        const int a = i < j ? i : j, b = i < j ? j : i;
        const double arg = r2[a * N - a * (a + 1) / 2 + (b - a - 1)] * nih2;
        if (arg > -746.0) S += m[j] * (c * hs_exp(arg));
      }
      const double Si = fmax(S, 1.0e-30);
      double v = eta * sqrt(hs_div(m[i], Si));
      if (!is_finite(v) || v <= 0.0) v = h[i];
      if (v < flo) v = flo;
      else if (v > cap) v = cap;
      // max_i |v - h_i| / max(h_i, 1e-12) < 1e-6, without the division
      conv = conv && (fabs(v - h[i]) < 1.0e-6 * hj);
      hn[i] = v;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) h[i] = hn[i];
    ++it;
    if (conv) break;
  }
  return it;
}

// hamsoft_eps_model.py:240-289: soft-min of the h_i with temperature alpha_run (+ clamp under the soft policy)
template <int N>
__device__ __forceinline__ double hs_softmin(const double (&h)[N], const HsPar& P, double inv_alpha) {
  const double alpha = P.alpha;
  double t[N];
  double tmax = -h[0] * inv_alpha;
  t[0] = tmax;
#pragma unroll
  for (int i = 1; i < N; ++i) { t[i] = -h[i] * inv_alpha; tmax = fmax(tmax, t[i]); }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double d = t[i] - tmax;                // <= 0; the largest term is exp(0) = 1 exactly
    s += (d == 0.0) ? 1.0 : ((d > -746.0) ? hs_exp(d) : 0.0);
  }
  double es;
  if (s <= 0.0 || !is_finite(s)) es = P.s0;
  else es = -alpha * (tmax + hs_log(s));
  if (P.policy == 0) {
    double lo = P.eps_min, hi = P.eps_max;
    if (hi < lo) { const double t2 = lo; lo = hi; hi = t2; }
    if (es < lo) es = lo;
    else if (es > hi) es = hi;
  }
  return es;
}

// Lanes per system.  4 N + 1 evaluations fit a half warp for N <= 3, so two systems share a warp there.  For N = 4
// (17) and N = 8 (33) the odd one out -- the UNPERTURBED evaluation -- is computed cooperatively instead (one body per
// lane, Jacobi sweeps exchanged by shuffles; same arithmetic per body, so the same bits as the serial solve at 1/N of
// its cost): the 4 N perturbed evaluations then fit a half warp (N = 4: two systems per warp) or one pass (N = 8).
template <int N>
struct HsLanes {
  static constexpr bool COOP = (N == 4 || N == 8);
  static constexpr int NE = COOP ? 4 * N : 4 * N + 1;        // evaluations, one per lane
  static constexpr int LPS = (NE <= 16) ? 16 : 32;
  static_assert(NE <= LPS, "one finite-difference evaluation per lane");
};

// eps_target of the unperturbed configuration, cooperatively over the lanes of one group (hamsoft_eps_model.py:316-400
// + :240-289); lane k < N ends up with h_k and publishes it in sh.h.  Warp-uniform control flow: the sweep loop always
// runs 8 times and a converged group simply stops updating (the two systems of a warp may converge at different sweeps
// and the shuffles need the whole warp).
template <int N>
__device__ __forceinline__ double hs_eps_target_coop(HsSh<N>& sh, double eps_cur, int lane, int base, int& sweeps) {
  constexpr int LPS = HsLanes<N>::LPS;
  const HsPar& P = sh.P;
  const int i = lane < N ? lane : 0;                        // spare lanes shadow body 0
  const double xi = sh.x[i], yi = sh.y[i], mi = sh.m[i];
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
    const double xj = sh.x[j], yj = sh.y[j];
    const double dx = (i < j) ? xi - xj : xj - xi, dy = (i < j) ? yi - yj : yj - yi;
    r2[j] = dx * dx + dy * dy;
  }
  double hcur = h0;
  bool done = false;
  int used = 0;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const double hj = fmax(hcur, 1.0e-12);
    const double inv = hs_rcp(hj * hj);
    const double c = inv * NB_INV_PI;
    const double nih2 = -inv;
    double S = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double arg = r2[j] * nih2;
      if (j != i && arg > -746.0) S += sh.m[j] * (c * hs_exp(arg));
    }
    const double Si = fmax(S, 1.0e-30);
    double v = P.eta * sqrt(hs_div(mi, Si));
    if (!is_finite(v) || v <= 0.0) v = hcur;
    if (v < flo) v = flo;
    else if (v > cap) v = cap;
    const bool conv_i = (lane >= N) || (fabs(v - hcur) < 1.0e-6 * hj);
    const bool conv = grp_max<LPS>(conv_i ? 0.0 : 1.0) == 0.0;
    if (!done) { hcur = v; ++used; }
    if (conv) done = true;                                  // group-uniform
  }
  if (lane == 0) sweeps += used;
  if (lane < N) sh.h[lane] = hcur;
  double h[N];
#pragma unroll
  for (int k = 0; k < N; ++k) h[k] = __shfl_sync(0xffffffffu, hcur, base + k);
  return hs_softmin<N>(h, P, sh.inv_alpha);
}

// hamsoft_eps_model.py:451-556 analytic SPH gradient + softening.py:86-131 sign reference, one body per lane.
// Runs when ANY group of the warp needs it (warp-uniform branch, so the butterflies inside stay legal); a group
// commits the result only if it needs it itself.  sh.h holds the smoothing lengths of the unperturbed configuration
// (the reference calls _solve_hi again with identical inputs).
template <int N>
__device__ __forceinline__ void hs_fallback_grp(HsSh<N>& sh, int lane, bool commit) {
  constexpr int LPS = HsLanes<N>::LPS;
  constexpr int NP = HsSh<N>::NP;
  const HsPar& P = sh.P;
  const bool mine = lane < N;
  const int i = mine ? lane : 0;
  const double xi = sh.x[i], yi = sh.y[i], mi = sh.m[i];
  const double flo = fmax(P.eps_min, 1.0e-12);
  const double hmin = fmax(1.0e-12, 0.1 * flo);
  const double ti = -sh.h[i] * sh.inv_alpha;
  const double tmax = grp_max<LPS>(ti);
  const double di = ti - tmax;
  const double ei = (di == 0.0) ? 1.0 : ((di > -746.0) ? hs_exp(di) : 0.0);
  __syncwarp();
  if (mine) sh.bx[i] = ei;
  __syncwarp();
  double den = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) den += sh.bx[k];
  const bool bad = (den <= 0.0 || !is_finite(den));
  const double wi = bad ? 0.0 : hs_div(ei, den);
  const double hj = fmax(sh.h[i], hmin);
  const double ihj = hs_rcp(hj), ih2 = ihj * ihj;           // 1/h, 1/h^2: every quotient of this body reuses them
  const double c = ih2 * NB_INV_PI;
  double S = 0.0, Sd = 0.0;
  double W[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
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
  __syncwarp();
  if (mine) { sh.bx[i] = si; sh.by[i] = ih2; }
  __syncwarp();
  // the reference scatters s_i m_j coef_i(r_ij) (q_i - q_j) onto i (+) and j (-); gathered per body here
  double gx = 0.0, gy = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    if (j == i || W[j] == 0.0) continue;
    const double rx = xi - sh.x[j], ry = yi - sh.y[j];
    const double coef = -2.0 * W[j] * ih2;
    gx += si * sh.m[j] * (coef * rx);
    gy += si * sh.m[j] * (coef * ry);
  }
#pragma unroll
  for (int a = 0; a < N; ++a) {
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
  // sign alignment against the legacy gradient: only sum(g_use . g_legacy) matters
  // (only the SIGN of sum(g_use . g_legacy) is used, and only when the analytic gradient is non-zero)
  const bool any_g = grp_max<LPS>((mine && (gx != 0.0 || gy != 0.0)) ? 1.0 : 0.0) != 0.0;
  double sg = 1.0;
  {
    const double dp = lane < NP ? hs_rcp(fmax(sh.rs[lane < NP ? lane : 0], 1.0e-15) + 1.0e-12) : 0.0;
    const double D = grp_sum<LPS>(dp);
    const double cp = P.lam * ((double)N / (D * D));
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const int a = i < j ? i : j, b = i < j ? j : i;
      const double r = fmax(sh.rs[hs_pair_index<N>(a, b)], 1.0e-15);
      const double dn = r + 1.0e-12;
      const double A = hs_rcp(r * dn * dn);
      sx += A * (xi - sh.x[j]);
      sy += A * (yi - sh.y[j]);
    }
    const double lx = -cp * sx, ly = -cp * sy;
    const double nok = grp_max<LPS>((mine && !(is_finite(lx) && is_finite(ly))) ? 1.0 : 0.0);
    const double dot = grp_sum<LPS>(mine ? gx * lx + gy * ly : 0.0);
    if (any_g && is_finite(D) && D > 0.0 && nok == 0.0 && is_finite(dot) && dot < 0.0) sg = -1.0;
  }
  if (commit && mine) { sh.gx[i] = sg * gx; sh.gy[i] = sg * gy; }
}

// eps*(q) and its gradient (hamsoft_eps_model.py:94-234) for the system in `sh`: lane e of the group evaluates ONE
// configuration (e = 0: unperturbed unless COOP; then coordinate c = 2 i + a perturbed by +h / -h), body lanes assemble
// the central differences, and the gradient lands in sh.gx / sh.gy.  Returns eps* on every lane of the group.
template <int N>
__device__ __forceinline__ double hs_eps_star_and_grad(HsSh<N>& sh, double eps_cur, int lane_full, bool& used_fallback,
                                                       int& sweeps) {
  constexpr bool COOP = HsLanes<N>::COOP;
  constexpr int NE = HsLanes<N>::NE;
  constexpr int LPS = HsLanes<N>::LPS;
  constexpr int OFF = COOP ? 0 : 1;           // evaluation index of the first perturbed configuration
  constexpr int NP = HsSh<N>::NP;
  const int lane = lane_full & (LPS - 1);     // lane within this system's group
  const int base = lane_full - lane;          // first lane of the group
  const HsPar& P = sh.P;
  __syncwarp();                               // positions written by the body lanes are visible to every lane
  double f;
  {
    const int ee = lane < NE ? lane : 0;      // idle lanes redo evaluation 0
    const bool pert = COOP || ee >= 1;        // without COOP evaluation 0 is the unperturbed configuration
    const int c = (ee - OFF) >> 1;            // perturbed coordinate
    const double sgn = ((ee - OFF) & 1) ? -1.0 : 1.0;
    double r2[NP > 0 ? NP : 1];
    {
      double px[N], py[N];
      // this lane's perturbed coordinate (body c >> 1, axis c & 1): one finite-difference step, then a select per body
      const int pb = pert ? (c >> 1) : 0;
      const bool on_y = (c & 1) != 0;
      const double base_v = on_y ? sh.y[pb] : sh.x[pb];
      const double pert_v = base_v + sgn * hs_fd_step(base_v);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        px[i] = sh.x[i];
        py[i] = sh.y[i];
        if (pert && i == pb) { if (on_y) py[i] = pert_v; else px[i] = pert_v; }
      }
      int p = 0;
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j) {
          const double dx = px[i] - px[j], dy = py[i] - py[j];
          r2[p++] = dx * dx + dy * dy;
        }
    }
    double h[N];
    const int used = hs_solve_regs<N>(r2, sh.m, eps_cur, P, h);
    if (lane < NE) sweeps += used;
    if (!COOP && lane == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) sh.h[i] = h[i];
    }
    f = hs_softmin<N>(h, P, sh.inv_alpha);
  }
  double es;
  if constexpr (COOP) es = hs_eps_target_coop<N>(sh, eps_cur, lane, base, sweeps);
  else es = __shfl_sync(0xffffffffu, f, base);
  // central differences, one body per lane
  const bool mine = lane < N;
  const int i = mine ? lane : N - 1;
  const int src = base + OFF + 4 * i;
  const double fpx = __shfl_sync(0xffffffffu, f, src), fmx = __shfl_sync(0xffffffffu, f, src + 1);
  const double fpy = __shfl_sync(0xffffffffu, f, src + 2), fmy = __shfl_sync(0xffffffffu, f, src + 3);
  const double xi = sh.x[i], yi = sh.y[i];
  // a clamped eps* has an exactly zero central difference (the common case under the soft policy)
  const double dfx = fpx - fmx, dfy = fpy - fmy;
  double gx = (dfx == 0.0) ? 0.0 : dfx / (2.0 * hs_fd_step(xi));
  double gy = (dfy == 0.0) ? 0.0 : dfy / (2.0 * hs_fd_step(yi));
  if (!is_finite(gx)) gx = 0.0;
  if (!is_finite(gy)) gy = 0.0;
  const double gmax = sqrt(grp_max<LPS>(mine ? gx * gx + gy * gy : 0.0));
  if (mine) { sh.gx[i] = gx; sh.gy[i] = gy; }
  // median pair separation: one pair per lane, rank selection through shared memory
  double rsp = 0.0;
  if (lane < NP) {
    int pa = 0, pb = 1, idx = 0;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
      for (int b = a + 1; b < N; ++b) { if (idx == lane) { pa = a; pb = b; } ++idx; }
    const double dx = sh.x[pa] - sh.x[pb], dy = sh.y[pa] - sh.y[pb];
    rsp = sqrt(dx * dx + dy * dy);
    sh.rs[lane] = rsp;
  }
  __syncwarp();
  const double ninf = __longlong_as_double(0xfff0000000000000LL);
  double lo_c = ninf, hi_c = ninf;
  if (lane < NP) {
    constexpr int klo = (NP - 1) / 2, khi = NP / 2;
    int less = 0, eq = 0;
#pragma unroll
    for (int r = 0; r < NP; ++r) { const double v = sh.rs[r]; less += v < rsp; eq += v == rsp; }
    if (less <= klo && klo < less + eq) lo_c = rsp;
    if (less <= khi && khi < less + eq) hi_c = rsp;
  }
  const double lo_v = grp_max<LPS>(lo_c), hi_v = grp_max<LPS>(hi_c);
  const double rmed = (lo_v > ninf && hi_v > ninf) ? 0.5 * (lo_v + hi_v) : 0.0;
  used_fallback = (gmax <= 1.0e-12) || (gmax <= 1.0e-9 * rmed);
  if (__any_sync(0xffffffffu, used_fallback)) hs_fallback_grp<N>(sh, lane, used_fallback);
  __syncwarp();
  return es;
}

// S half-flow: hamsoft_stepper.py:47-88 -> spring_oscillation (live definition) hamsoft_flows.py:427-762.
// eps, pi are replicated scalars of the group; momenta are updated by the body lanes.
template <int N>
__device__ __forceinline__ void hs_s_half(HsSh<N>& sh, double& eps, double& pi, double h, int lane_full, bool act,
                                          int& sweeps) {
  constexpr int LPS = HsLanes<N>::LPS;
  const int lane = lane_full & (LPS - 1);
  const HsPar& P = sh.P;
  const double dt = 0.5 * h;
  double eps0 = eps, pi0 = pi;
  hs_fold(eps0, pi0, P);                         // hamsoft_stepper.py:107-113
  bool fb;
  const double es = hs_eps_star_and_grad<N>(sh, eps0, lane_full, fb, sweeps);
  const HsSpring& R = sh.spr;                    // k, mu, h are frozen for the launch: rotation constants precomputed
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
  const bool mine = lane < N;
  const int i = mine ? lane : 0;
  const double mi = sh.m[i], vx = sh.vx[i], vy = sh.vy[i], gx = sh.gx[i], gy = sh.gy[i];
  const double px = mi * vx, py = mi * vy;
  const double jx = J * gx, jy = J * gy;
  const double pmax2 = grp_max<LPS>(mine ? px * px + py * py : 0.0);
  const double dmax2 = grp_max<LPS>(mine ? jx * jx + jy * jy : 0.0);
  const double p_scale = fmax(sqrt(pmax2), 1.0e-12);
  const double dp_inf = sqrt(dmax2);
  const double thr = P.jcap * p_scale;
  const double Ja = (dp_inf > thr && dp_inf > 0.0) ? J * (thr / dp_inf) : J;
  // SimConfig.freeze_s_subsystem (hamsoft_stepper.py:119-124): the half-flow ends after the fold; predicated, not
  // branched, because two systems may share the warp and the code above shuffles
  const bool frozen = (P.flags & NB_HS_FLAG_FREEZE_S) != 0;
  if (act && mine && Ja != 0.0 && !frozen) {    // p += J grad eps*; v = p / m
    const double im = hs_rcp(mi);
    sh.vx[i] = (mi * vx + Ja * gx) * im;
    sh.vy[i] = (mi * vy + Ja * gy) * im;
  }
  double pi_out = eta_t + kick2;
  hs_fold(eps_rot, pi_out, P);                   // hamsoft_stepper.py:72-80
  if (act) { eps = frozen ? eps0 : eps_rot; pi = frozen ? pi0 : pi_out; }
}

// V half-kick: hamsoft_stepper.py:543-663 + pi_half_kick hamsoft_flows.py:1102-1132; body i on lane i.  The pair factor
// G m_lo m_hi rho^-3 is formed identically on both lanes of a pair, so Newton's third law holds to the bit.
template <int N>
__device__ __forceinline__ void hs_v_half(HsSh<N>& sh, double eps, double& pi, double G, double h, int lane_full,
                                          bool act) {
  constexpr int LPS = HsLanes<N>::LPS;
  const int lane = lane_full & (LPS - 1);
  const HsPar& P = sh.P;
  const double hh = 0.5 * h;
  const double e2 = eps * eps;
  const bool mine = lane < N;
  const int i = mine ? lane : 0;
  __syncwarp();
  const double xi = sh.x[i], yi = sh.y[i], mi = sh.m[i];
  double fx = 0.0, fy = 0.0, s3 = 0.0;
  if (G != 0.0) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
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
  if (act && mine) {
    const double im = hs_rcp(mi);
    sh.vx[i] = (mi * sh.vx[i] + hh * fx) * im;
    sh.vy[i] = (mi * sh.vy[i] + hh * fy) * im;
  }
  const double s3t = grp_sum<LPS>(mine ? s3 : 0.0);
  const double dU = (eps == 0.0 || G == 0.0) ? 0.0 : G * eps * s3t;
  const double dB = (P.policy == 0) ? -hs_barrier_force(eps, P) : 0.0;
  if (act && !(P.flags & NB_HS_FLAG_FREEZE_S)) pi = pi - (dU + dB) * hh;    // hamsoft_stepper.py:592-600
}

template <int N>
__device__ __forceinline__ void hs_t_drift(HsSh<N>& sh, double h, int lane_full, bool act) {
  const int lane = lane_full & (HsLanes<N>::LPS - 1);
  __syncwarp();
  if (act && lane < N) {
    sh.x[lane] = fma(h, sh.vx[lane], sh.x[lane]);
    sh.y[lane] = fma(h, sh.vy[lane], sh.y[lane]);
  }
}

// S(h/2) V(h/2) T(h) V(h/2) S(h/2), hamsoft_stepper.py:247-308.  Written as a two-trip loop so that the S half-flow
// (the eps* model: by far the largest piece of code) is instantiated ONCE.
template <int N>
__device__ __forceinline__ void hs_strang(HsSh<N>& sh, double& eps, double& pi, double G, double h, int lane, bool act,
                                          int& sweeps) {
  const HsPar& P = sh.P;
  if (act) hs_fold(eps, pi, P);                  // hamsoft_stepper.py:261-264
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    hs_s_half<N>(sh, eps, pi, h, lane, act, sweeps);
    if (half == 0) {
      const bool vt = act && !(P.flags & NB_HS_FLAG_S_ONLY);       // cfg._validate_S_only: S S (hamsoft_stepper.py:270-284)
      hs_v_half<N>(sh, eps, pi, G, h, lane, vt);
      hs_t_drift<N>(sh, h, lane, vt);
      hs_v_half<N>(sh, eps, pi, G, h, lane, vt);
    }
  }
  if (act) hs_fold(eps, pi, P);                  // hamsoft_stepper.py:300-303
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

// eps_target of the system in shared memory, serially on the calling lane (energy taps: twice per run)
template <int N>
__device__ __noinline__ double hs_eps_target_sh(const HsSh<N>* sh, double eps_cur) {
  double x[N], y[N], m[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { x[i] = sh->x[i]; y[i] = sh->y[i]; m[i] = sh->m[i]; }
  const HsPar P = sh->P;
  return hs_eps_target<N>(x, y, m, eps_cur, P);
}

// ---------------------------------------------------------------------------------------------
// run kernel: one group of LPS lanes per system
// ---------------------------------------------------------------------------------------------

template <int N>
static inline int hs_run_blocks(int B) {
  const int spw = 32 / HsLanes<N>::LPS;
  const int warps = (B + spw - 1) / spw;
  return (warps + 3) / 4;
}

template <int N>
__device__ __forceinline__ void hs_load_system(HsSh<N>& sh, const double* m, const double* q, const double* v,
                                               const double* hs, int sys, int lane) {
  if (lane < N) {
    sh.m[lane] = m[(size_t)sys * N + lane];
    sh.x[lane] = q[((size_t)sys * N + lane) * 2 + 0];
    sh.y[lane] = q[((size_t)sys * N + lane) * 2 + 1];
    sh.vx[lane] = v ? v[((size_t)sys * N + lane) * 2 + 0] : 0.0;
    sh.vy[lane] = v ? v[((size_t)sys * N + lane) * 2 + 1] : 0.0;
    sh.gx[lane] = 0.0; sh.gy[lane] = 0.0; sh.h[lane] = 0.0;
    sh.drx[lane] = 0.0; sh.dry[lane] = 0.0; sh.dvx[lane] = 0.0; sh.dvy[lane] = 0.0;
  }
  if (lane == 0) {
    sh.P = hs_load(hs + (size_t)sys * NB_HS_NPARAM);
    sh.inv_alpha = 1.0 / sh.P.alpha;
#pragma unroll
    for (int k = 0; k < HS_NACC; ++k) sh.acc[k] = 0.0;
    sh.acc[HA_COM_MAX] = -1.0; sh.acc[HA_VAR_MAX] = -1.0; sh.acc[HA_COS_MIN] = 2.0;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    sh.acc[HA_E0] = qnan; sh.acc[HA_L0] = qnan; sh.acc[HA_E1] = qnan; sh.acc[HA_L1] = qnan;
  }
}

// Resident CTAs per SM: the kernel is latency-bound (top stalls `wait`, `no_instruction`), so occupancy is worth more than
// spill-free code from N = 5 up: 4 (N = 5, 6) and 3 (N = 7, 8) CTAs/SM with 80-230 B of spills instead of 3 / 2 without:
// +12 % at N = 6, +30 % at N = 8; for N <= 4 more than 6 / 5 CTAs/SM gains nothing (N = 3: -4 % on the C1 batch).
template <int N>
__global__ void __launch_bounds__(128, (N <= 3 ? 6 : (N <= 4 ? 5 : (N <= 6 ? 4 : 3)))) hamsoft_run_kernel(HsArgs a) {
  constexpr int LPS = HsLanes<N>::LPS, SPW = 32 / LPS;   // lanes per system, systems per warp
  __shared__ HsSh<N> shs[4 * SPW];
  const int lane_full = threadIdx.x & 31;
  const int lane = lane_full & (LPS - 1);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp * SPW >= a.B) return;                        // warp-uniform
  stamp_begin(a.tstamp);
  // an odd tail slot shadows the warp's first system (same arithmetic, no global writes) so the warp stays converged
  const int slot = warp * SPW + lane_full / LPS;
  const bool live = slot < a.B;
  const int t = live ? slot : warp * SPW;
  const int sys = a.perm ? a.perm[t] : t;               // n_sub-sorted: neighbours in a warp have equal sub-step counts
  HsSh<N>& sh = shs[(threadIdx.x >> 5) * SPW + lane_full / LPS];
  hs_load_system<N>(sh, a.m, a.q, a.v, a.hs, sys, lane);
  double eps = a.eps_pi[2 * (size_t)sys];
  double pi = a.eps_pi[2 * (size_t)sys + 1];
  const double G = a.G;
  // cfg._validate_S_only: one sub-step per macro step whatever the frozen schedule (hamiltonian_softening_integrator.py:804-835)
  const bool s_only = ((int)a.hs[(size_t)sys * NB_HS_NPARAM + NB_HS_FLAGS] & NB_HS_FLAG_S_ONLY) != 0;
  const int n_sub = s_only ? 1 : max(1, a.n_sub ? a.n_sub[sys] : 1);
  const int n_sub_warp = SPW > 1 ? __reduce_max_sync(0xffffffffu, n_sub) : n_sub;
  const double h = a.dt / (double)n_sub;
  const double dt = a.dt;
  __syncwarp();
  // hamiltonian_softening_integrator.py:232-242: mu is raised to k (dt/theta_imp)^2 on the first step
  if (lane == 0 && a.n_steps + a.n_megno > 0 && is_finite(sh.P.k) && sh.P.k > 0.0) {
    const double mu_macro = sh.P.k * (fabs(dt) / sh.P.theta_imp) * (fabs(dt) / sh.P.theta_imp);
    if (sh.P.mu < mu_macro) sh.P.mu = mu_macro;
  }
  if (lane == 0) hs_spring_setup(sh.spr, sh.P, h);
  __syncwarp();
  const bool want_energy = (a.flags & NB_RUN_ENERGY) != 0;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  int sweeps = 0, next_sample = 0;
  // MEGNO (evolution_features.py:34-66); the tangent map uses the post-step epsilon^2 (softening_manager.py:359-366)
  double tt = 0.0, accum = 0.0;
  const int n_total = a.n_steps + a.n_megno;
  const bool mine = lane < N;
  const int bi = mine ? lane : 0;
  // ONE loop over the main steps and the MEGNO steps, so that the macro step is instantiated once:
  // step < n_steps samples step_metrics, step >= n_steps advances the tangent vectors
#pragma unroll 1
  for (int step = 0; step <= n_total; ++step) {
    if (want_energy && (step == 0 || step == a.n_steps)) {
      __syncwarp();
      const double es = hs_eps_target_sh<N>(&sh, eps);
      const double E = hs_energy<N>(sh.m, sh.x, sh.y, sh.vx, sh.vy, eps, pi, is_finite(es) ? es : sh.P.s0, sh.P, G);
      double Lz = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) Lz += sh.m[i] * (sh.x[i] * sh.vy[i] - sh.y[i] * sh.vx[i]);
      if (lane == 0) {
        if (step == 0) { sh.acc[HA_E0] = E; sh.acc[HA_L0] = Lz; }
        if (step == a.n_steps) { sh.acc[HA_E1] = E; sh.acc[HA_L1] = Lz; }
      }
    }
    if (step == a.n_steps && a.n_megno > 0) {
      __syncwarp();
      if (lane == 0) {
        double M = 0.0, cx = 0.0, cy = 0.0, ux = 0.0, uy = 0.0;
        for (int i = 0; i < N; ++i) {
          const double rx = a.raw_dr[((size_t)sys * N + i) * 2 + 0], ry = a.raw_dr[((size_t)sys * N + i) * 2 + 1];
          const double wx = a.raw_dv[((size_t)sys * N + i) * 2 + 0], wy = a.raw_dv[((size_t)sys * N + i) * 2 + 1];
          sh.drx[i] = rx; sh.dry[i] = ry; sh.dvx[i] = wx; sh.dvy[i] = wy;
          M += sh.m[i];
          cx += sh.m[i] * rx; cy += sh.m[i] * ry; ux += sh.m[i] * wx; uy += sh.m[i] * wy;
        }
        cx /= M; cy /= M; ux /= M; uy /= M;
        double nr = 0.0, nv = 0.0;
        for (int i = 0; i < N; ++i) {
          sh.drx[i] -= cx; sh.dry[i] -= cy; sh.dvx[i] -= ux; sh.dvy[i] -= uy;
          nr += sh.drx[i] * sh.drx[i] + sh.dry[i] * sh.dry[i]; nv += sh.dvx[i] * sh.dvx[i] + sh.dvy[i] * sh.dvy[i];
        }
        nr = sqrt(nr); nv = sqrt(nv);
        for (int i = 0; i < N; ++i) { sh.drx[i] /= nr; sh.dry[i] /= nr; sh.dvx[i] /= nv; sh.dvy[i] /= nv; }
      }
      __syncwarp();
    }
    if (step == n_total) break;
    // one macro step = n_sub Strang sub-steps; systems sharing a warp run to the larger count, excess predicated off
#pragma unroll 1
    for (int k = 0; k < n_sub_warp; ++k) hs_strang<N>(sh, eps, pi, G, h, lane_full, k < n_sub, sweeps);
    __syncwarp();
    if (step < a.n_steps) {
      if (a.sample_interval > 0 && step == next_sample) {   // diagnostics.py:241-285
        next_sample += a.sample_interval;
        if (lane == 0) {
          double* A = sh.acc;
          double cx = 0.0, cy = 0.0, Lt = 0.0, Li[N];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            cx += sh.m[i] * sh.x[i]; cy += sh.m[i] * sh.y[i];
            Li[i] = sh.m[i] * (sh.x[i] * sh.vy[i] - sh.y[i] * sh.vx[i]);
            Lt += Li[i];
          }
          const double mean = Lt / N;
          double var = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) var += (Li[i] - mean) * (Li[i] - mean);
          var /= N;
          hs_sample_scalars(A, sqrt(cx * cx + cy * cy), var, Lt, eps, pi, sh.P.mu);
        }
      }
    } else {
      // tangent map, body i on lane i: dr += dv dt; dv += da(dr) dt  (tangent_map.py:21-59 at the post-step epsilon)
      double drx = sh.drx[bi], dry = sh.dry[bi];
      drx = fma(sh.dvx[bi], dt, drx); dry = fma(sh.dvy[bi], dt, dry);
      __syncwarp();
      if (mine) { sh.drx[bi] = drx; sh.dry[bi] = dry; }
      __syncwarp();
      const double xi = sh.x[bi], yi = sh.y[bi], e2 = eps * eps;
      double dax = 0.0, day = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
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
      double nr = sqrt(grp_sum<LPS>(mine ? drx * drx + dry * dry : 0.0));
      tt += dt;
      if (nr < 1e-12) {
        drx /= nr; dry /= nr; dvx /= nr; dvy /= nr;
        nr = 1.0;
      }
      const double nv = sqrt(grp_sum<LPS>(mine ? dvx * dvx + dvy * dvy : 0.0));
      __syncwarp();
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
  const double sweeps_tot = grp_sum<LPS>((double)sweeps);
  __syncwarp();
  stamp_end(a.tstamp);
  if (lane != 0 || !live) return;
  bool finite = is_finite(eps) && is_finite(pi);
#pragma unroll
  for (int i = 0; i < N; ++i)
    finite = finite && is_finite(sh.x[i]) && is_finite(sh.y[i]) && is_finite(sh.vx[i]) && is_finite(sh.vy[i]);
  int st = finite ? 0 : NB_STATUS_NONFINITE;
  {
    const double R = sh.P.eps_max - sh.P.eps_min;
    if (finite && (eps < sh.P.eps_min - R || eps > sh.P.eps_max + R)) st |= NB_STATUS_EPS_OOB;
  }
  if (a.status) a.status[sys] = st;
  if (a.work) {   // counted work for the roofline: Jacobi sweeps summed over all finite-difference evaluations
    a.work[2 * (size_t)sys] = sweeps_tot;
    a.work[2 * (size_t)sys + 1] = 2.0 * (double)n_sub * (double)n_total;   // S half-flows
  }
  if (a.flags & NB_RUN_WRITE_STATE) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      a.q[((size_t)sys * N + i) * 2 + 0] = sh.x[i]; a.q[((size_t)sys * N + i) * 2 + 1] = sh.y[i];
      a.v[((size_t)sys * N + i) * 2 + 0] = sh.vx[i]; a.v[((size_t)sys * N + i) * 2 + 1] = sh.vy[i];
    }
    a.eps_pi[2 * (size_t)sys] = eps;
    a.eps_pi[2 * (size_t)sys + 1] = pi;
  }
  if (a.dyn) hs_write_dyn(a.dyn + (size_t)sys * NB_N_DYN, sh.acc, want_energy, megno, lyap, t_end);
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

// eps*(q), its gradient and H_ext for B systems (parity taps; one warp per system, same group code as the run kernel)
template <int N>
__global__ void __launch_bounds__(128) hamsoft_probe_kernel(const double* m_, const double* q_, const double* v_,
                                                            double G, int B, const double* eps_pi, const double* hs,
                                                            double* out /*[B][3+2N]: eps*, H, fallback, grad*/) {
  constexpr int LPS = HsLanes<N>::LPS, SPW = 32 / LPS;
  __shared__ HsSh<N> shs[4 * SPW];
  const int lane_full = threadIdx.x & 31;
  const int lane = lane_full & (LPS - 1);
  const int sys = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (sys >= B) return;
  HsSh<N>& sh = shs[(threadIdx.x >> 5) * SPW + lane_full / LPS];   // with LPS = 16 both half warps take the same system
  hs_load_system<N>(sh, m_, q_, v_, hs, sys, lane);
  __syncwarp();
  const double eps = eps_pi[2 * (size_t)sys], pi = eps_pi[2 * (size_t)sys + 1];
  bool fb;
  int sweeps = 0;
  const double es = hs_eps_star_and_grad<N>(sh, eps, lane_full, fb, sweeps);
  const double H = hs_energy<N>(sh.m, sh.x, sh.y, sh.vx, sh.vy, eps, pi, es, sh.P, G);
  if (lane_full == 0) {
    double* o = out + (size_t)sys * (3 + 2 * N);
    o[0] = es; o[1] = H; o[2] = fb ? 1.0 : 0.0;
    for (int i = 0; i < N; ++i) { o[3 + 2 * i] = sh.gx[i]; o[4 + 2 * i] = sh.gy[i]; }
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
                const double* raw_dv, double* eps_pi, const double* hs, double* dyn, int32_t* status, double* work,
                unsigned long long* tstamp, cudaStream_t st) {
  NvtxRange r("nb_ensemble_run: ham_soft");
  HsArgs a{m, q, v, G, B, flags, dt, n_steps, sample_interval, n_megno, n_sub, perm, raw_dr, raw_dv, eps_pi, hs, dyn,
           status, work, tstamp};
  if (N > NB_MAX_N && N <= NB_MAX_N_MID) return hamsoft_mid_run(a, N, st);    // 9..64 bodies: one CTA per system
  NB_HS_DISPATCH(N, (hamsoft_run_kernel<NN><<<hs_run_blocks<NN>(a.B), 128, 0, st>>>(a)));
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int hamsoft_setup(const double* m, const double* q, double G, int B, int N, unsigned flags, double dt, double* hs,
                  double* eps_pi, int32_t* n_sub, cudaStream_t st) {
  if (N > NB_MAX_N && N <= NB_MAX_N_MID) return hamsoft_mid_setup(m, q, G, B, N, flags, dt, hs, eps_pi, n_sub, st);
  const int blocks = (B + 63) / 64;
  NB_HS_DISPATCH(N, (hamsoft_setup_kernel<NN><<<blocks, 64, 0, st>>>(m, q, G, B, flags, dt, hs, eps_pi, n_sub)));
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int hamsoft_probe(const double* m, const double* q, const double* v, double G, int B, int N, const double* eps_pi,
                  const double* hs, double* out, cudaStream_t st) {
  if (N > NB_MAX_N && N <= NB_MAX_N_MID) return hamsoft_mid_probe(m, q, v, G, B, N, eps_pi, hs, out, st);
  const int blocks = (B + 3) / 4;
  NB_HS_DISPATCH(N, (hamsoft_probe_kernel<NN><<<blocks, 128, 0, st>>>(m, q, v, G, B, eps_pi, hs, out)));
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
