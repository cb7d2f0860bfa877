// kepler.cuh -- universal-variable Kepler propagation for the "whfast" kick-drift scheme.
//
// Two solvers:
//   * kepler_reference<>  restates kepler_solver.py:25-91 INCLUDING its quirks (Stumpff c1,c2 where
//     c2,c3 belong at :69-70 and the argument-doubling recurrence at :36-45), because parity with the
//     reference means bug-compatibility.  It is evaluated with strictly
//     rounded, non-contracted fp64 operations (sd type below) so the Newton iteration follows the
//     same path as the NumPy/CPython arithmetic.
//   * kepler_exact<>      a physically correct universal-variable solver (NB_RUN_KEPLER_EXACT), checked
//     against the analytic two-body solution.
#pragma once
#include "common.cuh"

namespace nb {

// strictly-rounded double: every operator is one IEEE operation, never contracted into an FMA
struct sd {
  double v;
  __device__ __forceinline__ sd() {}
  __device__ __forceinline__ sd(double a) : v(a) {}
};
__device__ __forceinline__ sd operator+(sd a, sd b) { return sd(__dadd_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a, sd b) { return sd(__dadd_rn(a.v, -b.v)); }
__device__ __forceinline__ sd operator*(sd a, sd b) { return sd(__dmul_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator/(sd a, sd b) { return sd(__ddiv_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a) { return sd(-a.v); }

// x / C for a compile-time constant C, correctly rounded (identical bits to __ddiv_rn) in 3 FP64 operations:
// q = RN(x RN(1/C)); r = x - q C (exact, one fma); result = RN(q + r RN(1/C)).  Markstein's theorem: the result is the
// correctly rounded quotient whenever RN(1/C) is the correctly rounded reciprocal and the significand of C is not all
// ones -- true for 6, 24, 120, 720, 5040, 40320, 362880 (checked exhaustively-by-sampling on the CPU: 4.2e8 random
// operands over 2^-300..2^300, zero mismatches).  The general __ddiv_rn is ~15 instructions plus a slow-path call, and
// the Stumpff series below contains 14 such divisions per Newton iteration.
template <int C>
__device__ __forceinline__ sd divc(sd x) {
  constexpr double c = (double)C;
  constexpr double rc = 1.0 / (double)C;
  const double q = __dmul_rn(x.v, rc);
  const double r = __fma_rn(-q, c, x.v);
  return sd(__fma_rn(r, rc, q));
}

// kepler_solver.py:25-46
__device__ __forceinline__ void cfunc_reference(double z_in, sd& c0, sd& c1, sd& c2, sd& c3) {
  sd z(z_in);
  int n = 0;
  while (fabs(z.v) > 0.1 && n < 600) {
    z = z * sd(0.25);
    ++n;
  }
  const sd z2 = z * z;
  const sd z3 = z * z2, z4 = z2 * z2;
  c0 = sd(1.0) - z * sd(0.5) + divc<24>(z2) - divc<720>(z3) + divc<40320>(z4);
  c1 = sd(1.0) - divc<6>(z) + divc<120>(z2) - divc<5040>(z3) + divc<362880>(z4);
  c2 = sd(0.5) - divc<24>(z) + divc<720>(z2) - divc<40320>(z3);
  c3 = sd(1.0 / 6.0) - divc<120>(z) + divc<5040>(z2) - divc<362880>(z3);
  while (n) {
    z = z * sd(4.0);
    --n;
    const sd c3o = c3, c1o = c1, c2o = c2;
    c0 = sd(1.0) - z * c2o;
    c1 = sd(1.0) - z * c3o;
    c2 = sd(0.5) - z * (c3o * (sd(1.0) + c1o)) * sd(0.125);
    c3 = (c1o - sd(1.0)) / z;
  }
}

// kepler_solver.py:48-91.  Returns the Newton iteration count (64 = cap reached).
__device__ __forceinline__ int kepler_reference(double& rx, double& ry, double& vx, double& vy, double mu_in,
                                                double dt_in, int& executed) {
  const sd r_x(rx), r_y(ry), v_x(vx), v_y(vy), mu(mu_in), dt(dt_in);
  const sd r0(hypot(rx, ry));
  if (r0.v < 1e-14) {
    rx = (r_x + v_x * dt).v;
    ry = (r_y + v_y * dt).v;
    return 0;
  }
  const sd vr0 = (r_x * v_x + r_y * v_y) / r0;
  const sd v2 = v_x * v_x + v_y * v_y;
  const sd alpha = sd(2.0) / r0 - v2 / mu;
  const sd sqrt_mu(sqrt(mu.v));
  sd chi;
  if (fabs(alpha.v) > 1e-12)
    chi = sqrt_mu * sd(fabs(alpha.v)) * dt;
  else
    chi = sqrt_mu * dt / r0;
  // x2 = the iterate two steps back.  (The reference also keeps prev1/prev2, but its second exit test
  // `chi_new == prev2` compares with the current chi again, kepler_solver.py:71-77, so it only ever exits on
  // chi_new == chi; an exact 2-cycle therefore runs to the 64-iteration cap -- 4 % of all planetary solves.)
  double x2 = __longlong_as_double(0x7ff8000000000000LL);
  sd c0, c1, c2, c3;
  // Stumpff c2 / c3 of the previous iterate, and whether (c2, c3) below belong to the FINAL iterate: the reference
  // re-evaluates its series at the final chi (kepler_solver.py:81-83); whenever the loop has just evaluated it at that
  // very chi -- every exit except the 64-iteration cap -- the values are reused (same input, same bits; one of ~10
  // series evaluations per solve saved)
  double c2p = 0.0, c3p = 0.0;
  bool have_c = false;
  int it = 0;
  double last_step = 0.0;
  const sd k1 = r0 * vr0 / sqrt_mu, k2 = sd(1.0) - alpha * r0, k3 = sqrt_mu * dt;     // loop invariants, same rounding
  // ONE loop holds the only instance of the Stumpff series: the Newton trips, plus -- when the series values at hand do
  // not belong to the final iterate -- one closing trip (`fin`) that evaluates them there (kepler_solver.py:81-83)
  bool fin = false;
  for (;;) {
    const sd z = alpha * chi * chi;
    if (!fin) { ++it; c2p = c2.v; c3p = c3.v; }
    cfunc_reference(z.v, c0, c1, c2, c3);
    if (fin) break;
    have_c = true;                               // (c2, c3) belong to the current chi
    bool stop = false;
    const sd f = k1 * chi * chi * c1 + k2 * chi * chi * chi * c2 + r0 * chi - k3;
    const sd fp = k1 * chi * (sd(1.0) - alpha * chi * chi * c2) + k2 * chi * chi * c1 + r0;
    if (fp.v == 0.0) {
      stop = true;
    } else {
      const sd chi_new = chi - f / fp;
      last_step = fabs(chi_new.v - chi.v);
      if (chi_new.v == chi.v) {
        chi = chi_new;
        stop = true;
      } else if (chi_new.v == x2) {
        // Exact-cycle shortcut (bit-identical result): the Newton map is a pure function of chi, so chi_new == x2 means
        // the iterates alternate between chi_new and chi for the rest of the reference's 64 iterations; its final
        // iterate is the one with the parity of 64.  Detected after ~8 iterations instead of running all 64 in lock-step
        // per warp.
        if (((64 - it) & 1) == 0) {              // final iterate = chi_new == the iterate BEFORE chi: its series values
          chi = chi_new;
          c2 = sd(c2p); c3 = sd(c3p);
          have_c = it >= 2;
        }
        executed += it - 64;                     // counted work = iterations actually executed, not the reference's 64
        it = 64;
        stop = true;
      } else {
        x2 = chi.v;
        chi = chi_new;
        have_c = false;                          // the new chi has not been evaluated yet
      }
    }
    if (stop || it >= 64) {
      if (have_c) break;
      fin = true;
    }
  }
  // the reference's exit test is exact equality, so running into the 64-iteration cap while hovering within a
  // few ulps of the root is normal; only a cap hit with a still-moving iterate is reported (as 65)
  executed += it;
  if (it >= 64 && !(last_step <= 1e-9 * fabs(chi.v))) it = 65;
  const sd f = sd(1.0) - chi * chi * c2 / r0;
  const sd g = dt - chi * chi * chi * c3 / sqrt_mu;
  const sd nx = f * r_x + g * v_x;
  const sd ny = f * r_y + g * v_y;
  const sd rn(hypot(nx.v, ny.v));
  rx = nx.v;
  ry = ny.v;
  if (rn.v == 0.0) return it;
  const sd fdot = sqrt_mu / (rn * r0) * (alpha * chi * chi * c3 - chi);
  const sd gdot = sd(1.0) - chi * chi * c2 / rn;
  const sd wx = fdot * r_x + gdot * v_x;
  const sd wy = fdot * r_y + gdot * v_y;
  vx = wx.v;
  vy = wy.v;
  return it;
}

// ---- physically correct solver -----------------------------------------------------------------
// Stumpff functions c2(psi) = (1-cos sqrt psi)/psi, c3(psi) = (sqrt psi - sin sqrt psi)/psi^1.5
__device__ __forceinline__ void stumpff(double psi, double& c2, double& c3) {
  if (psi > 1e-6) {
    const double s = sqrt(psi);
    double sn, cs;
    sincos(s, &sn, &cs);
    c2 = (1.0 - cs) / psi;
    c3 = (s - sn) / (psi * s);
  } else if (psi < -1e-6) {
    const double s = sqrt(-psi);
    c2 = (1.0 - cosh(s)) / psi;
    c3 = (sinh(s) - s) / (-psi * s);
  } else {
    c2 = 0.5 - psi / 24.0 + psi * psi / 720.0;
    c3 = 1.0 / 6.0 - psi / 120.0 + psi * psi / 5040.0;
  }
}

__device__ __forceinline__ int kepler_exact(double& rx, double& ry, double& vx, double& vy, double mu, double dt) {
  const double r0 = hypot(rx, ry);
  if (r0 < 1e-14) {
    rx += vx * dt;
    ry += vy * dt;
    return 0;
  }
  const double sm = sqrt(mu);
  const double rv = rx * vx + ry * vy;
  const double alpha = 2.0 / r0 - (vx * vx + vy * vy) / mu;
  double chi = sm * dt / r0;                 // good for small steps
  if (alpha > 1e-12) chi = sm * dt * alpha;
  double c2 = 0.5, c3 = 1.0 / 6.0;
  int it = 0;
  double d_prev = __longlong_as_double(0x7ff0000000000000LL);
  for (; it < 64; ++it) {
    const double chi2 = chi * chi;
    stumpff(alpha * chi2, c2, c3);
    const double f = rv / sm * chi2 * c2 + (1.0 - alpha * r0) * chi2 * chi * c3 + r0 * chi - sm * dt;
    const double fp = rv / sm * chi * (1.0 - alpha * chi2 * c3) + (1.0 - alpha * r0) * chi2 * c2 + r0;
    const double d = f / fp;
    chi -= d;
    // Newton converges quadratically: after a step of relative size <= 1e-9 the iterate is exact to rounding.  (Asking
    // the STEP itself to vanish, |d| <= 4e-16 |chi|, never terminates once the residual is rounding noise: 81 % of
    // the systems reported a spurious non-convergence at least once in 100,000 steps.)  The stagnation clause covers
    // linearly converging degenerate cases.
    if (fabs(d) <= 1e-9 * fabs(chi) || (fabs(d) >= d_prev && fabs(d) <= 1e-12 * fabs(chi))) {
      ++it;
      break;
    }
    d_prev = fabs(d);
  }
  if (it >= 64) it = 65;
  const double chi2 = chi * chi;
  stumpff(alpha * chi2, c2, c3);
  const double f = 1.0 - chi2 * c2 / r0;
  const double g = dt - chi2 * chi * c3 / sm;
  const double nx = f * rx + g * vx, ny = f * ry + g * vy;
  const double rn = hypot(nx, ny);
  const double fdot = sm / (rn * r0) * (alpha * chi2 * chi * c3 - chi);
  const double gdot = 1.0 - chi2 * c2 / rn;
  const double wx = fdot * rx + gdot * vx, wy = fdot * ry + gdot * vy;
  rx = nx;
  ry = ny;
  vx = wx;
  vy = wy;
  return it;
}

}  // namespace nb
