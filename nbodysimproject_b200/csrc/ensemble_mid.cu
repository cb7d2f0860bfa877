// ensemble_mid.cu -- systems of 9 .. 64 bodies: ONE CTA PER SYSTEM, one body per thread, positions in shared memory.
//
// The reference accepts any body count in fp64 (simulation.py:39-162, simulation_state.py:98-144); the register-resident
// thread-per-system kernels stop at N = 8 (N(N-1)/2 unrolled pairs), the fp32 large-N path starts where fp64 per-body
// accuracy no longer matters.  In between, thread i of a 64-thread CTA owns body i: it reads every other body's position
// from shared memory (broadcast reads), accumulates its own acceleration over j in ASCENDING order with the same
// operations as pair_small.cuh (so a body's acceleration is bit-identical to the small-N kernels'), and integrates its
// own body.  Two barriers per force evaluation.  Order-sensitive scalar reductions (COM, momenta, sampling statistics,
// the double-double energies, the tangent norms) are done by thread 0 in ascending body order -- the same summation
// order as the thread-per-system kernels and the reference's loops; this path serves single simulations and small
// batches of mid-sized systems, not the throughput benchmarks, so the serial O(N) / O(N^2) tails are irrelevant.
// Restates the same reference lines as ensemble_run.cuh / ensemble_misc.cu (integrator.py:78-136, 200-227,
// integration_scheme_base.py:41-54, 129-192, timestep_manager.py:139-253, tangent_map.py:21-59,
// evolution_features.py:34-66, diagnostics.py:241-285, 457-549, dynamical_features.py:27-155, whfast_scheme.py:22-123,
// simulation.py:487-534).
#include "pair_small.cuh"
#include "kepler.cuh"
#include "args.cuh"

namespace nb {

constexpr int MID_MAX = NB_MAX_N_MID;      // 64
constexpr int MID_THREADS = 64;

struct MidSh {
  double x[MID_MAX], y[MID_MAX], m[MID_MAX], gm[MID_MAX];
  double a[MID_MAX], b[MID_MAX], c[MID_MAX], d[MID_MAX];   // exchange buffers (velocities, tangent vectors, partial results)
  double red[MID_MAX];
  double scal[8];
};

__device__ __forceinline__ double mid_drift_of(double a0, double a1) {
  if (is_finite(a0) && fabs(a0) > 0.0 && is_finite(a1)) return fabs((a1 - a0) / a0);
  if (is_finite(a0) && is_finite(a1)) return fabs(a1 - a0);
  return __longlong_as_double(0x7ff0000000000000LL);
}

// acceleration of body i from the positions in shared memory (and optionally the variational acceleration)
template <bool TANGENT>
__device__ __forceinline__ void mid_accel(const MidSh& sh, int N, int i, double xi, double yi, double eps2, double& ax,
                                          double& ay, double dri_x, double dri_y, double& dax, double& day) {
  ax = 0.0; ay = 0.0;
  if (TANGENT) { dax = 0.0; day = 0.0; }
  for (int j = 0; j < N; ++j) {
    if (j == i) continue;
    const double dx = xi - sh.x[j], dy = yi - sh.y[j];
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
    const double cj = sh.gm[j] * w3;
    ax = fma(-cj, dx, ax);
    ay = fma(-cj, dy, ay);
    if (TANGENT) {
      const double ex = sh.c[j] - dri_x, ey = sh.d[j] - dri_y;
      const double dot = -fma(dx, ex, dy * ey);
      const double c5 = 3.0 * dot * w2 * w3;
      dax = fma(sh.gm[j], fma(ex, w3, c5 * dx), dax);
      day = fma(sh.gm[j], fma(ey, w3, c5 * dy), day);
    }
  }
}

struct MidBody {
  double x, y, vx, vy, ax, ay;
};

// publish positions, evaluate the acceleration of the own body
template <bool TANGENT>
__device__ __forceinline__ void mid_force(MidSh& sh, int N, int i, bool mine, MidBody& s, double eps2, double drx,
                                          double dry, double& dax, double& day) {
  if (mine) { sh.x[i] = s.x; sh.y[i] = s.y; if (TANGENT) { sh.c[i] = drx; sh.d[i] = dry; } }
  __syncthreads();
  if (mine) mid_accel<TANGENT>(sh, N, i, s.x, s.y, eps2, s.ax, s.ay, drx, dry, dax, day);
  __syncthreads();
}

// pseudo-Jacobi Kepler drift (whfast_scheme.py:22-37, simulation.py:487-534): the two prefix sums are sequential
// (each body is referred to the centre of mass of its predecessors' NEW positions) and run on thread 0; the N-1 Kepler
// solves run one per thread.
template <bool EXACT>
__device__ __forceinline__ int mid_kepler_drift(MidSh& sh, int N, int i, bool mine, MidBody& s, double G, double tau,
                                                int& iters) {
  if (mine) { sh.x[i] = s.x; sh.y[i] = s.y; sh.a[i] = s.vx; sh.b[i] = s.vy; }
  __syncthreads();
  if (threadIdx.x == 0) {           // to_jacobi: c,d <- reference position / velocity of body k's predecessors
    double Rx = sh.m[0] * sh.x[0], Ry = sh.m[0] * sh.y[0], Vx = sh.m[0] * sh.a[0], Vy = sh.m[0] * sh.b[0], M = sh.m[0];
    sh.c[0] = 0.0; sh.d[0] = 0.0; sh.red[0] = sh.m[0];
    for (int k = 1; k < N; ++k) {
      const double jx = sh.x[k] - Rx / M, jy = sh.y[k] - Ry / M, ju = sh.a[k] - Vx / M, jv = sh.b[k] - Vy / M;
      Rx = __dadd_rn(Rx, __dmul_rn(sh.m[k], sh.x[k]));
      Ry = __dadd_rn(Ry, __dmul_rn(sh.m[k], sh.y[k]));
      Vx = __dadd_rn(Vx, __dmul_rn(sh.m[k], sh.a[k]));
      Vy = __dadd_rn(Vy, __dmul_rn(sh.m[k], sh.b[k]));
      sh.red[k] = M;                // cumulative mass of the predecessors
      M += sh.m[k];
      sh.x[k] = jx; sh.y[k] = jy; sh.a[k] = ju; sh.b[k] = jv;
    }
  }
  __syncthreads();
  int it = 0;
  if (mine) {
    double rx = sh.x[i], ry = sh.y[i], ux = sh.a[i], uy = sh.b[i];
    if (i == 0) {
      rx = __dadd_rn(rx, __dmul_rn(ux, tau));
      ry = __dadd_rn(ry, __dmul_rn(uy, tau));
    } else {
      const double mu = G * (sh.red[i] + sh.m[i]);
      int done = 0;
      it = EXACT ? kepler_exact(rx, ry, ux, uy, mu, tau) : kepler_reference(rx, ry, ux, uy, mu, tau, done);
      iters += EXACT ? min(it, 64) : done;
    }
    sh.x[i] = rx; sh.y[i] = ry; sh.a[i] = ux; sh.b[i] = uy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {           // from_jacobi
    double Rx = sh.m[0] * sh.x[0], Ry = sh.m[0] * sh.y[0], Vx = sh.m[0] * sh.a[0], Vy = sh.m[0] * sh.b[0], M = sh.m[0];
    for (int k = 1; k < N; ++k) {
      const double px = sh.x[k] + Rx / M, py = sh.y[k] + Ry / M, pu = sh.a[k] + Vx / M, pv = sh.b[k] + Vy / M;
      sh.x[k] = px; sh.y[k] = py; sh.a[k] = pu; sh.b[k] = pv;
      Rx = __dadd_rn(Rx, __dmul_rn(sh.m[k], px));
      Ry = __dadd_rn(Ry, __dmul_rn(sh.m[k], py));
      Vx = __dadd_rn(Vx, __dmul_rn(sh.m[k], pu));
      Vy = __dadd_rn(Vy, __dmul_rn(sh.m[k], pv));
      M += sh.m[k];
    }
  }
  __syncthreads();
  if (mine) { s.x = sh.x[i]; s.y = sh.y[i]; s.vx = sh.a[i]; s.vy = sh.b[i]; }
  // worst iteration count of the CTA
  __syncthreads();
  if (mine) sh.red[i] = (double)it;
  __syncthreads();
  int worst = 0;
  for (int k = 0; k < N; ++k) worst = max(worst, (int)sh.red[k]);
  __syncthreads();
  return worst;
}

// whfast_scheme.py:39-69 "interaction acceleration" (corrector only), serially on thread 0 into sh.c / sh.d
__device__ void mid_wh_interaction(MidSh& sh, int N, double G, double eps2) {
  // Jacobi positions in a/b, cumulative masses in red
  double Rx = sh.m[0] * sh.x[0], Ry = sh.m[0] * sh.y[0], M = sh.m[0];
  sh.a[0] = sh.x[0]; sh.b[0] = sh.y[0]; sh.red[0] = sh.m[0];
  for (int k = 1; k < N; ++k) {
    sh.a[k] = sh.x[k] - Rx / M; sh.b[k] = sh.y[k] - Ry / M;
    Rx += sh.m[k] * sh.x[k]; Ry += sh.m[k] * sh.y[k]; M += sh.m[k];
    sh.red[k] = sh.red[k - 1] + sh.m[k];
  }
  for (int k = 0; k < N; ++k) { sh.c[k] = 0.0; sh.d[k] = 0.0; }
  for (int k = 2; k < N; ++k) {
    const double rn2 = sh.a[k] * sh.a[k] + sh.b[k] * sh.b[k] + eps2;
    if (rn2 > 0.0) {
      const double f = G * sh.red[k - 1] / (rn2 * sqrt(rn2));
      const double gx = f * sh.a[k], gy = f * sh.b[k];
      for (int l = 0; l < k; ++l) {
        const double w = sh.m[k] * (sh.m[l] / sh.red[k - 1]);
        sh.c[l] -= w * gx; sh.d[l] -= w * gy;
      }
      sh.c[k] += sh.red[k - 1] * gx; sh.d[k] += sh.red[k - 1] * gy;
    }
  }
  for (int k = 1; k < N; ++k)
    for (int l = k + 1; l < N; ++l) {
      const double dx = sh.x[l] - sh.x[k], dy = sh.y[l] - sh.y[k];
      const double r2 = dx * dx + dy * dy + eps2;
      const double w = G / (r2 * sqrt(r2));
      sh.c[k] -= sh.m[l] * w * dx; sh.d[k] -= sh.m[l] * w * dy;
      sh.c[l] += sh.m[k] * w * dx; sh.d[l] += sh.m[k] * w * dy;
    }
}

template <int MODE, bool EXACT, bool TANGENT>
__device__ __forceinline__ int mid_substep(MidSh& sh, int N, int i, bool mine, MidBody& s, double G, double eps2, double h,
                                           double drx, double dry, double& dax, double& day, int& iters) {
  double d1 = 0.0, d2 = 0.0;
  int kep = 0;
  if (MODE == NB_MODE_VERLET) {
    const double h2 = 0.5 * h;
    s.vx = fma(h2, s.ax, s.vx); s.vy = fma(h2, s.ay, s.vy);
    s.x = fma(h, s.vx, s.x); s.y = fma(h, s.vy, s.y);
    mid_force<TANGENT>(sh, N, i, mine, s, eps2, drx, dry, dax, day);
    s.vx = fma(h2, s.ax, s.vx); s.vy = fma(h2, s.ay, s.vy);
  } else if (MODE == NB_MODE_YOSHIDA4) {
    const double cbrt2 = 1.2599210498948731648;
    const double w1 = 1.0 / (2.0 - cbrt2), w2 = -cbrt2 / (2.0 - cbrt2);
    const double ha = w1 * h, hb = w2 * h;
    const double hab = 0.5 * ha + 0.5 * hb;          // merged adjacent half kicks, exactly as ensemble_run.cuh substep<>
    s.vx = fma(0.5 * ha, s.ax, s.vx); s.vy = fma(0.5 * ha, s.ay, s.vy);
    s.x = fma(ha, s.vx, s.x); s.y = fma(ha, s.vy, s.y);
    mid_force<false>(sh, N, i, mine, s, eps2, 0.0, 0.0, d1, d2);
    s.vx = fma(hab, s.ax, s.vx); s.vy = fma(hab, s.ay, s.vy);
    s.x = fma(hb, s.vx, s.x); s.y = fma(hb, s.vy, s.y);
    mid_force<false>(sh, N, i, mine, s, eps2, 0.0, 0.0, d1, d2);
    s.vx = fma(hab, s.ax, s.vx); s.vy = fma(hab, s.ay, s.vy);
    s.x = fma(ha, s.vx, s.x); s.y = fma(ha, s.vy, s.y);
    mid_force<TANGENT>(sh, N, i, mine, s, eps2, drx, dry, dax, day);
    s.vx = fma(0.5 * ha, s.ax, s.vx); s.vy = fma(0.5 * ha, s.ay, s.vy);
  } else {      // whfast: Kepler(h/2) . kick(h) . Kepler(h/2)   whfast_scheme.py:71-93
    kep = mid_kepler_drift<EXACT>(sh, N, i, mine, s, G, 0.5 * h, iters);
    if (EXACT) {
      if (mine) { sh.x[i] = s.x; sh.y[i] = s.y; }
      __syncthreads();
      if (threadIdx.x == 0) mid_wh_interaction(sh, N, G, eps2);
      __syncthreads();
      if (mine) { s.ax = sh.c[i]; s.ay = sh.d[i]; }
      __syncthreads();
    } else {
      mid_force<false>(sh, N, i, mine, s, eps2, 0.0, 0.0, d1, d2);
    }
    s.vx = fma(h, s.ax, s.vx); s.vy = fma(h, s.ay, s.vy);
    kep = max(kep, mid_kepler_drift<EXACT>(sh, N, i, mine, s, G, 0.5 * h, iters));
    if (TANGENT) mid_force<true>(sh, N, i, mine, s, eps2, drx, dry, dax, day);
  }
  return kep;
}

// T + U in double-double, L in strictly rounded fp64 (energy_kernel of ensemble_misc.cu, for one system, thread 0)
__device__ void mid_energy(const MidSh& sh, int N, double G, double eps, double& E, double& L) {
  dd T = dd_make(0.0);
  double Lz = 0.0;
  for (int i = 0; i < N; ++i) {
    const double vx = sh.a[i], vy = sh.b[i];
    dd v2 = dd_add(two_prod(vx, vx), two_prod(vy, vy));
    T = dd_add(T, dd_mul_d(dd_mul_d(v2, sh.m[i]), 0.5));
    Lz += sh.m[i] * __dadd_rn(__dmul_rn(sh.x[i], vy), -__dmul_rn(sh.y[i], vx));
  }
  dd S = dd_make(0.0);
  const dd e2 = two_prod(eps, eps);
  if (G != 0.0)
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) {
        dd dx = two_sum(sh.x[i], -sh.x[j]);
        dd dy = two_sum(sh.y[i], -sh.y[j]);
        dd r2 = dd_add(dd_add(dd_mul(dx, dx), dd_mul(dy, dy)), e2);
        if (!(r2.hi > 0.0)) r2 = dd_make(1e-300);
        S = dd_add(S, dd_mul(two_prod(sh.m[i], sh.m[j]), dd_div(dd_make(1.0), dd_sqrt(r2))));
      }
  E = dd_to_double(T) + dd_to_double(dd_mul_d(S, -G));
  L = Lz;
}

// ---------------------------------------------------------------------------------------------
// run kernel: E0 -> n_steps (sampling) -> E1 -> n_megno tangent steps -> features, one CTA per system
// ---------------------------------------------------------------------------------------------
template <int MODE, bool EXACT>
__global__ void __launch_bounds__(MID_THREADS) mid_run_kernel(RunArgs a, int N) {
  __shared__ MidSh sh;
  const int sys = blockIdx.x;
  const int i = threadIdx.x;
  const bool mine = i < N;
  const double G = a.G;
  MidBody s = {0, 0, 0, 0, 0, 0};
  if (mine) {
    sh.m[i] = a.m[(size_t)sys * N + i];
    sh.gm[i] = G * sh.m[i];
    s.x = a.q[((size_t)sys * N + i) * 2 + 0]; s.y = a.q[((size_t)sys * N + i) * 2 + 1];
    s.vx = a.v[((size_t)sys * N + i) * 2 + 0]; s.vy = a.v[((size_t)sys * N + i) * 2 + 1];
  }
  const double eps = a.eps[sys];
  const double eps2 = eps * eps;
  const int n_sub = a.n_sub ? max(1, a.n_sub[sys]) : 1;
  const double h = a.dt / (double)n_sub;
  const double dt = a.dt;
  const bool want_energy = (a.flags & NB_RUN_ENERGY) != 0 && a.dyn != nullptr;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double E0 = nan, L0 = nan, E1 = nan, L1 = nan;
  double d1 = 0.0, d2 = 0.0;
  int kep_worst = 0, kep_iters = 0;
  __syncthreads();
  auto energy = [&](double& E, double& L) {
    if (mine) { sh.x[i] = s.x; sh.y[i] = s.y; sh.a[i] = s.vx; sh.b[i] = s.vy; }
    __syncthreads();
    if (i == 0) mid_energy(sh, N, G, eps, E, L);
    __syncthreads();
  };
  if (want_energy) energy(E0, L0);
  if (MODE != NB_MODE_WHFAST) mid_force<false>(sh, N, i, mine, s, eps2, 0.0, 0.0, d1, d2);   // FSAL start
  // ---- main loop with step_metrics sampling (thread 0 keeps the accumulators)
  double com_sum = 0.0, com_max = -1.0, var_sum = 0.0, var_max = -1.0, cos_sum = 0.0, cos_min = 2.0, th_sum = 0.0;
  double Lfirst = 0.0;
  bool have_first = false, cos_nan = false;
  int n_samp = 0, next_sample = 0;
  const double theta_eps = (eps != 0.0) ? atan2(0.0, eps) : nan;
  for (int step = 0; step < a.n_steps; ++step) {
    for (int k = 0; k < n_sub; ++k)
      kep_worst = max(kep_worst, mid_substep<MODE, EXACT, false>(sh, N, i, mine, s, G, eps2, h, 0.0, 0.0, d1, d2, kep_iters));
    if (a.sample_interval > 0 && step == next_sample) {       // diagnostics.py:241-285
      next_sample += a.sample_interval;
      if (mine) { sh.x[i] = s.x; sh.y[i] = s.y; sh.a[i] = s.vx; sh.b[i] = s.vy; }
      __syncthreads();
      if (i == 0) {
        double cx = 0.0, cy = 0.0, Lt = 0.0;
        for (int k = 0; k < N; ++k) {
          cx += sh.m[k] * sh.x[k]; cy += sh.m[k] * sh.y[k];
          sh.red[k] = sh.m[k] * (sh.x[k] * sh.b[k] - sh.y[k] * sh.a[k]);
          Lt += sh.red[k];
        }
        const double com = sqrt(cx * cx + cy * cy);
        const double mean = Lt / N;
        double var = 0.0;
        for (int k = 0; k < N; ++k) var += (sh.red[k] - mean) * (sh.red[k] - mean);
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
      __syncthreads();
    }
  }
  if (want_energy) energy(E1, L1);
  // ---- MEGNO (evolution_features.py:34-66)
  double megno = 2.0, lyap = inf, t_end = 0.0;
  if (a.n_megno > 0) {
    double drx = 0.0, dry = 0.0, dvx = 0.0, dvy = 0.0, dax = 0.0, day = 0.0;
    if (mine) {
      drx = a.raw_dr[((size_t)sys * N + i) * 2 + 0]; dry = a.raw_dr[((size_t)sys * N + i) * 2 + 1];
      dvx = a.raw_dv[((size_t)sys * N + i) * 2 + 0]; dvy = a.raw_dv[((size_t)sys * N + i) * 2 + 1];
      sh.a[i] = drx; sh.b[i] = dry; sh.c[i] = dvx; sh.d[i] = dvy;
    }
    __syncthreads();
    if (i == 0) {       // COM removal and normalisation of the raw draws, ascending sums
      double M = 0.0, cx = 0.0, cy = 0.0, ux = 0.0, uy = 0.0;
      for (int k = 0; k < N; ++k) {
        M += sh.m[k];
        cx += sh.m[k] * sh.a[k]; cy += sh.m[k] * sh.b[k]; ux += sh.m[k] * sh.c[k]; uy += sh.m[k] * sh.d[k];
      }
      cx /= M; cy /= M; ux /= M; uy /= M;
      double nr = 0.0, nv = 0.0;
      for (int k = 0; k < N; ++k) {
        const double rx = sh.a[k] - cx, ry = sh.b[k] - cy, wx = sh.c[k] - ux, wy = sh.d[k] - uy;
        nr += rx * rx + ry * ry; nv += wx * wx + wy * wy;
      }
      sh.scal[0] = cx; sh.scal[1] = cy; sh.scal[2] = ux; sh.scal[3] = uy; sh.scal[4] = sqrt(nr); sh.scal[5] = sqrt(nv);
    }
    __syncthreads();
    if (mine) {
      drx = (drx - sh.scal[0]) / sh.scal[4]; dry = (dry - sh.scal[1]) / sh.scal[4];
      dvx = (dvx - sh.scal[2]) / sh.scal[5]; dvy = (dvy - sh.scal[3]) / sh.scal[5];
    }
    __syncthreads();
    double tt = 0.0, accum = 0.0;
    for (int step = 0; step < a.n_megno; ++step) {
      for (int k = 0; k < n_sub - 1; ++k)
        kep_worst = max(kep_worst, mid_substep<MODE, EXACT, false>(sh, N, i, mine, s, G, eps2, h, 0.0, 0.0, d1, d2, kep_iters));
      drx = fma(dvx, dt, drx); dry = fma(dvy, dt, dry);
      kep_worst = max(kep_worst, mid_substep<MODE, EXACT, true>(sh, N, i, mine, s, G, eps2, h, drx, dry, dax, day, kep_iters));
      dvx = fma(dax, dt, dvx); dvy = fma(day, dt, dvy);
      tt += dt;
      if (mine) { sh.a[i] = drx * drx + dry * dry; }
      __syncthreads();
      if (i == 0) { double nr = 0.0; for (int k = 0; k < N; ++k) nr += sh.a[k]; sh.scal[0] = sqrt(nr); }
      __syncthreads();
      double nr = sh.scal[0];
      if (nr < 1e-12) { drx /= nr; dry /= nr; dvx /= nr; dvy /= nr; }
      if (mine) { sh.b[i] = dvx * dvx + dvy * dvy; }
      __syncthreads();
      if (i == 0) { double nv = 0.0; for (int k = 0; k < N; ++k) nv += sh.b[k]; sh.scal[1] = sqrt(nv); }
      __syncthreads();
      if (nr < 1e-12) nr = 1.0;
      accum += (sh.scal[1] / nr) * tt * dt;
      __syncthreads();
    }
    megno = 2.0 * accum / tt;
    lyap = (megno == 0.0) ? inf : tt / fabs(megno);
    t_end = tt;
  }
  // ---- outputs
  const bool write = (a.flags & NB_RUN_WRITE_STATE) != 0;
  if (mine) sh.red[i] = (is_finite(s.x) && is_finite(s.y) && is_finite(s.vx) && is_finite(s.vy)) ? 0.0 : 1.0;
  if (mine && write) {
    a.q[((size_t)sys * N + i) * 2 + 0] = s.x; a.q[((size_t)sys * N + i) * 2 + 1] = s.y;
    a.v[((size_t)sys * N + i) * 2 + 0] = s.vx; a.v[((size_t)sys * N + i) * 2 + 1] = s.vy;
  }
  if (mine) sh.a[i] = (double)kep_iters;
  __syncthreads();
  if (i != 0) return;
  int st = 0;
  double iters_all = 0.0;
  for (int k = 0; k < N; ++k) { if (sh.red[k] != 0.0) st = NB_STATUS_NONFINITE; iters_all += sh.a[k]; }
  if (MODE == NB_MODE_WHFAST && kep_worst > 64) st |= NB_STATUS_KEPLER_NOCONV;
  if (a.status) a.status[sys] = st;
  if (MODE == NB_MODE_WHFAST && a.work) {
    a.work[2 * (size_t)sys] = iters_all;
    a.work[2 * (size_t)sys + 1] = 2.0 * (N - 1) * (double)n_sub * (double)(a.n_steps + a.n_megno);
  }
  if (a.dyn) {
    double* f = a.dyn + (size_t)sys * NB_N_DYN;
    const double inv = n_samp > 0 ? 1.0 / (double)n_samp : nan;
    const double ed = want_energy ? mid_drift_of(E0, E1) : nan, ld = want_energy ? mid_drift_of(L0, L1) : nan;
    f[NB_F_ENERGY_DRIFT] = ed; f[NB_F_ANGMOM_DRIFT] = ld;
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
    f[NB_F_MEGNO] = megno; f[NB_F_LYAP_TIME] = lyap; f[NB_F_T_END] = t_end;
    f[NB_F_E0] = want_energy ? E0 : nan; f[NB_F_E1] = want_energy ? E1 : nan;
    f[NB_F_L0] = want_energy ? L0 : nan; f[NB_F_L1] = want_energy ? L1 : nan;
    f[NB_F_IS_STABLE] = ((ed < 0.01) && (ld < 0.01) && (f[NB_F_COM_MEAN] < 1.0) && (megno < 10.0)) ? 1.0 : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// prepare kernel (COM removal, corrector kicks, frozen schedule, static features)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MID_THREADS) mid_prepare_kernel(PrepArgs a, int N) {
  __shared__ MidSh sh;
  __shared__ double dist[MID_MAX * (MID_MAX - 1) / 2];     // pair separations of the static features (16 KB)
  const int sys = blockIdx.x;
  const int i = threadIdx.x;
  const bool mine = i < N;
  const double G = a.G;
  MidBody s = {0, 0, 0, 0, 0, 0};
  if (mine) {
    sh.m[i] = a.m[(size_t)sys * N + i];
    sh.gm[i] = G * sh.m[i];
    s.x = a.q[((size_t)sys * N + i) * 2 + 0]; s.y = a.q[((size_t)sys * N + i) * 2 + 1];
    s.vx = a.v[((size_t)sys * N + i) * 2 + 0]; s.vy = a.v[((size_t)sys * N + i) * 2 + 1];
    sh.x[i] = s.x; sh.y[i] = s.y; sh.a[i] = s.vx; sh.b[i] = s.vy;
  }
  const double eps = a.eps[sys];
  const double eps2 = eps * eps;
  __syncthreads();
  bool touched = false;
  if (a.flags & NB_PREP_REMOVE_COM) {     // physics_utils.py:16-26
    if (i == 0) {
      double M = 0.0, px = 0.0, py = 0.0;
      for (int k = 0; k < N; ++k) { M += sh.m[k]; px += sh.m[k] * sh.a[k]; py += sh.m[k] * sh.b[k]; }
      sh.scal[0] = M; sh.scal[1] = M != 0.0 ? px / M : 0.0; sh.scal[2] = M != 0.0 ? py / M : 0.0;
    }
    __syncthreads();
    if (sh.scal[0] != 0.0) { s.vx -= sh.scal[1]; s.vy -= sh.scal[2]; }
    touched = true;
    __syncthreads();
  }
  const int n_kicks = ((a.flags & NB_PREP_CTOR_KICK) ? 1 : 0) + ((a.flags & NB_PREP_SNAPSHOT_KICK) ? 1 : 0);
  if (n_kicks > 0 && G != 0.0) {          // integration_scheme_base.py:154-192 / whfast_scheme.py:95-123
    double d1 = 0.0, d2 = 0.0;
    if (a.mode == NB_MODE_WHFAST) {
      if (i == 0) mid_wh_interaction(sh, N, G, eps2);
      __syncthreads();
      if (mine) { s.ax = sh.c[i]; s.ay = sh.d[i]; }
      __syncthreads();
    } else {
      mid_force<false>(sh, N, i, mine, s, eps2, 0.0, 0.0, d1, d2);
    }
    for (int k = 0; k < n_kicks; ++k) { s.vx = fma(0.5 * a.kick_dt, s.ax, s.vx); s.vy = fma(0.5 * a.kick_dt, s.ay, s.vy); }
    touched = true;
  }
  if (touched && mine) {
    a.v[((size_t)sys * N + i) * 2 + 0] = s.vx;
    a.v[((size_t)sys * N + i) * 2 + 1] = s.vy;
  }
  if (mine) { sh.a[i] = s.vx; sh.b[i] = s.vy; }
  __syncthreads();
  // ---- frozen sub-step schedule (timestep_manager.py:139-253): a minimum over pairs is order-independent
  if (a.h_sub_ref || a.n_sub) {
    double tau = __longlong_as_double(0x7ff0000000000000LL);
    if (G != 0.0 && mine) {
      for (int j = i + 1; j < N; ++j) {
        const double dx = s.x - sh.x[j], dy = s.y - sh.y[j];
        const double r = sqrt(dx * dx + dy * dy);
        const double r3 = r * r * r;
        const double den = G * (sh.m[i] + sh.m[j]);
        if (is_finite(r3) && is_finite(den) && den > 0.0) tau = fmin(tau, sqrt(r3 / den));
      }
    }
    sh.red[i] = tau;
    __syncthreads();
    if (i == 0) {
      for (int k = 1; k < N; ++k) tau = fmin(tau, sh.red[k]);
      const double dt_user = fabs(a.sched_dt);
      double hs = 0.9 * tau;
      if (!is_finite(hs) || hs <= 0.0) hs = dt_user > 0.0 ? dt_user : 1.0;
      if (a.split_n_max > 0) {
        if (ceil(dt_user / fmax(hs, 1e-30)) > (double)a.split_n_max) hs = dt_user / (double)a.split_n_max;
      }
      if (a.h_sub_ref) a.h_sub_ref[sys] = hs;
      if (a.n_sub) {
        const double need = ceil(fabs(a.dt) / hs);
        const int ns = need > (double)a.split_n_max ? a.split_n_max : (int)need;
        a.n_sub[sys] = max(1, ns);
      }
    }
    __syncthreads();
  }
  // ---- static features (dynamical_features.py:27-155), thread 0, the sums in the order of ensemble_prepare_kernel
  if ((a.flags & NB_PREP_STATIC_FEATURES) && a.stat && i == 0) {
    double* f = a.stat + (size_t)sys * NB_N_STATIC;
    double M = 0.0, mmin = sh.m[0], mmax = sh.m[0], xs = 0.0, ys = 0.0;
    for (int k = 0; k < N; ++k) {
      M += sh.m[k]; mmin = fmin(mmin, sh.m[k]); mmax = fmax(mmax, sh.m[k]);
      xs += sh.m[k] * sh.x[k]; ys += sh.m[k] * sh.y[k];
    }
    double mvar = 0.0;
    { const double mu = M / N; for (int k = 0; k < N; ++k) mvar += (sh.m[k] - mu) * (sh.m[k] - mu); mvar /= N; }
    f[NB_S_TOTAL_MASS] = M;
    f[NB_S_MASS_VAR] = mvar;
    f[NB_S_MASS_RATIO_MAX] = mmin > 0.0 ? mmax / mmin : 1.0;
    f[NB_S_MASS_CENTER_OFFSET] = (M != 0.0) ? sqrt((xs / M) * (xs / M) + (ys / M) * (ys / M)) : 0.0;
    const int NP = N * (N - 1) / 2;
    double dsum = 0.0, dmin = __longlong_as_double(0x7ff0000000000000LL), dmax = 0.0, rsum = 0.0, rmax = 0.0, PE = 0.0;
    int p = 0;
    for (int k = 0; k < N; ++k)
      for (int l = k + 1; l < N; ++l) {
        const double dx = sh.x[l] - sh.x[k], dy = sh.y[l] - sh.y[k];
        const double r = sqrt(dx * dx + dy * dy);
        dist[p++] = r; dsum += r; dmin = fmin(dmin, r); dmax = fmax(dmax, r);
        const double ux = sh.a[l] - sh.a[k], uy = sh.b[l] - sh.b[k];
        const double dv = sqrt(ux * ux + uy * uy);
        rsum += dv; rmax = fmax(rmax, dv);
        PE -= G * sh.m[k] * sh.m[l] / sqrt(dx * dx + dy * dy + eps2);
      }
    const double dmean = dsum / NP;
    double dvar = 0.0;
    for (int q = 0; q < NP; ++q) dvar += (dist[q] - dmean) * (dist[q] - dmean);
    f[NB_S_MEAN_SEP] = dmean;
    f[NB_S_STD_SEP] = sqrt(dvar / NP);
    f[NB_S_MIN_SEP] = dmin;
    f[NB_S_MAX_SEP] = dmax;
    f[NB_S_SEP_RATIO] = dmin > 0.0 ? dmax / dmin : 1.0;
    double ssum = 0.0, smax = 0.0, KE = 0.0, L = 0.0, spsum = 0.0;
    for (int k = 0; k < N; ++k) {
      const double v2 = sh.a[k] * sh.a[k] + sh.b[k] * sh.b[k];
      sh.c[k] = sqrt(v2); ssum += sh.c[k]; smax = fmax(smax, sh.c[k]);
      KE += 0.5 * sh.m[k] * v2;
      const double li = sh.m[k] * (sh.x[k] * sh.b[k] - sh.y[k] * sh.a[k]);
      L += li;
      sh.d[k] = fabs(li) / sh.m[k]; spsum += sh.d[k];
    }
    const double smean = ssum / N, spmean = spsum / N;
    double svar = 0.0, spvar = 0.0;
    for (int k = 0; k < N; ++k) { svar += (sh.c[k] - smean) * (sh.c[k] - smean); spvar += (sh.d[k] - spmean) * (sh.d[k] - spmean); }
    f[NB_S_MEAN_SPEED] = smean;
    f[NB_S_STD_SPEED] = sqrt(svar / N);
    f[NB_S_MAX_SPEED] = smax;
    f[NB_S_MEAN_RELVEL] = rsum / NP;
    f[NB_S_MAX_RELVEL] = rmax;
    const double E = KE + PE;
    f[NB_S_KINETIC] = KE;
    f[NB_S_POTENTIAL] = PE;
    f[NB_S_TOTAL_ENERGY] = E;
    f[NB_S_VIRIAL] = PE != 0.0 ? 2.0 * KE / fabs(PE) : 0.0;
    f[NB_S_ENERGY_PER_MASS] = E / M;
    f[NB_S_IS_BOUND] = E < 0.0 ? 1.0 : 0.0;
    f[NB_S_TOTAL_ANGMOM] = fabs(L);
    f[NB_S_MEAN_SPEC_ANGMOM] = spmean;
    f[NB_S_ANGMOM_VAR] = spvar / N;
    f[NB_S_SOFT_MEAN] = eps;
    f[NB_S_SOFT_STD] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// stand-alone pair / variational calls
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MID_THREADS) mid_pair_kernel(const double* __restrict__ q, const double* __restrict__ m,
                                                               const double* __restrict__ eps, double G, int N,
                                                               double* acc, double* U, double* dV) {
  __shared__ MidSh sh;
  const int sys = blockIdx.x, i = threadIdx.x;
  const bool mine = i < N;
  MidBody s = {0, 0, 0, 0, 0, 0};
  if (mine) {
    sh.m[i] = m[(size_t)sys * N + i];
    sh.gm[i] = G * sh.m[i];
    s.x = q[((size_t)sys * N + i) * 2 + 0]; s.y = q[((size_t)sys * N + i) * 2 + 1];
  }
  const double e = eps[sys], eps2 = e * e;
  double d1 = 0.0, d2 = 0.0;
  mid_force<false>(sh, N, i, mine, s, eps2, 0.0, 0.0, d1, d2);
  if (acc && mine) {
    acc[((size_t)sys * N + i) * 2 + 0] = s.ax;
    acc[((size_t)sys * N + i) * 2 + 1] = s.ay;
  }
  if ((U || dV) && i == 0) {       // pair_scalars of pair_small.cuh, same pair order
    double u = 0.0, s3 = 0.0;
    for (int k = 0; k < N; ++k)
      for (int l = k + 1; l < N; ++l) {
        const double dx = sh.x[k] - sh.x[l], dy = sh.y[k] - sh.y[l];
        const double w = rsqrt_f64<true>(fma(dx, dx, fma(dy, dy, eps2)));
        const double mm = sh.gm[k] * sh.m[l];
        u = fma(mm, w, u);
        s3 = fma(mm, w * w * w, s3);
      }
    if (U) U[sys] = (G == 0.0) ? 0.0 : -u;
    if (dV) dV[sys] = (e == 0.0 || G == 0.0) ? 0.0 : e * s3;
  }
}

__global__ void __launch_bounds__(MID_THREADS) mid_variational_kernel(const double* __restrict__ q, const double* __restrict__ m,
                                                                      const double* __restrict__ s2, const double* __restrict__ dr,
                                                                      double G, int N, double* da) {
  __shared__ MidSh sh;
  const int sys = blockIdx.x, i = threadIdx.x;
  const bool mine = i < N;
  MidBody s = {0, 0, 0, 0, 0, 0};
  double drx = 0.0, dry = 0.0, dax = 0.0, day = 0.0;
  if (mine) {
    sh.m[i] = m[(size_t)sys * N + i];
    sh.gm[i] = G * sh.m[i];
    s.x = q[((size_t)sys * N + i) * 2 + 0]; s.y = q[((size_t)sys * N + i) * 2 + 1];
    drx = dr[((size_t)sys * N + i) * 2 + 0]; dry = dr[((size_t)sys * N + i) * 2 + 1];
  }
  mid_force<true>(sh, N, i, mine, s, s2[sys], drx, dry, dax, day);
  if (mine) {
    da[((size_t)sys * N + i) * 2 + 0] = dax;
    da[((size_t)sys * N + i) * 2 + 1] = day;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int mid_run(const RunArgs& a, int N, int mode, cudaStream_t st) {
  if (N <= NB_MAX_N || N > MID_MAX) { set_error("mid-N kernels cover 9..64 bodies"); return NB_ERR_ARG; }
  const bool exact = mode == NB_MODE_WHFAST && (a.flags & NB_RUN_KEPLER_EXACT) != 0;
  switch (mode) {
    case NB_MODE_VERLET: mid_run_kernel<NB_MODE_VERLET, false><<<a.B, MID_THREADS, 0, st>>>(a, N); break;
    case NB_MODE_YOSHIDA4: mid_run_kernel<NB_MODE_YOSHIDA4, false><<<a.B, MID_THREADS, 0, st>>>(a, N); break;
    case NB_MODE_WHFAST:
      if (exact) mid_run_kernel<NB_MODE_WHFAST, true><<<a.B, MID_THREADS, 0, st>>>(a, N);
      else mid_run_kernel<NB_MODE_WHFAST, false><<<a.B, MID_THREADS, 0, st>>>(a, N);
      break;
    default: set_error("nb_ensemble_run_f64: systems of more than 8 bodies run verlet / yoshida4 / whfast"); return NB_ERR_UNSUPPORTED;
  }
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int mid_prepare(const PrepArgs& a, int N, cudaStream_t st) {
  if (N <= NB_MAX_N || N > MID_MAX) { set_error("mid-N kernels cover 9..64 bodies"); return NB_ERR_ARG; }
  mid_prepare_kernel<<<a.B, MID_THREADS, 0, st>>>(a, N);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int mid_pair(const double* q, const double* m, const double* eps, double G, int B, int N, double* acc, double* U, double* dV,
             cudaStream_t st) {
  mid_pair_kernel<<<B, MID_THREADS, 0, st>>>(q, m, eps, G, N, acc, U, dV);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int mid_variational(const double* q, const double* m, const double* s2, const double* dr, double G, int B, int N, double* da,
                    cudaStream_t st) {
  mid_variational_kernel<<<B, MID_THREADS, 0, st>>>(q, m, s2, dr, G, N, da);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
