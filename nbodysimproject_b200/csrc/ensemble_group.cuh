// ensemble_group.cuh -- latency-optimised mapping for systems with many sub-steps: one BODY per lane,
// floor(32/N) systems per warp, positions exchanged with warp shuffles.
//
// Why: a system's sub-steps are inherently sequential.  In the thread-per-system mapping one force evaluation
// is ~17 N(N-1)/2 dependent-issue FP64 instructions (N = 8: ~520, >= 1040 cycles on one SMSP), so a system
// that needs n_sub = 50 sub-steps per step (integrator.py:86-92 caps at split_n_max = 50) keeps ONE warp busy
// for ~100 ms while the rest of the GPU has long finished.  Here lane b evaluates only a_b = sum_j (N pair
// terms, ~14 (N-1) FP64 instructions), i.e. a ~5x shorter critical path at N = 8, at the price of evaluating every
// ordered pair (no Newton's-third-law sharing).  Used for the heavy head (nb_sort_by_nsub) of the n_sub-sorted
// permutation: CTAs [0, group_blocks) of the main / MEGNO kernels run this body, the rest run the
// throughput-optimal thread-per-system body, so both mappings overlap inside one launch.
#pragma once
#include "pair_small.cuh"
#include "args.cuh"

namespace nb {

constexpr int NB_HEAVY_NSUB = 4;   // smallest threshold: systems with n_sub <= this never leave the thread mapping

template <int N>
struct GroupCtx {
  unsigned mask;
  int base;
  double gmk[N - 1];     // G m of the t-th partner (ascending body index, self skipped)
  int src[N - 1];        // lane that holds the t-th partner
  double eps2;
};

template <int N, bool TANGENT, bool GUARD>
__device__ __forceinline__ void group_accel(const GroupCtx<N>& c, int b, double x, double y, double& ax, double& ay,
                                            double drx, double dry, double& dax, double& day) {
  ax = 0.0; ay = 0.0;
  if (TANGENT) { dax = 0.0; day = 0.0; }
  // partners in ascending body index, the self pair skipped: t-th partner of body b is src[t] = t + (t >= b)
#pragma unroll
  for (int t = 0; t < N - 1; ++t) {
    const int src = c.src[t];
    const double xj = __shfl_sync(0xffffffffu, x, src);
    const double yj = __shfl_sync(0xffffffffu, y, src);
    const double dx = x - xj, dy = y - yj;
    const double r2 = fma(dx, dx, fma(dy, dy, c.eps2));
    double w2, w3;                                   // same arithmetic as pair_small.cuh: bit-identical accelerations
    if (TANGENT) {
      const double w = rsqrt_f64<GUARD>(r2);
      w2 = w * w;
      w3 = w2 * w;
    } else {
      w2 = 0.0;
      w3 = rsqrt3_f64<GUARD>(r2);
    }
    const double cj = c.gmk[t] * w3;
    ax = fma(-cj, dx, ax);
    ay = fma(-cj, dy, ay);
    if (TANGENT) {
      const double ex = __shfl_sync(0xffffffffu, drx, src) - drx;   // d = dr_j - dr_i ; D = q_j - q_i = -(dx,dy)
      const double ey = __shfl_sync(0xffffffffu, dry, src) - dry;
      const double dot = -fma(dx, ex, dy * ey);
      const double c5 = 3.0 * dot * w2 * w3;
      dax = fma(c.gmk[t], fma(ex, w3, c5 * dx), dax);
      day = fma(c.gmk[t], fma(ey, w3, c5 * dy), day);
    }
  }
}

template <int N>
__device__ __forceinline__ double group_sum(const GroupCtx<N>& c, double v) {
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) s += __shfl_sync(0xffffffffu, v, c.base + j);
  return s;
}

// one sub-step for the lane-owned body (x, y, vx, vy, ax, ay); FSAL like the thread mapping
template <int N, int MODE, bool TANGENT, bool GUARD>
__device__ __forceinline__ void group_substep(const GroupCtx<N>& c, int b, double h, double& x, double& y, double& vx,
                                              double& vy, double& ax, double& ay, double drx, double dry, double& dax,
                                              double& day) {
  double dum1 = 0.0, dum2 = 0.0;
  if (MODE == NB_MODE_VERLET) {
    const double h2 = 0.5 * h;
    vx = fma(h2, ax, vx); vy = fma(h2, ay, vy);
    x = fma(h, vx, x); y = fma(h, vy, y);
    group_accel<N, TANGENT, GUARD>(c, b, x, y, ax, ay, drx, dry, dax, day);
    vx = fma(h2, ax, vx); vy = fma(h2, ay, vy);
  } else {
    const double cbrt2 = 1.2599210498948731648;
    const double w1 = 1.0 / (2.0 - cbrt2), w2 = -cbrt2 / (2.0 - cbrt2);
    const double ha = w1 * h, hb = w2 * h;
    const double hab = 0.5 * ha + 0.5 * hb;          // merged adjacent half kicks, exactly as ensemble_run.cuh substep<>
    vx = fma(0.5 * ha, ax, vx); vy = fma(0.5 * ha, ay, vy);
    x = fma(ha, vx, x); y = fma(ha, vy, y);
    group_accel<N, false, GUARD>(c, b, x, y, ax, ay, 0.0, 0.0, dum1, dum2);
    vx = fma(hab, ax, vx); vy = fma(hab, ay, vy);
    x = fma(hb, vx, x); y = fma(hb, vy, y);
    group_accel<N, false, GUARD>(c, b, x, y, ax, ay, 0.0, 0.0, dum1, dum2);
    vx = fma(hab, ax, vx); vy = fma(hab, ay, vy);
    x = fma(ha, vx, x); y = fma(ha, vy, y);
    group_accel<N, TANGENT, GUARD>(c, b, x, y, ax, ay, drx, dry, dax, day);
    vx = fma(0.5 * ha, ax, vx); vy = fma(0.5 * ha, ay, vy);
  }
}

// phase 0: main loop + step_metrics sampling ; phase 1: MEGNO.
template <int N, int MODE, bool GUARD>
__device__ __forceinline__ void group_body(const RunArgs& a, int phase, int write_state, int bid) {
  constexpr int G = 32 / N;
  const int lane = threadIdx.x & 31;
  const int warp = (bid * (int)blockDim.x + (int)threadIdx.x) >> 5;
  const int n_heavy = min(*a.n_heavy, a.B);
  if (warp * G >= n_heavy) return;                      // warp-uniform exit
  // The warp stays fully converged (full-mask shuffles compile to bare SHFL; partial masks cost a WARPSYNC
  // per shuffle and serialise the dependency chain): spare lanes and empty slots shadow a live system
  // and simply never write.
  int g = lane / N;
  int b = lane - g * N;
  bool live = g < G;
  if (!live) { g = G - 1; b = 0; }
  int slot = warp * G + g;
  if (slot >= n_heavy) { slot = n_heavy - 1; live = false; }
  GroupCtx<N> c;
  c.mask = 0xffffffffu;
  c.base = (slot == warp * G + g) ? g * N : ((n_heavy - 1) - warp * G) * N;   // lanes that hold this slot's bodies
  const int sys = a.perm[slot];
  const double G_ = a.G;
#pragma unroll
  for (int t = 0; t < N - 1; ++t) {
    const int k = t < b ? t : t + 1;
    c.gmk[t] = G_ * a.m[(size_t)sys * N + k];
    c.src[t] = c.base + k;
  }
  const double mb = a.m[(size_t)sys * N + b];
  double x = a.q[((size_t)sys * N + b) * 2 + 0], y = a.q[((size_t)sys * N + b) * 2 + 1];
  double vx = a.v[((size_t)sys * N + b) * 2 + 0], vy = a.v[((size_t)sys * N + b) * 2 + 1];
  const double eps = a.eps[sys];
  c.eps2 = eps * eps;
  const int n_sub = max(1, a.n_sub[sys]);
  const int n_sub_warp = __reduce_max_sync(0xffffffffu, n_sub);   // sorted by n_sub: nearly uniform per warp
  const double h = a.dt / (double)n_sub;
  double ax, ay, d1 = 0.0, d2 = 0.0;
  group_accel<N, false, GUARD>(c, b, x, y, ax, ay, 0.0, 0.0, d1, d2);
  double* f = a.dyn ? a.dyn + (size_t)sys * NB_N_DYN : nullptr;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);

  if (phase == 0) {
    double com_sum = 0.0, com_max = -1.0, var_sum = 0.0, var_max = -1.0, cos_sum = 0.0, cos_min = 2.0, th_sum = 0.0;
    double Lfirst = 0.0;
    bool have_first = false, cos_nan = false;
    int n_samp = 0, next_sample = 0;
    const int interval = a.sample_interval;
  const double theta_eps = (eps != 0.0) ? atan2(0.0, eps) : nan;   // diagnostics.py:246-249 with pi = 0: loop-invariant
    for (int step = 0; step < a.n_steps; ++step) {
#pragma unroll 1
      for (int k = 0; k < n_sub_warp; ++k) {
        double xs = x, ys = y, us = vx, ws = vy, as = ax, bs = ay;
        group_substep<N, MODE, false, GUARD>(c, b, h, xs, ys, us, ws, as, bs, 0.0, 0.0, d1, d2);
        if (k < n_sub) { x = xs; y = ys; vx = us; vy = ws; ax = as; ay = bs; }
      }
      if (interval > 0 && step == next_sample) {
        next_sample += interval;
        const double Li = mb * (x * vy - y * vx);
        const double cx = group_sum<N>(c, mb * x), cy = group_sum<N>(c, mb * y), Lt = group_sum<N>(c, Li);
        const double mean = Lt / N;
        const double var = group_sum<N>(c, (Li - mean) * (Li - mean)) / N;
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
    if (f && b == 0 && live) {
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
  } else {
    // MEGNO (evolution_features.py:34-66)
    double drx = a.raw_dr[((size_t)sys * N + b) * 2 + 0], dry = a.raw_dr[((size_t)sys * N + b) * 2 + 1];
    double dvx = a.raw_dv[((size_t)sys * N + b) * 2 + 0], dvy = a.raw_dv[((size_t)sys * N + b) * 2 + 1];
    {
      const double M = group_sum<N>(c, mb);
      const double cx = group_sum<N>(c, mb * drx) / M, cy = group_sum<N>(c, mb * dry) / M;
      const double ux = group_sum<N>(c, mb * dvx) / M, uy = group_sum<N>(c, mb * dvy) / M;
      drx -= cx; dry -= cy; dvx -= ux; dvy -= uy;
      const double nr = sqrt(group_sum<N>(c, drx * drx + dry * dry));
      const double nv = sqrt(group_sum<N>(c, dvx * dvx + dvy * dvy));
      drx /= nr; dry /= nr; dvx /= nv; dvy /= nv;
    }
    double tt = 0.0, accum = 0.0, dax = 0.0, day = 0.0;
    const double dt = a.dt;
    for (int step = 0; step < a.n_megno; ++step) {
#pragma unroll 1
      for (int k = 0; k < n_sub_warp - 1; ++k) {
        double xs = x, ys = y, us = vx, ws = vy, as = ax, bs = ay;
        group_substep<N, MODE, false, GUARD>(c, b, h, xs, ys, us, ws, as, bs, 0.0, 0.0, d1, d2);
        if (k < n_sub - 1) { x = xs; y = ys; vx = us; vy = ws; ax = as; ay = bs; }
      }
      drx = fma(dvx, dt, drx); dry = fma(dvy, dt, dry);
      group_substep<N, MODE, true, GUARD>(c, b, h, x, y, vx, vy, ax, ay, drx, dry, dax, day);
      dvx = fma(dax, dt, dvx); dvy = fma(day, dt, dvy);
      tt += dt;
      double nr = sqrt(group_sum<N>(c, drx * drx + dry * dry));
      if (nr < 1e-12) { drx /= nr; dry /= nr; dvx /= nr; dvy /= nr; nr = 1.0; }
      const double nv = sqrt(group_sum<N>(c, dvx * dvx + dvy * dvy));
      accum += (nv / nr) * tt * dt;
    }
    if (f && b == 0 && live) {
      const double megno = 2.0 * accum / tt;
      f[NB_F_MEGNO] = megno;
      f[NB_F_LYAP_TIME] = (megno == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : tt / fabs(megno);
      f[NB_F_T_END] = tt;
    }
  }
  const bool fin = is_finite(x) && is_finite(y) && is_finite(vx) && is_finite(vy);
  const unsigned grp = ((1u << N) - 1u) << (c.base);
  const unsigned bad = __ballot_sync(0xffffffffu, !fin) & grp;
  if (write_state && live) {
    a.q[((size_t)sys * N + b) * 2 + 0] = x;
    a.q[((size_t)sys * N + b) * 2 + 1] = y;
    a.v[((size_t)sys * N + b) * 2 + 0] = vx;
    a.v[((size_t)sys * N + b) * 2 + 1] = vy;
  }
  if (a.status && b == 0 && live) {
    const int st = bad ? NB_STATUS_NONFINITE : 0;
    if (phase == 0) a.status[sys] = st; else a.status[sys] |= st;
  }
}

// CTAs needed if every system were heavy (the real count is only known on the device)
template <int N>
static inline int group_blocks_for(int B) {
  constexpr int G = 32 / N;
  const int warps = (B + G - 1) / G;
  return (warps + 3) / 4;
}

}  // namespace nb
