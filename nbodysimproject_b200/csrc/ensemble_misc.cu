// ensemble_misc.cu -- construction-time kernel, stand-alone pair kernels and the n_sub counting sort.
#include <cstdlib>
#include <mutex>
#include <cstdint>
#include "ensemble_run.cuh"

namespace nb {

int ensemble_run_verlet(const RunArgs& a, int N, int phase, int write_state, cudaStream_t st);
int ensemble_run_yoshida4(const RunArgs& a, int N, int phase, int write_state, cudaStream_t st);
int ensemble_run_whfast(const RunArgs& a, int N, int phase, int write_state, cudaStream_t st);

// T + U with double-double accumulation; each part rounded to fp64 and then added, like
// diagnostics.py:543-549 does with its long-double Kahan sums.  One thread per system, straight from
// global memory (runs twice per analysis, so it is kept out of the register-resident hot kernels).
__global__ void __launch_bounds__(128) energy_kernel(const double* __restrict__ m, const double* __restrict__ q,
                                                     const double* __restrict__ v, const double* __restrict__ eps_arr,
                                                     double G, int B, int N, double* dyn, int slot) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= B) return;
  const double* mm = m + (size_t)sys * N;
  const double* qq = q + (size_t)sys * N * 2;
  const double* vv = v + (size_t)sys * N * 2;
  const double eps = eps_arr[sys];
  dd T = dd_make(0.0);
  double L = 0.0;
  for (int i = 0; i < N; ++i) {
    const double vx = vv[2 * i], vy = vv[2 * i + 1];
    dd v2 = dd_add(two_prod(vx, vx), two_prod(vy, vy));
    T = dd_add(T, dd_mul_d(dd_mul_d(v2, mm[i]), 0.5));
    L += mm[i] * __dadd_rn(__dmul_rn(qq[2 * i], vy), -__dmul_rn(qq[2 * i + 1], vx));   // diagnostics.py:553-557
  }
  dd S = dd_make(0.0);
  const dd e2 = two_prod(eps, eps);
  if (G != 0.0) {
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) {
        dd dx = two_sum(qq[2 * i], -qq[2 * j]);
        dd dy = two_sum(qq[2 * i + 1], -qq[2 * j + 1]);
        dd r2 = dd_add(dd_add(dd_mul(dx, dx), dd_mul(dy, dy)), e2);
        if (!(r2.hi > 0.0)) r2 = dd_make(1e-300);
        dd inv = dd_div(dd_make(1.0), dd_sqrt(r2));
        S = dd_add(S, dd_mul(two_prod(mm[i], mm[j]), inv));
      }
  }
  const double Tf = dd_to_double(T);
  const double Vf = dd_to_double(dd_mul_d(S, -G));
  double* f = dyn + (size_t)sys * NB_N_DYN;
  f[NB_F_E0 + slot] = Tf + Vf;
  f[NB_F_L0 + slot] = L;
}

// drifts + the is_stable predicate (stability_analyzer.py:147-231)
__global__ void __launch_bounds__(128) finalize_kernel(double* dyn, int B, int have_energy, int have_megno) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= B) return;
  double* f = dyn + (size_t)sys * NB_N_DYN;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (!have_megno) {
    f[NB_F_MEGNO] = 2.0;
    f[NB_F_LYAP_TIME] = __longlong_as_double(0x7ff0000000000000LL);
    f[NB_F_T_END] = 0.0;
  }
  double ed = nan, ld = nan;
  if (have_energy) {
    ed = drift_of(f[NB_F_E0], f[NB_F_E1]);
    ld = drift_of(f[NB_F_L0], f[NB_F_L1]);
  } else {
    f[NB_F_E0] = nan; f[NB_F_E1] = nan; f[NB_F_L0] = nan; f[NB_F_L1] = nan;
  }
  f[NB_F_ENERGY_DRIFT] = ed;
  f[NB_F_ANGMOM_DRIFT] = ld;
  f[NB_F_IS_STABLE] = ((ed < 0.01) && (ld < 0.01) && (f[NB_F_COM_MEAN] < 1.0) && (f[NB_F_MEGNO] < 10.0)) ? 1.0 : 0.0;
}

// launchers for the other translation units (ensemble_adaptive.cu)
int launch_energy(const double* m, const double* q, const double* v, const double* eps, double G, int B, int N, double* dyn,
                  int slot, cudaStream_t st) {
  energy_kernel<<<(B + 127) / 128, 128, 0, st>>>(m, q, v, eps, G, B, N, dyn, slot);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}
int launch_finalize(double* dyn, int B, int have_energy, int have_megno, cudaStream_t st) {
  finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(dyn, B, have_energy, have_megno);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

// ---------------------------------------------------------------------------------------------
// Heads first.  Each bucket's launch ends in a latency-bound tail: the sub-step-heavy systems at the head of the
// n_sub-sorted permutation run long sequential chains while the bulk is throughput-bound.  When several buckets are in
// flight (one stream each) the block scheduler works through the kernels roughly in launch order, so the chains of
// the buckets launched later START late and finish last (measured: 54.5 ms for the six C3 main kernels, 48.5 ms when
// every head starts first).  The main kernel is therefore split into a HEAD launch (the latency-mapped CTAs plus the
// first 5 % of the thread-mapped ones) and a REST launch, and the head -- together with the small energy kernel that
// must precede it -- runs on a stream of HIGHER priority than the rest: its CTAs are dispatched ahead of every queued
// bulk CTA of every bucket.  If the caller's stream already has the highest priority (the *_host slots) the rest is
// demoted to an internal normal-priority stream, otherwise the head is promoted to an internal high-priority one.
// Internal streams / events are created once per device (the only allocation the device-pointer entry points make).
// ---------------------------------------------------------------------------------------------
constexpr int NB_SIDE = 32;   // distinct caller streams served without sharing a side stream (more: hashed)
struct SidePool {
  bool init = false;
  int prio_hi = 0, prio_lo = 0;
  cudaStream_t hi[NB_SIDE], lo[NB_SIDE];
  cudaEvent_t ev[NB_SIDE][3];
  cudaStream_t owner[NB_SIDE];
  int n_owner = 0;
};
static SidePool g_side[16];
static std::mutex g_side_mu;
constexpr int NB_SPLIT_MIN_B = 4096;   // below this a bucket is one launch (nothing to overlap)

static int side_pool(int dev, cudaStream_t st, SidePool** out, int* k) {
  SidePool& p = g_side[dev & 15];
  if (!p.init) {
    NB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&p.prio_lo, &p.prio_hi));
    for (int i = 0; i < NB_SIDE; ++i) {
      NB_CUDA_CHECK(cudaStreamCreateWithPriority(&p.hi[i], cudaStreamNonBlocking, p.prio_hi));
      NB_CUDA_CHECK(cudaStreamCreateWithPriority(&p.lo[i], cudaStreamNonBlocking, p.prio_lo));
      for (int e = 0; e < 3; ++e) NB_CUDA_CHECK(cudaEventCreateWithFlags(&p.ev[i][e], cudaEventDisableTiming));
    }
    p.init = true;
  }
  int idx = -1;
  for (int i = 0; i < p.n_owner; ++i) if (p.owner[i] == st) idx = i;
  if (idx < 0) {
    idx = p.n_owner < NB_SIDE ? p.n_owner++ : (int)(((uintptr_t)st >> 6) % NB_SIDE);
    p.owner[idx] = st;
  }
  *out = &p;
  *k = idx;
  return NB_OK;
}

int ensemble_run_classic(const RunArgs& a_in, int N, int mode, cudaStream_t st) {
  if (N < NB_MIN_N || N > NB_MAX_N) { set_error("N must be in 2..8"); return NB_ERR_ARG; }
  RunArgs a = a_in;
  auto phase = [&](int ph, int write_state, const RunArgs& ra, cudaStream_t s) -> int {
    switch (mode) {
      case NB_MODE_VERLET: return ensemble_run_verlet(ra, N, ph, write_state, s);
      case NB_MODE_YOSHIDA4: return ensemble_run_yoshida4(ra, N, ph, write_state, s);
      case NB_MODE_WHFAST: return ensemble_run_whfast(ra, N, ph, write_state, s);
      default: set_error("nb_ensemble_run_f64: unsupported mode"); return NB_ERR_UNSUPPORTED;
    }
  };
  const int threads = 128, blocks = (a.B + threads - 1) / threads;
  const bool energy = (a.flags & NB_RUN_ENERGY) != 0 && a.dyn != nullptr;
  const bool megno = a.n_megno > 0;
  const int write = ((a.flags & NB_RUN_WRITE_STATE) || energy || megno) ? 1 : 0;
  const bool split = a.perm && a.n_heavy && a.B >= NB_SPLIT_MIN_B && a.n_steps > 0 &&
                     (mode == NB_MODE_VERLET || mode == NB_MODE_YOSHIDA4);
  int rc;
  NvtxRange r_all("nb_ensemble_run: E0 + main");
  if (!split) {
    if (energy) energy_kernel<<<blocks, threads, 0, st>>>(a.m, a.q, a.v, a.eps, a.G, a.B, N, a.dyn, 0);
    rc = phase(0, write, a, st);
    if (rc != NB_OK) return rc;
  } else {
    int total = 0, prefix = 0;
    if (mode == NB_MODE_YOSHIDA4) main_blocks_n<NB_MODE_YOSHIDA4>(a, N, &total, &prefix);
    else main_blocks_n<NB_MODE_VERLET>(a, N, &total, &prefix);
    int head = prefix + (blocks + 19) / 20;
    if (head > total) head = total;
    int dev = 0;
    NB_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_side_mu);
    SidePool* sp = nullptr;
    int k = 0;
    rc = side_pool(dev, st, &sp, &k);
    if (rc != NB_OK) return rc;
    int prio = 0;
    NB_CUDA_CHECK(cudaStreamGetPriority(st, &prio));
    const bool st_is_high = prio <= sp->prio_hi && sp->prio_hi != sp->prio_lo;
    cudaStream_t s_head = st_is_high ? st : sp->hi[k];
    cudaStream_t s_rest = st_is_high ? sp->lo[k] : st;
    cudaStream_t side = st_is_high ? s_rest : s_head;
    NB_CUDA_CHECK(cudaEventRecord(sp->ev[k][0], st));              // fork: the side stream joins the caller's order
    NB_CUDA_CHECK(cudaStreamWaitEvent(side, sp->ev[k][0], 0));
    if (energy) energy_kernel<<<blocks, threads, 0, s_head>>>(a.m, a.q, a.v, a.eps, a.G, a.B, N, a.dyn, 0);
    NB_CUDA_CHECK(cudaEventRecord(sp->ev[k][1], s_head));          // E0 is read before anything advances the state
    NB_CUDA_CHECK(cudaStreamWaitEvent(s_rest, sp->ev[k][1], 0));
    RunArgs h = a, r = a;
    h.block0 = 0; h.block_count = head;
    r.block0 = head; r.block_count = total - head;
    rc = phase(0, write, h, s_head);
    if (rc != NB_OK) return rc;
    if (r.block_count > 0) {
      rc = phase(0, write, r, s_rest);
      if (rc != NB_OK) return rc;
    }
    NB_CUDA_CHECK(cudaEventRecord(sp->ev[k][2], side));            // join
    NB_CUDA_CHECK(cudaStreamWaitEvent(st, sp->ev[k][2], 0));
  }
  NvtxRange r_tail("nb_ensemble_run: E1 + MEGNO + finalize");
  if (energy) energy_kernel<<<blocks, threads, 0, st>>>(a.m, a.q, a.v, a.eps, a.G, a.B, N, a.dyn, 1);
  if (megno) {
    rc = phase(1, (a.flags & NB_RUN_WRITE_STATE) ? 1 : 0, a, st);
    if (rc != NB_OK) return rc;
  }
  if (a.dyn) finalize_kernel<<<blocks, threads, 0, st>>>(a.dyn, a.B, energy ? 1 : 0, megno ? 1 : 0);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

// ---------------------------------------------------------------------------------------------
// prepare kernel
// ---------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) ensemble_prepare_kernel(PrepArgs a) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= a.B) return;
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
  bool touched = false;
  if (a.flags & NB_PREP_REMOVE_COM) {   // physics_utils.py:16-26
    double M = 0.0, px = 0.0, py = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) { M += m[i]; px += m[i] * s.vx[i]; py += m[i] * s.vy[i]; }
    if (M != 0.0) {
      px /= M; py /= M;
#pragma unroll
      for (int i = 0; i < N; ++i) { s.vx[i] -= px; s.vy[i] -= py; }
    }
    touched = true;
  }
  const int n_kicks = ((a.flags & NB_PREP_CTOR_KICK) ? 1 : 0) + ((a.flags & NB_PREP_SNAPSHOT_KICK) ? 1 : 0);
  if (n_kicks > 0 && G != 0.0) {        // integration_scheme_base.py:154-192 / whfast_scheme.py:95-123
    if (a.mode == NB_MODE_WHFAST) {
      double ax[N], ay[N];
      wh_interaction_accel<N>(s, m, G, ax, ay);
#pragma unroll
      for (int i = 0; i < N; ++i) { s.ax[i] = ax[i]; s.ay[i] = ay[i]; }
    } else {
      pair_pass<N, false, true>(s, nullptr, nullptr, nullptr, nullptr);
    }
    for (int k = 0; k < n_kicks; ++k) kick<N>(s, 0.5 * a.kick_dt);
    touched = true;
  }
  if (touched) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      a.v[((size_t)sys * N + i) * 2 + 0] = s.vx[i];
      a.v[((size_t)sys * N + i) * 2 + 1] = s.vy[i];
    }
  }
  // ---- frozen sub-step schedule (timestep_manager.py:139-253, classic branch)
  if (a.h_sub_ref || a.n_sub) {
    double tau = __longlong_as_double(0x7ff0000000000000LL);
    if (G != 0.0) {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j) {
          const double dx = s.x[i] - s.x[j], dy = s.y[i] - s.y[j];
          const double r = sqrt(dx * dx + dy * dy);
          const double r3 = r * r * r;
          const double den = G * (m[i] + m[j]);
          if (is_finite(r3) && is_finite(den) && den > 0.0) tau = fmin(tau, sqrt(r3 / den));
        }
    }
    const double dt_user = fabs(a.sched_dt);
    double hs = 0.9 * tau;
    if (!is_finite(hs) || hs <= 0.0) hs = dt_user > 0.0 ? dt_user : 1.0;
    if (a.split_n_max > 0) {
      if (ceil(dt_user / fmax(hs, 1e-30)) > (double)a.split_n_max) hs = dt_user / (double)a.split_n_max;
    }
    if (a.h_sub_ref) a.h_sub_ref[sys] = hs;
    if (a.n_sub) {
      double need = ceil(fabs(a.dt) / hs);
      int ns = need > (double)a.split_n_max ? a.split_n_max : (int)need;
      a.n_sub[sys] = max(1, ns);
    }
  }
  // ---- static features (dynamical_features.py:27-155)
  if ((a.flags & NB_PREP_STATIC_FEATURES) && a.stat) {
    double* f = a.stat + (size_t)sys * NB_N_STATIC;
    double M = 0.0, mmin = m[0], mmax = m[0], xs = 0.0, ys = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      M += m[i]; mmin = fmin(mmin, m[i]); mmax = fmax(mmax, m[i]);
      xs += m[i] * s.x[i]; ys += m[i] * s.y[i];
    }
    double mvar = 0.0;
    { const double mu = M / N;
#pragma unroll
      for (int i = 0; i < N; ++i) mvar += (m[i] - mu) * (m[i] - mu);
      mvar /= N; }
    f[NB_S_TOTAL_MASS] = M;
    f[NB_S_MASS_VAR] = mvar;
    f[NB_S_MASS_RATIO_MAX] = mmin > 0.0 ? mmax / mmin : 1.0;
    f[NB_S_MASS_CENTER_OFFSET] = (M != 0.0) ? sqrt((xs / M) * (xs / M) + (ys / M) * (ys / M)) : 0.0;
    constexpr int NP = N * (N - 1) / 2;
    double dsum = 0.0, dmin = __longlong_as_double(0x7ff0000000000000LL), dmax = 0.0, rsum = 0.0, rmax = 0.0;
    double dist[NP];
    double PE = 0.0;
    { int p = 0;
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j) {
          const double dx = s.x[j] - s.x[i], dy = s.y[j] - s.y[i];
          const double r = sqrt(dx * dx + dy * dy);
          dist[p++] = r; dsum += r; dmin = fmin(dmin, r); dmax = fmax(dmax, r);
          const double ux = s.vx[j] - s.vx[i], uy = s.vy[j] - s.vy[i];
          const double dv = sqrt(ux * ux + uy * uy);
          rsum += dv; rmax = fmax(rmax, dv);
          PE -= G * m[i] * m[j] / sqrt(dx * dx + dy * dy + s.eps2);
        } }
    const double dmean = dsum / NP;
    double dvar = 0.0;
#pragma unroll
    for (int p = 0; p < NP; ++p) dvar += (dist[p] - dmean) * (dist[p] - dmean);
    f[NB_S_MEAN_SEP] = dmean;
    f[NB_S_STD_SEP] = sqrt(dvar / NP);
    f[NB_S_MIN_SEP] = dmin;
    f[NB_S_MAX_SEP] = dmax;
    f[NB_S_SEP_RATIO] = dmin > 0.0 ? dmax / dmin : 1.0;
    double sp[N], ssum = 0.0, smax = 0.0, KE = 0.0, L = 0.0, spec[N], spsum = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double v2 = s.vx[i] * s.vx[i] + s.vy[i] * s.vy[i];
      sp[i] = sqrt(v2); ssum += sp[i]; smax = fmax(smax, sp[i]);
      KE += 0.5 * m[i] * v2;
      const double li = m[i] * (s.x[i] * s.vy[i] - s.y[i] * s.vx[i]);
      L += li;
      spec[i] = fabs(li) / m[i]; spsum += spec[i];
    }
    const double smean = ssum / N, spmean = spsum / N;
    double svar = 0.0, spvar = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) { svar += (sp[i] - smean) * (sp[i] - smean); spvar += (spec[i] - spmean) * (spec[i] - spmean); }
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
    f[NB_S_SOFT_MEAN] = eps;     // history == [s] for a sim that has not been stepped; the host overrides otherwise
    f[NB_S_SOFT_STD] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// pair / variational batched kernels (a1-a6 as stand-alone calls)
// ---------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) pair_batched_kernel(const double* __restrict__ q, const double* __restrict__ m,
                                                           const double* __restrict__ eps, double G, int B,
                                                           double* acc, double* U, double* dV) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= B) return;
  SysState<N> s;
  double mm[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    mm[i] = m[(size_t)sys * N + i];
    s.gm[i] = G * mm[i];
    s.x[i] = q[((size_t)sys * N + i) * 2 + 0];
    s.y[i] = q[((size_t)sys * N + i) * 2 + 1];
  }
  const double e = eps[sys];
  s.eps2 = e * e;
  if (acc) {
    pair_pass<N, false, true>(s, nullptr, nullptr, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      acc[((size_t)sys * N + i) * 2 + 0] = s.ax[i];
      acc[((size_t)sys * N + i) * 2 + 1] = s.ay[i];
    }
  }
  if (U || dV) {
    double u, s3;
    pair_scalars<N, true>(s.gm, mm, s.x, s.y, s.eps2, u, s3);
    if (U) U[sys] = (G == 0.0) ? 0.0 : u;
    if (dV) dV[sys] = (e == 0.0 || G == 0.0) ? 0.0 : e * s3;   // forces.py:96-98
  }
}

template <int N>
__global__ void __launch_bounds__(128) variational_batched_kernel(const double* __restrict__ q, const double* __restrict__ m,
                                                                  const double* __restrict__ s2, const double* __restrict__ dr,
                                                                  double G, int B, double* da) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= B) return;
  SysState<N> s;
  double drx[N], dry[N], dax[N], day[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.gm[i] = G * m[(size_t)sys * N + i];
    s.x[i] = q[((size_t)sys * N + i) * 2 + 0];
    s.y[i] = q[((size_t)sys * N + i) * 2 + 1];
    drx[i] = dr[((size_t)sys * N + i) * 2 + 0];
    dry[i] = dr[((size_t)sys * N + i) * 2 + 1];
  }
  s.eps2 = s2[sys];
  pair_pass<N, true, true>(s, drx, dry, dax, day);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    da[((size_t)sys * N + i) * 2 + 0] = dax[i];
    da[((size_t)sys * N + i) * 2 + 1] = day[i];
  }
}

// ---------------------------------------------------------------------------------------------
// counting sort by n_sub (descending): histogram -> exclusive scan (65 bins) -> scatter
// ---------------------------------------------------------------------------------------------
__global__ void sort_hist_kernel(const int32_t* __restrict__ n_sub, int B, int32_t* bins) {
  __shared__ int sh[64];
  if (threadIdx.x < 64) sh[threadIdx.x] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x)
    atomicAdd(&sh[min(max(n_sub[i], 0), 63)], 1);
  __syncthreads();
  if (threadIdx.x < 64 && sh[threadIdx.x]) atomicAdd(&bins[threadIdx.x], sh[threadIdx.x]);
}
// Which systems leave the thread-per-system mapping?  The latency-optimised mappings shorten the sequential
// chain of a sub-step-heavy system (by the factor `speedup`, measured per N) but spend 2-3.5x the FP64-pipe time per
// force evaluation, so only systems whose own chain would otherwise set the run time should use them.  A launch
// cannot finish before a system at the split_n_max cap (50 sub-steps per step) does on the fast mapping, which is as
// long as 50 / speedup thread-mapped sub-steps: nothing below that gains anything.  The threshold is a function of N
// alone -- NOT of the batch -- so that a system is integrated by the same arithmetic however the ensemble is
// sharded over GPUs or batches (the mappings agree to rounding, not to the bit).
__global__ void sort_scan_kernel(int32_t* bins, float speedup, int fixed_thr) {
  if (threadIdx.x == 0) {                          // descending: bin 63 first
    int thr = fixed_thr;
    if (thr < 0) thr = min(max((int)floorf(50.0f / speedup), NB_HEAVY_NSUB), 63);
    int run = 0;
    for (int b = 63; b >= 0; --b) {
      if (b == thr) bins[64] = run;                // number of systems with n_sub > thr (the heavy head of perm)
      int c = bins[b]; bins[b] = run; run += c;
    }
    bins[65] = thr;
  }
}
__global__ void sort_scatter_kernel(const int32_t* __restrict__ n_sub, int B, int32_t* bins, int32_t* perm) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
    const int b = min(max(n_sub[i], 0), 63);
    perm[atomicAdd(&bins[b], 1)] = i;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int ensemble_prepare(const PrepArgs& a, int N, cudaStream_t st) {
  NvtxRange r("nb_ensemble_prepare");
  const int threads = 128, blocks = (a.B + threads - 1) / threads;
  switch (N) {
    case 2: ensemble_prepare_kernel<2><<<blocks, threads, 0, st>>>(a); break;
    case 3: ensemble_prepare_kernel<3><<<blocks, threads, 0, st>>>(a); break;
    case 4: ensemble_prepare_kernel<4><<<blocks, threads, 0, st>>>(a); break;
    case 5: ensemble_prepare_kernel<5><<<blocks, threads, 0, st>>>(a); break;
    case 6: ensemble_prepare_kernel<6><<<blocks, threads, 0, st>>>(a); break;
    case 7: ensemble_prepare_kernel<7><<<blocks, threads, 0, st>>>(a); break;
    case 8: ensemble_prepare_kernel<8><<<blocks, threads, 0, st>>>(a); break;
    default: set_error("N must be in 2..8"); return NB_ERR_ARG;
  }
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int pair_batched(const double* q, const double* m, const double* eps, double G, int B, int N, double* acc, double* U,
                 double* dV, cudaStream_t st) {
  const int threads = 128, blocks = (B + threads - 1) / threads;
  switch (N) {
    case 2: pair_batched_kernel<2><<<blocks, threads, 0, st>>>(q, m, eps, G, B, acc, U, dV); break;
    case 3: pair_batched_kernel<3><<<blocks, threads, 0, st>>>(q, m, eps, G, B, acc, U, dV); break;
    case 4: pair_batched_kernel<4><<<blocks, threads, 0, st>>>(q, m, eps, G, B, acc, U, dV); break;
    case 5: pair_batched_kernel<5><<<blocks, threads, 0, st>>>(q, m, eps, G, B, acc, U, dV); break;
    case 6: pair_batched_kernel<6><<<blocks, threads, 0, st>>>(q, m, eps, G, B, acc, U, dV); break;
    case 7: pair_batched_kernel<7><<<blocks, threads, 0, st>>>(q, m, eps, G, B, acc, U, dV); break;
    case 8: pair_batched_kernel<8><<<blocks, threads, 0, st>>>(q, m, eps, G, B, acc, U, dV); break;
    default: set_error("N must be in 2..8"); return NB_ERR_ARG;
  }
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int variational_batched(const double* q, const double* m, const double* s2, const double* dr, double G, int B, int N,
                        double* da, cudaStream_t st) {
  const int threads = 128, blocks = (B + threads - 1) / threads;
  switch (N) {
    case 2: variational_batched_kernel<2><<<blocks, threads, 0, st>>>(q, m, s2, dr, G, B, da); break;
    case 3: variational_batched_kernel<3><<<blocks, threads, 0, st>>>(q, m, s2, dr, G, B, da); break;
    case 4: variational_batched_kernel<4><<<blocks, threads, 0, st>>>(q, m, s2, dr, G, B, da); break;
    case 5: variational_batched_kernel<5><<<blocks, threads, 0, st>>>(q, m, s2, dr, G, B, da); break;
    case 6: variational_batched_kernel<6><<<blocks, threads, 0, st>>>(q, m, s2, dr, G, B, da); break;
    case 7: variational_batched_kernel<7><<<blocks, threads, 0, st>>>(q, m, s2, dr, G, B, da); break;
    case 8: variational_batched_kernel<8><<<blocks, threads, 0, st>>>(q, m, s2, dr, G, B, da); break;
    default: set_error("N must be in 2..8"); return NB_ERR_ARG;
  }
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int sort_by_nsub(const int32_t* n_sub, int B, int N, int32_t* perm, int32_t* ws, int heavy_threshold, cudaStream_t st) {
  NvtxRange r("nb_sort_by_nsub");
  if (heavy_threshold < -1 || heavy_threshold > 63) { set_error("nb_sort_by_nsub: heavy_threshold must be -1 (automatic) or 0..63"); return NB_ERR_ARG; }
  NB_CUDA_CHECK(cudaMemsetAsync(ws, 0, 66 * sizeof(int32_t), st));
  const int threads = 256;
  const int blocks = min((B + threads - 1) / threads, 148 * 8);
  sort_hist_kernel<<<blocks, threads, 0, st>>>(n_sub, B, ws);
  // measured on B200 (tools/check_thr.py): chain time of an n_sub = 50 system, thread mapping / fast mapping
  static const float speedup[9] = {1.f, 1.f, 1.f, 1.f, 1.8f, 2.1f, 3.0f, 3.5f, 4.5f};
  const int n = (N >= 2 && N <= 8) ? N : 0;
  sort_scan_kernel<<<1, 32, 0, st>>>(ws, speedup[n], heavy_threshold);
  sort_scatter_kernel<<<blocks, threads, 0, st>>>(n_sub, B, ws, perm);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
