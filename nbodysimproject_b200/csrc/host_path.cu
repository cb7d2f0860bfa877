// host_path.cu -- the *_host entry points: BatchStabilityAnalyzer.analyze_batch (batch_stability_analyzer.py:62-80) for one
// (N, mode) bucket with HOST buffers in and out, as ONE C-ABI call.
//
// A call flows through a three-stage pipeline on the slot's own streams
//     s_in : H2D of the inputs
//     s_k  : prepare -> (ham_soft: calibrate / freeze) -> sort -> run            (two compute streams, alternating chunks)
//     s_out: D2H of what is final after prepare (kicked v, static features, n_sub), then of the dynamic features
// and NB_HOST_SLOTS independent slots let a caller keep several buckets -- and several consecutive batches -- in flight,
// which is where the overlap of copies and kernels comes from (bench.py keeps two steps in flight: step k+1's inputs and
// step k-1's results move while step k computes).  opts->n_chunks > 1 additionally cuts ONE bucket into chunks that
// pipeline through the three stages.  That is off by default: every chunk of a classic-mode bucket ends in the same
// latency-bound tail (the sequential chains of its n_sub ~ 50 systems, ~30 ms at 1000 steps whatever the chunk size), so
// chunking the C3 buckets multiplies tails (measured: e2e step 57 -> 82 ms with ~96k-system chunks); it pays for
// buckets without such tails (whfast, ham_soft) and for very large ones.  The sort threshold that sends sub-step-heavy
// systems to the latency mappings depends on N only (ensemble_misc.cu), so a system is integrated by the same arithmetic
// whatever chunk it lands in: chunked == unchunked bit for bit (tests/test_gpu_host_path.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "common.cuh"
#include "args.cuh"

namespace nb {

int ensemble_run_classic(const RunArgs& a, int N, int mode, cudaStream_t st);
int ensemble_prepare(const PrepArgs& a, int N, cudaStream_t st);
int sort_by_nsub(const int32_t* n_sub, int B, int N, int32_t* perm, int32_t* ws, int heavy_threshold, cudaStream_t st);
int ensemble_analyze_adaptive(const double* m, double* q, double* v, double* eps, const double* eps_energy,
                              const double* soft_par, double G, int B, int N, int mode, double dt, int n_steps,
                              int sample_interval, int n_megno, const int32_t* n_sub, const double* raw_dr,
                              const double* raw_dv, double k_wall, int n_exp, double* e_delta, double* dyn, int32_t* status,
                              cudaStream_t st);
int hamsoft_run(const double* m, double* q, double* v, double G, int B, int N, unsigned flags, double dt, int n_steps,
                int sample_interval, int n_megno, const int32_t* n_sub, const int32_t* perm, const double* raw_dr,
                const double* raw_dv, double* eps_pi, const double* hs, double* dyn, int32_t* status, double* work,
                unsigned long long* tstamp, cudaStream_t st);
int hamsoft_setup(const double* m, const double* q, double G, int B, int N, unsigned flags, double dt, double* hs,
                  double* eps_pi, int32_t* n_sub, cudaStream_t st);
int generate_tangent(int N, int B, uint64_t seed, uint64_t first, double* dr, double* dv, cudaStream_t st);
int mid_run(const RunArgs& a, int N, int mode, cudaStream_t st);
int mid_prepare(const PrepArgs& a, int N, cudaStream_t st);

// RAII: entry points that take a `device` argument leave the caller's current device as they found it
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// ---- small device helpers of the host path ---------------------------------------------------------------------
// the 17 user-visible dynamic columns (stability_analyzer.py:226-252) without the E0/E1/L0/L1/t_end taps
__global__ void compact_dyn_kernel(const double* __restrict__ dyn, double* __restrict__ out, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * NB_N_DYN_USER) return;
  const int s = i / NB_N_DYN_USER, c = i - s * NB_N_DYN_USER;
  out[i] = dyn[(size_t)s * NB_N_DYN + c];
}

// SimConfig defaults of a ham_soft simulation (sim_config.py:27-57, simulation.py:88-114,
// hamiltonian_softening_integrator.py:47-141) from the per-system softening; eps_pi = (max(s0, eps_min), 0)
__global__ void hs_defaults_kernel(const double* __restrict__ soft, int B, double* hs, double* eps_pi, int fill_eps_pi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  double s = soft[i];
  double mn = 0.0;
  if (s < 0.0) s = mn;
  if (s > 0.0) mn = 0.1 * s;
  const double s0 = fmax(s, mn);
  double* p = hs + (size_t)i * NB_HS_NPARAM;
  p[NB_HS_K_SOFT] = 1.0e3; p[NB_HS_MU_SOFT] = 1.0; p[NB_HS_EPS_MIN] = mn; p[NB_HS_EPS_MAX] = 10.0 * s0;
  p[NB_HS_ALPHA_RUN] = 0.1; p[NB_HS_K_WALL] = 1.0e9; p[NB_HS_BARRIER_N] = 5.0; p[NB_HS_ETA] = 1.35;
  p[NB_HS_J_MAX_CAP] = 0.02; p[NB_HS_LAMBDA] = 0.3; p[NB_HS_POLICY] = 0.0; p[NB_HS_THETA_IMP] = 0.5;
  p[NB_HS_THETA_CAP] = 0.1; p[NB_HS_CHI_PI] = 0.2; p[NB_HS_OMEGA_SPR0] = 0.0; p[NB_HS_S0] = s0;
  p[NB_HS_FLAGS] = 0.0;
  if (fill_eps_pi) { eps_pi[2 * (size_t)i] = fmax(s0, mn); eps_pi[2 * (size_t)i + 1] = 0.0; }
}

// hamiltonian_softening_integrator.py:232-242 on the parameter table (the run kernel applies the same floor internally;
// the re-frozen schedule must see it too)
__global__ void hs_bump_mu_kernel(double* hs, int B, double dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  double* p = hs + (size_t)i * NB_HS_NPARAM;
  const double k = p[NB_HS_K_SOFT];
  if (is_finite(k) && k > 0.0) {
    const double r = fabs(dt) / p[NB_HS_THETA_IMP];
    const double mu = k * r * r;
    if (p[NB_HS_MU_SOFT] < mu) p[NB_HS_MU_SOFT] = mu;
  }
}

// ham_soft: the static features see the CALIBRATED epsilon (manager.step_s2 after update_continuous, simulation.py:116-117)
__global__ void hs_eps_gather_kernel(const double* __restrict__ eps_pi, double* __restrict__ eps, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) eps[i] = eps_pi[2 * (size_t)i];
}

// ---- per-slot workspace ----------------------------------------------------------------------------------------
constexpr int NB_HOST_SLOTS = 16;
constexpr int NB_MAX_CHUNKS = 32;
struct HostWs {
  int device = -1;
  cudaStream_t s_in = nullptr, s_out = nullptr, s_k[2] = {nullptr, nullptr};
  cudaEvent_t in_done[NB_MAX_CHUNKS], prep_done[NB_MAX_CHUNKS], run_done[NB_MAX_CHUNKS];
  cudaEvent_t idle = nullptr;            // everything of the previous call on this slot has left the workspace
  void* buf = nullptr;
  size_t cap = 0;
  bool ev_init = false;
};
static HostWs g_ws[NB_HOST_SLOTS];
static std::mutex g_ws_mu;

static void ws_release(HostWs& w) {
  if (w.device < 0) return;
  cudaSetDevice(w.device);
  if (w.buf) cudaFree(w.buf);
  if (w.s_in) cudaStreamDestroy(w.s_in);
  if (w.s_out) cudaStreamDestroy(w.s_out);
  for (int k = 0; k < 2; ++k) if (w.s_k[k]) cudaStreamDestroy(w.s_k[k]);
  if (w.ev_init) {
    for (int c = 0; c < NB_MAX_CHUNKS; ++c) {
      cudaEventDestroy(w.in_done[c]); cudaEventDestroy(w.prep_done[c]); cudaEventDestroy(w.run_done[c]);
    }
    cudaEventDestroy(w.idle);
  }
  w = HostWs();
}

static int ws_reserve(int slot, int device, size_t bytes) {
  HostWs& w = g_ws[slot];
  if (w.device != device) {
    ws_release(w);
    // compute streams at high priority: everything of a bucket except the bulk of its main kernel (which
    // ensemble_run_classic moves to a normal-priority side stream) is dispatched ahead of other buckets' queued bulk
    int prio_lo = 0, prio_hi = 0;
    NB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int k = 0; k < 2; ++k) NB_CUDA_CHECK(cudaStreamCreateWithPriority(&w.s_k[k], cudaStreamNonBlocking, prio_hi));
    NB_CUDA_CHECK(cudaStreamCreateWithFlags(&w.s_in, cudaStreamNonBlocking));
    NB_CUDA_CHECK(cudaStreamCreateWithFlags(&w.s_out, cudaStreamNonBlocking));
    for (int c = 0; c < NB_MAX_CHUNKS; ++c) {
      NB_CUDA_CHECK(cudaEventCreateWithFlags(&w.in_done[c], cudaEventDisableTiming));
      NB_CUDA_CHECK(cudaEventCreateWithFlags(&w.prep_done[c], cudaEventDisableTiming));
      NB_CUDA_CHECK(cudaEventCreateWithFlags(&w.run_done[c], cudaEventDisableTiming));
    }
    NB_CUDA_CHECK(cudaEventCreateWithFlags(&w.idle, cudaEventDisableTiming));
    w.ev_init = true;
    NB_CUDA_CHECK(cudaEventRecord(w.idle, w.s_out));
    w.device = device;
  }
  if (w.cap < bytes) {
    NB_CUDA_CHECK(cudaStreamSynchronize(w.s_in));
    NB_CUDA_CHECK(cudaStreamSynchronize(w.s_k[0]));
    NB_CUDA_CHECK(cudaStreamSynchronize(w.s_k[1]));
    NB_CUDA_CHECK(cudaStreamSynchronize(w.s_out));
    if (w.buf) cudaFree(w.buf);
    w.buf = nullptr;
    w.cap = 0;
    NB_CUDA_CHECK(cudaMalloc(&w.buf, bytes));
    w.cap = bytes;
  }
  return NB_OK;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int analyze_host_ex(const double* m, const double* q, double* v, const double* eps, double G, int B, int N, int mode,
                           unsigned prep_flags, double kick_dt, double sched_dt, double dt, int n_steps, int n_megno,
                           int split_n_max, const double* raw_dr, const double* raw_dv, double* dyn_features,
                           double* static_features, int32_t* n_sub_out, int32_t* status, int device, int slot,
                           const nb_host_opts* opts_in) {
  nb_host_opts o;
  std::memset(&o, 0, sizeof(o));
  if (opts_in) {
    if (opts_in->size < 8 || opts_in->size > 4096) { set_error("nb_ensemble_analyze_host_ex: opts->size is not sizeof(nb_host_opts)"); return NB_ERR_ARG; }
    std::memcpy(&o, opts_in, opts_in->size < sizeof(o) ? opts_in->size : sizeof(o));
  }
  const bool hamsoft = mode == NB_MODE_HAMSOFT;
  const bool adaptive = (o.flags & NB_HOST_ADAPTIVE) != 0;
  const bool dev_tangent = (o.flags & NB_HOST_DEVICE_TANGENT) != 0;
  const bool compact = (o.flags & NB_HOST_COMPACT_DYN) != 0;
  const bool keep_v = (o.flags & NB_HOST_KEEP_V) != 0;
  if (slot < 0 || slot >= NB_HOST_SLOTS) { set_error("nb_ensemble_analyze_host: slot out of range"); return NB_ERR_ARG; }
  if (!m || !q || !v || !eps || !dyn_features || B < 0 || N < NB_MIN_N || N > NB_MAX_N_MID || n_steps < 0 || n_megno < 0) { set_error("nb_ensemble_analyze_host: bad arguments"); return NB_ERR_ARG; }
  if (mode != NB_MODE_VERLET && mode != NB_MODE_YOSHIDA4 && mode != NB_MODE_WHFAST && !hamsoft) { set_error("nb_ensemble_analyze_host: unknown integrator mode"); return NB_ERR_ARG; }
  if (n_megno > 0 && !dev_tangent && (!raw_dr || !raw_dv)) { set_error("nb_ensemble_analyze_host: n_megno > 0 needs raw_dr/raw_dv (or NB_HOST_DEVICE_TANGENT)"); return NB_ERR_ARG; }
  if (adaptive && (hamsoft || mode == NB_MODE_WHFAST || !o.soft_par)) { set_error("nb_ensemble_analyze_host: NB_HOST_ADAPTIVE needs verlet / yoshida4 and opts->soft_par"); return NB_ERR_UNSUPPORTED; }
  if (B == 0) return NB_OK;
  NvtxRange r_host("nb_ensemble_analyze_host");
  std::lock_guard<std::mutex> lock(g_ws_mu);
  DeviceGuard guard(device);
  if (!guard.ok) { set_error("nb_ensemble_analyze_host: cudaSetDevice failed"); return NB_ERR_CUDA; }
  // ---- chunking
  int n_chunks = o.n_chunks;
  if (n_chunks <= 0) n_chunks = 1;
  if (n_chunks > NB_MAX_CHUNKS) n_chunks = NB_MAX_CHUNKS;
  if (n_chunks > B) n_chunks = B;
  const int per = (B + n_chunks - 1) / n_chunks;
  // ---- workspace (sized for the whole bucket; a chunk uses its own slice of every array)
  const size_t bn = (size_t)B * N;
  const size_t sz_m = align256(bn * 8), sz_q = align256(bn * 16), sz_b = align256((size_t)B * 8);
  const size_t sz_dyn = align256((size_t)B * NB_N_DYN * 8), sz_stat = align256((size_t)B * NB_N_STATIC * 8);
  const size_t sz_i = align256((size_t)B * 4);
  const size_t sz_hs = hamsoft ? align256((size_t)B * NB_HS_NPARAM * 8) : 0, sz_ep = hamsoft ? align256((size_t)B * 16) : 0;
  const size_t sz_ad = adaptive ? (align256((size_t)B * 24) + 3 * sz_b) : 0;
  const size_t sz_bins = align256(128 * 4) * (size_t)n_chunks;
  const size_t total = sz_m + 5 * sz_q + sz_b + 2 * sz_dyn + sz_stat + 3 * sz_i + sz_hs + sz_ep + sz_ad + sz_bins + 512;
  int rc = ws_reserve(slot, device, total);
  if (rc != NB_OK) return rc;
  HostWs& w = g_ws[slot];
  char* p = (char*)w.buf;
  double* d_m = (double*)p; p += sz_m;
  double* d_q = (double*)p; p += sz_q;
  double* d_v = (double*)p; p += sz_q;
  double* d_dr = (double*)p; p += sz_q;
  double* d_dv = (double*)p; p += sz_q;
  double* d_vk = (double*)p; p += sz_q;      // kicked velocities, frozen for the early D2H while the run advances d_v
  double* d_eps = (double*)p; p += sz_b;
  double* d_dyn = (double*)p; p += sz_dyn;
  double* d_dyn_c = (double*)p; p += sz_dyn;  // compact copy (17 columns)
  double* d_stat = (double*)p; p += sz_stat;
  int32_t* d_nsub = (int32_t*)p; p += sz_i;
  int32_t* d_perm = (int32_t*)p; p += sz_i;
  int32_t* d_status = (int32_t*)p; p += sz_i;
  double* d_hs = (double*)p; p += sz_hs;
  double* d_ep = (double*)p; p += sz_ep;
  double *d_par = nullptr, *d_eps_e = nullptr, *d_edelta = nullptr, *d_eps_run = nullptr;
  if (adaptive) {
    d_par = (double*)p; p += align256((size_t)B * 24);
    d_eps_e = (double*)p; p += sz_b;
    d_edelta = (double*)p; p += sz_b;
    d_eps_run = (double*)p; p += sz_b;
  }
  char* d_bins0 = p;
  const int interval = n_steps / 100 > 1 ? n_steps / 100 : 1;
  unsigned pf = prep_flags & 0xffu;
  if (static_features) pf |= NB_PREP_STATIC_FEATURES; else pf &= ~NB_PREP_STATIC_FEATURES;
  const bool kicked = (pf & (hamsoft ? NB_PREP_REMOVE_COM : (NB_PREP_REMOVE_COM | NB_PREP_CTOR_KICK | NB_PREP_SNAPSHOT_KICK))) != 0;
  const int ndyn_out = compact ? NB_N_DYN_USER : NB_N_DYN;
  // the workspace is reused: the previous call's last copy must have left it
  NB_CUDA_CHECK(cudaStreamWaitEvent(w.s_in, w.idle, 0));
  NB_CUDA_CHECK(cudaStreamWaitEvent(w.s_k[0], w.idle, 0));
  NB_CUDA_CHECK(cudaStreamWaitEvent(w.s_k[1], w.idle, 0));
  for (int c = 0; c < n_chunks; ++c) {
    const int lo = c * per;
    const int nb = (lo + per <= B ? per : B - lo);
    if (nb <= 0) break;
    const size_t o1 = (size_t)lo, oN = (size_t)lo * N, cn = (size_t)nb * N;
    cudaStream_t sk = w.s_k[c & 1];
    // ---- stage 1: H2D
    NB_CUDA_CHECK(cudaMemcpyAsync(d_m + oN, m + oN, cn * 8, cudaMemcpyHostToDevice, w.s_in));
    NB_CUDA_CHECK(cudaMemcpyAsync(d_q + 2 * oN, q + 2 * oN, cn * 16, cudaMemcpyHostToDevice, w.s_in));
    NB_CUDA_CHECK(cudaMemcpyAsync(d_v + 2 * oN, v + 2 * oN, cn * 16, cudaMemcpyHostToDevice, w.s_in));
    NB_CUDA_CHECK(cudaMemcpyAsync(d_eps + o1, eps + o1, (size_t)nb * 8, cudaMemcpyHostToDevice, w.s_in));
    if (n_megno > 0 && !dev_tangent) {
      NB_CUDA_CHECK(cudaMemcpyAsync(d_dr + 2 * oN, raw_dr + 2 * oN, cn * 16, cudaMemcpyHostToDevice, w.s_in));
      NB_CUDA_CHECK(cudaMemcpyAsync(d_dv + 2 * oN, raw_dv + 2 * oN, cn * 16, cudaMemcpyHostToDevice, w.s_in));
    }
    if (hamsoft && o.hs_params) NB_CUDA_CHECK(cudaMemcpyAsync(d_hs + o1 * NB_HS_NPARAM, o.hs_params + o1 * NB_HS_NPARAM, (size_t)nb * NB_HS_NPARAM * 8, cudaMemcpyHostToDevice, w.s_in));
    if (hamsoft && o.eps_pi) NB_CUDA_CHECK(cudaMemcpyAsync(d_ep + 2 * o1, o.eps_pi + 2 * o1, (size_t)nb * 16, cudaMemcpyHostToDevice, w.s_in));
    if (adaptive) {
      NB_CUDA_CHECK(cudaMemcpyAsync(d_par + 3 * o1, o.soft_par + 3 * o1, (size_t)nb * 24, cudaMemcpyHostToDevice, w.s_in));
      NB_CUDA_CHECK(cudaMemcpyAsync(d_eps_e + o1, (o.eps_energy ? o.eps_energy : eps) + o1, (size_t)nb * 8, cudaMemcpyHostToDevice, w.s_in));
      NB_CUDA_CHECK(cudaMemcpyAsync(d_eps_run + o1, (o.eps_start ? o.eps_start : eps) + o1, (size_t)nb * 8, cudaMemcpyHostToDevice, w.s_in));
      NB_CUDA_CHECK(cudaMemsetAsync(d_edelta + o1, 0, (size_t)nb * 8, w.s_in));
    }
    NB_CUDA_CHECK(cudaEventRecord(w.in_done[c], w.s_in));
    // ---- stage 2: construction-time work
    NB_CUDA_CHECK(cudaStreamWaitEvent(sk, w.in_done[c], 0));
    if (n_megno > 0 && dev_tangent) {
      rc = generate_tangent(N, nb, o.tangent_seed, o.first_index + (uint64_t)lo, d_dr + 2 * oN, d_dv + 2 * oN, sk);
      if (rc != NB_OK) return rc;
    }
    // ham_soft: COM removal only here (no corrector kick, hamiltonian_softening_integrator.py:753-754); the static
    // features follow the calibration below
    const unsigned pf1 = hamsoft ? (pf & NB_PREP_REMOVE_COM) : pf;
    PrepArgs pa{d_m + oN, d_q + 2 * oN, d_v + 2 * oN, d_eps + o1, G, nb, hamsoft ? NB_MODE_VERLET : mode, pf1, kick_dt,
                sched_dt, dt, split_n_max, nullptr, d_nsub + o1, d_stat + o1 * NB_N_STATIC};
    rc = N > NB_MAX_N ? mid_prepare(pa, N, sk) : ensemble_prepare(pa, N, sk);
    if (rc != NB_OK) return rc;
    if (hamsoft) {
      const int blocks = (nb + 127) / 128;
      if (!o.hs_params) hs_defaults_kernel<<<blocks, 128, 0, sk>>>(d_eps + o1, nb, d_hs + o1 * NB_HS_NPARAM, d_ep + 2 * o1, o.eps_pi ? 0 : 1);
      const bool calibrate = (o.flags & NB_HOST_HS_NO_CALIBRATE) == 0;
      const double dt0 = sched_dt != 0.0 ? sched_dt : dt;
      rc = hamsoft_setup(d_m + oN, d_q + 2 * oN, G, nb, N, (calibrate ? 1u : 0u) | 2u, dt0, d_hs + o1 * NB_HS_NPARAM,
                         d_ep + 2 * o1, d_nsub + o1, sk);
      if (rc != NB_OK) return rc;
      if (fabs(fabs(dt) - fabs(dt0)) > 0.01 * fabs(dt0)) {   // strang_substeps: re-freeze outside 1 % of the frozen dt
        hs_bump_mu_kernel<<<blocks, 128, 0, sk>>>(d_hs + o1 * NB_HS_NPARAM, nb, dt);
        rc = hamsoft_setup(d_m + oN, d_q + 2 * oN, G, nb, N, 2u, dt, d_hs + o1 * NB_HS_NPARAM, d_ep + 2 * o1,
                           d_nsub + o1, sk);
        if (rc != NB_OK) return rc;
      }
      if (pf & NB_PREP_STATIC_FEATURES) {
        hs_eps_gather_kernel<<<blocks, 128, 0, sk>>>(d_ep + 2 * o1, d_eps + o1, nb);
        PrepArgs ps{d_m + oN, d_q + 2 * oN, d_v + 2 * oN, d_eps + o1, G, nb, NB_MODE_VERLET, NB_PREP_STATIC_FEATURES, 0.0,
                    sched_dt, dt, split_n_max, nullptr, nullptr, d_stat + o1 * NB_N_STATIC};
        rc = N > NB_MAX_N ? mid_prepare(ps, N, sk) : ensemble_prepare(ps, N, sk);
        if (rc != NB_OK) return rc;
      }
    }
    // everything that is final now goes back while the run is in flight: the kicked velocities (the reference mutates
    // the caller's sims), the static features and n_sub -- 60 % of the D2H bytes
    const bool send_v = kicked && !keep_v;
    if (send_v) NB_CUDA_CHECK(cudaMemcpyAsync(d_vk + 2 * oN, d_v + 2 * oN, cn * 16, cudaMemcpyDeviceToDevice, sk));
    NB_CUDA_CHECK(cudaEventRecord(w.prep_done[c], sk));
    NB_CUDA_CHECK(cudaStreamWaitEvent(w.s_out, w.prep_done[c], 0));
    if (send_v) NB_CUDA_CHECK(cudaMemcpyAsync(v + 2 * oN, d_vk + 2 * oN, cn * 16, cudaMemcpyDeviceToHost, w.s_out));
    if (static_features) NB_CUDA_CHECK(cudaMemcpyAsync(static_features + o1 * NB_N_STATIC, d_stat + o1 * NB_N_STATIC, (size_t)nb * NB_N_STATIC * 8, cudaMemcpyDeviceToHost, w.s_out));
    if (n_sub_out) NB_CUDA_CHECK(cudaMemcpyAsync(n_sub_out + o1, d_nsub + o1, (size_t)nb * 4, cudaMemcpyDeviceToHost, w.s_out));
    // ---- stage 3: the run
    int32_t* bins = (int32_t*)(d_bins0 + (size_t)c * align256(128 * 4));
    rc = sort_by_nsub(d_nsub + o1, nb, N, d_perm + o1, bins, -1, sk);
    if (rc != NB_OK) return rc;
    if (hamsoft) {
      rc = hamsoft_run(d_m + oN, d_q + 2 * oN, d_v + 2 * oN, G, nb, N, NB_RUN_ENERGY | NB_RUN_WRITE_STATE, dt, n_steps,
                       interval, n_megno, d_nsub + o1, d_perm + o1, d_dr + 2 * oN, d_dv + 2 * oN, d_ep + 2 * o1,
                       d_hs + o1 * NB_HS_NPARAM, d_dyn + o1 * NB_N_DYN, d_status + o1, nullptr, nullptr, sk);
    } else if (adaptive) {
      rc = ensemble_analyze_adaptive(d_m + oN, d_q + 2 * oN, d_v + 2 * oN, d_eps_run + o1, d_eps_e + o1, d_par + 3 * o1, G,
                                     nb, N, mode, dt, n_steps, interval, n_megno, d_nsub + o1, d_dr + 2 * oN,
                                     d_dv + 2 * oN, o.k_wall > 0.0 ? o.k_wall : 1.0e9,
                                     o.barrier_exponent > 0 ? o.barrier_exponent : 5, d_edelta + o1,
                                     d_dyn + o1 * NB_N_DYN, d_status + o1, sk);
    } else {
      RunArgs ra{d_m + oN, d_q + 2 * oN, d_v + 2 * oN, d_eps + o1, G, nb, NB_RUN_ENERGY, dt, n_steps, interval, n_megno,
                 d_nsub + o1, d_perm + o1, bins + 64, 0, 0, 0, d_dr + 2 * oN, d_dv + 2 * oN, d_dyn + o1 * NB_N_DYN,
                 d_status + o1, nullptr, nullptr};
      rc = N > NB_MAX_N ? mid_run(ra, N, mode, sk) : ensemble_run_classic(ra, N, mode, sk);
    }
    if (rc != NB_OK) return rc;
    if (compact) compact_dyn_kernel<<<(nb * NB_N_DYN_USER + 255) / 256, 256, 0, sk>>>(d_dyn + o1 * NB_N_DYN, d_dyn_c + o1 * NB_N_DYN_USER, nb);
    NB_CUDA_CHECK(cudaGetLastError());
    NB_CUDA_CHECK(cudaEventRecord(w.run_done[c], sk));
    // ---- stage 4: results
    NB_CUDA_CHECK(cudaStreamWaitEvent(w.s_out, w.run_done[c], 0));
    NB_CUDA_CHECK(cudaMemcpyAsync(dyn_features + o1 * ndyn_out, (compact ? d_dyn_c : d_dyn) + o1 * ndyn_out, (size_t)nb * ndyn_out * 8, cudaMemcpyDeviceToHost, w.s_out));
    if (status) NB_CUDA_CHECK(cudaMemcpyAsync(status + o1, d_status + o1, (size_t)nb * 4, cudaMemcpyDeviceToHost, w.s_out));
    if (hamsoft && o.eps_pi) NB_CUDA_CHECK(cudaMemcpyAsync(o.eps_pi + 2 * o1, d_ep + 2 * o1, (size_t)nb * 16, cudaMemcpyDeviceToHost, w.s_out));
    if (adaptive && o.energy_delta) NB_CUDA_CHECK(cudaMemcpyAsync(o.energy_delta + o1, d_edelta + o1, (size_t)nb * 8, cudaMemcpyDeviceToHost, w.s_out));
  }
  NB_CUDA_CHECK(cudaEventRecord(w.idle, w.s_out));
  return NB_OK;
}

}  // namespace nb

using namespace nb;

extern "C" {

int nb_ensemble_analyze_host_ex(const double* m, const double* q, double* v, const double* eps, double G, int B, int N,
                                int mode, unsigned prep_flags, double kick_dt, double sched_dt, double dt, int n_steps,
                                int n_megno, int split_n_max, const double* raw_dr, const double* raw_dv,
                                double* dyn_features, double* static_features, int32_t* n_sub_out, int32_t* status,
                                int device, int slot, const nb_host_opts* opts) {
  return analyze_host_ex(m, q, v, eps, G, B, N, mode, prep_flags, kick_dt, sched_dt, dt, n_steps, n_megno, split_n_max,
                         raw_dr, raw_dv, dyn_features, static_features, n_sub_out, status, device, slot, opts);
}

int nb_ensemble_analyze_host_async(const double* m, const double* q, double* v, const double* eps, double G, int B, int N,
                                   int mode, unsigned prep_flags, double kick_dt, double sched_dt, double dt, int n_steps,
                                   int n_megno, int split_n_max, const double* raw_dr, const double* raw_dv,
                                   double* dyn_features, double* static_features, int32_t* n_sub_out, int32_t* status,
                                   int device, int slot) {
  return analyze_host_ex(m, q, v, eps, G, B, N, mode, prep_flags, kick_dt, sched_dt, dt, n_steps, n_megno, split_n_max,
                         raw_dr, raw_dv, dyn_features, static_features, n_sub_out, status, device, slot, nullptr);
}

int nb_host_sync(int slot) {
  if (slot < 0 || slot >= NB_HOST_SLOTS) { set_error("nb_host_sync: slot out of range"); return NB_ERR_ARG; }
  HostWs& w = g_ws[slot];
  if (w.device < 0) return NB_OK;
  DeviceGuard guard(w.device);
  NB_CUDA_CHECK(cudaStreamSynchronize(w.s_in));
  NB_CUDA_CHECK(cudaStreamSynchronize(w.s_k[0]));
  NB_CUDA_CHECK(cudaStreamSynchronize(w.s_k[1]));
  NB_CUDA_CHECK(cudaStreamSynchronize(w.s_out));
  return NB_OK;
}

int nb_ensemble_analyze_host(const double* m, const double* q, double* v, const double* eps, double G, int B, int N,
                             int mode, unsigned prep_flags, double kick_dt, double sched_dt, double dt, int n_steps,
                             int n_megno, int split_n_max, const double* raw_dr, const double* raw_dv, double* dyn_features,
                             double* static_features, int32_t* n_sub_out, int32_t* status, int device) {
  int rc = analyze_host_ex(m, q, v, eps, G, B, N, mode, prep_flags, kick_dt, sched_dt, dt, n_steps, n_megno, split_n_max,
                           raw_dr, raw_dv, dyn_features, static_features, n_sub_out, status, device, 0, nullptr);
  if (rc != NB_OK) return rc;
  return nb_host_sync(0);
}

}  // extern "C"
