// capi.cu -- the extern "C" boundary declared in include/nbody_b200.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "common.cuh"
#include "args.cuh"

namespace nb {

// ---- error plumbing -----------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
}
int cuda_fail(cudaError_t e, const char* where) {
  std::snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
  return NB_ERR_CUDA;
}

// ---- declarations of the launchers in the other translation units -------------------------------------
int ensemble_run_classic(const RunArgs& a, int N, int mode, cudaStream_t st);
int ensemble_prepare(const PrepArgs& a, int N, cudaStream_t st);
int pair_batched(const double* q, const double* m, const double* eps, double G, int B, int N, double* acc, double* U,
                 double* dV, cudaStream_t st);
int variational_batched(const double* q, const double* m, const double* s2, const double* dr, double G, int B, int N,
                        double* da, cudaStream_t st);
int sort_by_nsub(const int32_t* n_sub, int B, int N, int32_t* perm, int32_t* ws, int heavy_threshold, cudaStream_t st);
int generate_ensemble(int cohort, int N, int B, uint64_t seed, uint64_t first, double* m, double* q, double* v, double* eps,
                      cudaStream_t st);
int ensemble_analyze_adaptive(const double* m, double* q, double* v, double* eps, const double* eps_energy,
                              const double* soft_par, double G, int B, int N, int mode, double dt, int n_steps,
                              int sample_interval, int n_megno, const int32_t* n_sub, const double* raw_dr,
                              const double* raw_dv, double k_wall, int n_exp, double* e_delta, double* dyn, int32_t* status,
                              cudaStream_t st);
int mlp_classify(const double* dyn, const double* stat, const int32_t* idx, int F, const float* mean,
                 const float* inv_scale, const float* w1, const float* b1, const float* w2, const float* b2,
                 const float* w3, float b3, float threshold, int B, float* prob, int32_t* label, cudaStream_t st);
int ensemble_run_adaptive(const double* m, double* q, double* v, double* eps, const double* soft_par, double G, int B,
                          int N, int mode, double dt, int n_steps, const int32_t* n_sub, double k_wall, int n_exp,
                          double* e_delta, double* eps_hist, int32_t* status, cudaStream_t st);
int mid_run(const RunArgs& a, int N, int mode, cudaStream_t st);
int mid_prepare(const PrepArgs& a, int N, cudaStream_t st);
int mid_pair(const double* q, const double* m, const double* eps, double G, int B, int N, double* acc, double* U, double* dV,
             cudaStream_t st);
int mid_variational(const double* q, const double* m, const double* s2, const double* dr, double G, int B, int N, double* da,
                    cudaStream_t st);
int generate_tangent(int N, int B, uint64_t seed, uint64_t first, double* dr, double* dv, cudaStream_t st);
int largeN_accel(const float* xym, int n_total, int i0, int ni, float eps, float G, float* acc, double* sums,
                 double* workspace, int variant, cudaStream_t st);
int largeN_pass(int kind, const float* xym, const float* jaux, int n_total, int i0, int ni, const float* iparam,
                float eps, double* out, const float* boxes, cudaStream_t st);
int largeN_tile_boxes(const float* xym, const float* jaux, int n_total, float* boxes, cudaStream_t st);
int largeN_kick_drift(float* xym_local, float* vel, const float* acc, int ni, float kick_h, float drift_h,
                      cudaStream_t st);
int hamsoft_run(const double* m, double* q, double* v, double G, int B, int N, unsigned flags, double dt, int n_steps,
                int sample_interval, int n_megno, const int32_t* n_sub, const int32_t* perm, const double* raw_dr,
                const double* raw_dv, double* eps_pi, const double* hs, double* dyn, int32_t* status, double* work,
                unsigned long long* tstamp, cudaStream_t st);

int hamsoft_setup(const double* m, const double* q, double G, int B, int N, unsigned flags, double dt, double* hs,
                  double* eps_pi, int32_t* n_sub, cudaStream_t st);
int hamsoft_probe(const double* m, const double* q, const double* v, double G, int B, int N, const double* eps_pi,
                  const double* hs, double* out, cudaStream_t st);

// ---- peak micro-benchmarks ------------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) peak_kernel(int iters, float seed, float* out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (WHICH == 0) {
    double a[8];
    const double b = 1.0000001 + seed, c = 1e-9 * tid;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = k + seed;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456) out[0] = (float)s;
  } else if (WHICH == 1) {
    float a[16];
    const float b = 1.0000001f + seed, c = 1e-9f * tid;
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = k + seed;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], b, c);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 123.456f) out[0] = s;
  } else if (WHICH == 2) {
    unsigned long long a[16];
    unsigned long long b, c;
    {
      float2 fb = make_float2(1.0000001f + seed, 0.9999999f + seed), fc = make_float2(1e-9f * tid, 2e-9f * tid);
      b = *reinterpret_cast<unsigned long long*>(&fb);
      c = *reinterpret_cast<unsigned long long*>(&fc);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float2 fa = make_float2(k + seed, k - seed);
        a[k] = *reinterpret_cast<unsigned long long*>(&fa);
      }
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[k]) : "l"(b), "l"(c));
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float2 fa = *reinterpret_cast<float2*>(&a[k]);
      s += fa.x + fa.y;
    }
    if (s == 123.456f) out[0] = s;
  } else if (WHICH == 3) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.5f + k + seed + 1e-3f * tid;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456f) out[0] = s;
  } else {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.5 + k + seed + 1e-3 * tid;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a[k]));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456) out[0] = (float)s;
  }
}

static int peak_flops(int which, int device, double* tflops) {
  if (!tflops || which < 0 || which > 4) { set_error("nb_peak_flops: bad arguments"); return NB_ERR_ARG; }
  // the caller's current device is restored on every path; resources are released on every path
  int prev = -1;
  cudaGetDevice(&prev);
  cudaError_t err = cudaSetDevice(device);
  float* out = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int sms = 0;
  double best = 0.0;
  if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaMalloc(&out, 4);
  if (err == cudaSuccess) err = cudaEventCreate(&e0);
  if (err == cudaSuccess) err = cudaEventCreate(&e1);
  const int blocks = sms * 8, threads = 256;
  const int iters = 4096;
  for (int rep = 0; rep < 6 && err == cudaSuccess; ++rep) {
    err = cudaEventRecord(e0, st);
    if (err != cudaSuccess) break;
    switch (which) {
      case 0: peak_kernel<0><<<blocks, threads, 0, st>>>(iters, 0.f, out); break;
      case 1: peak_kernel<1><<<blocks, threads, 0, st>>>(iters, 0.f, out); break;
      case 2: peak_kernel<2><<<blocks, threads, 0, st>>>(iters, 0.f, out); break;
      case 3: peak_kernel<3><<<blocks, threads, 0, st>>>(iters, 0.f, out); break;
      default: peak_kernel<4><<<blocks, threads, 0, st>>>(iters, 0.f, out); break;
    }
    err = cudaEventRecord(e1, st);
    if (err == cudaSuccess) err = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
    if (err != cudaSuccess) break;
    const double per_thread = which == 0 ? 8.0 * 2 : which == 1 ? 16.0 * 2 : which == 2 ? 16.0 * 4 : 8.0;
    const double ops = per_thread * iters * (double)blocks * threads;
    const double t = ops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && t > best) best = t;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (out) cudaFree(out);
  if (st) cudaStreamDestroy(st);
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  if (err != cudaSuccess) return cuda_fail(err, "nb_peak_flops");
  *tflops = best;
  return NB_OK;
}

}  // namespace nb

using namespace nb;

extern "C" {

const char* nb_last_error(void) { return g_err; }
int nb_version(void) { return 100; }

int nb_pair_batched_f64(const double* q, const double* m, const double* eps, double G, int B, int N, double* acc,
                        double* U, double* dVdeps, void* stream) {
  if (!q || !m || !eps || B < 0 || N < NB_MIN_N || N > NB_MAX_N_MID) { set_error("nb_pair_batched_f64: bad arguments"); return NB_ERR_ARG; }
  if (B == 0) return NB_OK;
  if (N > NB_MAX_N) return mid_pair(q, m, eps, G, B, N, acc, U, dVdeps, (cudaStream_t)stream);
  return pair_batched(q, m, eps, G, B, N, acc, U, dVdeps, (cudaStream_t)stream);
}

int nb_variational_batched_f64(const double* q, const double* m, const double* s2, const double* dr, double G, int B,
                               int N, double* da, void* stream) {
  if (!q || !m || !s2 || !dr || !da || B < 0 || N < NB_MIN_N || N > NB_MAX_N_MID) { set_error("nb_variational_batched_f64: bad arguments"); return NB_ERR_ARG; }
  if (B == 0) return NB_OK;
  if (N > NB_MAX_N) return mid_variational(q, m, s2, dr, G, B, N, da, (cudaStream_t)stream);
  return variational_batched(q, m, s2, dr, G, B, N, da, (cudaStream_t)stream);
}

int nb_ensemble_prepare_f64(const double* m, const double* q, double* v, const double* eps, double G, int B, int N,
                            int mode, unsigned flags, double kick_dt, double sched_dt, double dt, int split_n_max,
                            double* h_sub_ref, int32_t* n_sub, double* static_features, void* stream) {
  if (!m || !q || !v || !eps || B < 0 || N < NB_MIN_N || N > NB_MAX_N_MID) { set_error("nb_ensemble_prepare_f64: bad arguments"); return NB_ERR_ARG; }
  if (B == 0) return NB_OK;
  PrepArgs a{m, q, v, eps, G, B, mode, flags, kick_dt, sched_dt, dt, split_n_max, h_sub_ref, n_sub, static_features};
  if (N > NB_MAX_N) return mid_prepare(a, N, (cudaStream_t)stream);
  return ensemble_prepare(a, N, (cudaStream_t)stream);
}

int nb_ensemble_run_counted_f64(const double* m, double* q, double* v, const double* eps, double G, int B, int N, int mode,
                                unsigned flags, double dt, int n_steps, int sample_interval, int n_megno,
                                const int32_t* n_sub, const int32_t* perm, const int32_t* n_heavy, const double* raw_dr,
                                const double* raw_dv, double* eps_pi, const double* hs_params, double* dyn_features,
                                int32_t* status, double* work, uint64_t* t_main, void* stream) {
  if (!m || !q || !v || B < 0 || N < NB_MIN_N || N > NB_MAX_N_MID || n_steps < 0 || n_megno < 0) { set_error("nb_ensemble_run_f64: bad arguments"); return NB_ERR_ARG; }
  if (n_megno > 0 && (!raw_dr || !raw_dv)) { set_error("nb_ensemble_run_f64: n_megno > 0 needs raw_dr/raw_dv"); return NB_ERR_ARG; }
  if (B == 0) return NB_OK;
  if (mode == NB_MODE_HAMSOFT) {
    if (!eps_pi || !hs_params) { set_error("nb_ensemble_run_f64: ham_soft needs eps_pi and hs_params"); return NB_ERR_ARG; }
    return hamsoft_run(m, q, v, G, B, N, flags, dt, n_steps, sample_interval, n_megno, n_sub, perm, raw_dr, raw_dv,
                       eps_pi, hs_params, dyn_features, status, work, (unsigned long long*)t_main, (cudaStream_t)stream);
  }
  if (!eps) { set_error("nb_ensemble_run_f64: eps is required"); return NB_ERR_ARG; }
  RunArgs a{m, q, v, eps, G, B, flags, dt, n_steps, sample_interval, n_megno, n_sub, perm, perm ? n_heavy : nullptr, 0, 0, 0, raw_dr, raw_dv, dyn_features, status, (unsigned long long*)t_main, work};
  if (N > NB_MAX_N) return mid_run(a, N, mode, (cudaStream_t)stream);    // 9..64 bodies: one CTA per system
  return ensemble_run_classic(a, N, mode, (cudaStream_t)stream);
}

int nb_ensemble_run_f64(const double* m, double* q, double* v, const double* eps, double G, int B, int N, int mode,
                        unsigned flags, double dt, int n_steps, int sample_interval, int n_megno, const int32_t* n_sub,
                        const int32_t* perm, const int32_t* n_heavy, const double* raw_dr, const double* raw_dv,
                        double* eps_pi, const double* hs_params, double* dyn_features, int32_t* status, void* stream) {
  return nb_ensemble_run_counted_f64(m, q, v, eps, G, B, N, mode, flags, dt, n_steps, sample_interval, n_megno, n_sub, perm,
                                     n_heavy, raw_dr, raw_dv, eps_pi, hs_params, dyn_features, status, nullptr, nullptr, stream);
}

int nb_hamsoft_setup_f64(const double* m, const double* q, double G, int B, int N, unsigned flags, double dt,
                         double* hs_params, double* eps_pi, int32_t* n_sub, void* stream) {
  if (!m || !q || !hs_params || !eps_pi || B < 0 || N < NB_MIN_N || N > NB_MAX_N_MID || ((flags & 2u) && !n_sub)) { set_error("nb_hamsoft_setup_f64: bad arguments"); return NB_ERR_ARG; }
  if (B == 0) return NB_OK;
  return hamsoft_setup(m, q, G, B, N, flags, dt, hs_params, eps_pi, n_sub, (cudaStream_t)stream);
}

int nb_hamsoft_probe_f64(const double* m, const double* q, const double* v, double G, int B, int N,
                         const double* eps_pi, const double* hs_params, double* out, void* stream) {
  if (!m || !q || !v || !hs_params || !eps_pi || !out || B < 0 || N < NB_MIN_N || N > NB_MAX_N_MID) { set_error("nb_hamsoft_probe_f64: bad arguments"); return NB_ERR_ARG; }
  if (B == 0) return NB_OK;
  return hamsoft_probe(m, q, v, G, B, N, eps_pi, hs_params, out, (cudaStream_t)stream);
}

int nb_ensemble_run_adaptive_f64(const double* m, double* q, double* v, double* eps, const double* soft_par, double G,
                                 int B, int N, int mode, double dt, int n_steps, const int32_t* n_sub, double k_wall,
                                 int barrier_exponent, double* energy_delta, double* eps_hist, int32_t* status,
                                 void* stream) {
  return ensemble_run_adaptive(m, q, v, eps, soft_par, G, B, N, mode, dt, n_steps, n_sub, k_wall, barrier_exponent,
                               energy_delta, eps_hist, status, (cudaStream_t)stream);
}

int nb_ensemble_analyze_adaptive_f64(const double* m, double* q, double* v, double* eps, const double* eps_energy,
                                     const double* soft_par, double G, int B, int N, int mode, double dt, int n_steps,
                                     int sample_interval, int n_megno, const int32_t* n_sub, const double* raw_dr,
                                     const double* raw_dv, double k_wall, int barrier_exponent, double* energy_delta,
                                     double* dyn_features, int32_t* status, void* stream) {
  return ensemble_analyze_adaptive(m, q, v, eps, eps_energy, soft_par, G, B, N, mode, dt, n_steps, sample_interval, n_megno,
                                   n_sub, raw_dr, raw_dv, k_wall, barrier_exponent, energy_delta, dyn_features, status,
                                   (cudaStream_t)stream);
}

int nb_sort_by_nsub(const int32_t* n_sub, int B, int N, int32_t* perm, int32_t* workspace, int heavy_threshold,
                    void* stream) {
  if (!n_sub || !perm || !workspace || B < 0) { set_error("nb_sort_by_nsub: bad arguments"); return NB_ERR_ARG; }
  if (B == 0) return NB_OK;
  return sort_by_nsub(n_sub, B, N, perm, workspace, heavy_threshold, (cudaStream_t)stream);
}

int nb_largeN_accel_f32(const float* xym, int n_total, int i0, int ni, float eps, float G, float* acc, double* sums,
                        double* workspace, int variant, void* stream) {
  return largeN_accel(xym, n_total, i0, ni, eps, G, acc, sums, workspace, variant, (cudaStream_t)stream);
}

int nb_largeN_pass_f32(int kind, const float* xym, const float* jaux, int n_total, int i0, int ni, const float* iparam,
                       float eps, double* out, const float* tile_boxes, void* stream) {
  return largeN_pass(kind, xym, jaux, n_total, i0, ni, iparam, eps, out, tile_boxes, (cudaStream_t)stream);
}
int nb_largeN_tile_boxes_f32(const float* xym, const float* jaux, int n_total, float* tile_boxes, void* stream) {
  return largeN_tile_boxes(xym, jaux, n_total, tile_boxes, (cudaStream_t)stream);
}
int nb_largeN_kick_drift_f32(float* xym_local, float* vel, const float* acc, int ni, float kick_h, float drift_h,
                             void* stream) {
  return largeN_kick_drift(xym_local, vel, acc, ni, kick_h, drift_h, (cudaStream_t)stream);
}

int nb_mlp_classify_f32(const double* dyn_features, const double* static_features, const int32_t* feature_index, int F,
                        const float* mean, const float* inv_scale, const float* w1, const float* b1, const float* w2,
                        const float* b2, const float* w3, float b3, float threshold, int B, float* prob, int32_t* label,
                        void* stream) {
  return mlp_classify(dyn_features, static_features, feature_index, F, mean, inv_scale, w1, b1, w2, b2, w3, b3, threshold,
                      B, prob, label, (cudaStream_t)stream);
}

int nb_generate_ensemble_f64(int cohort, int N, int B, uint64_t seed, uint64_t first_index, double* m, double* q, double* v,
                             double* eps, void* stream) {
  return generate_ensemble(cohort, N, B, seed, first_index, m, q, v, eps, (cudaStream_t)stream);
}

int nb_generate_tangent_f64(int N, int B, uint64_t seed, uint64_t first_index, double* dr, double* dv, void* stream) {
  return generate_tangent(N, B, seed, first_index, dr, dv, (cudaStream_t)stream);
}

int nb_peak_flops(int which, int device, double* tflops) { return peak_flops(which, device, tflops); }

}  // extern "C"
