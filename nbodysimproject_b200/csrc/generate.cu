// generate.cu -- on-GPU initial conditions for ensembles.
//
// Same distributions as the host generators (initial_condition_generator.py:49-104, specialized_generators.py:23-94,
// ml_training_pipeline.py:44-122 as vectorised in generators.py: EnsembleInputs), drawn with a COUNTER-BASED RNG
// (Philox4x32-10 keyed by (seed, cohort) and counted by the GLOBAL system index), so a system's initial condition
// depends only on (seed, index): any sharding over ranks or batches reproduces the same ensemble, and building 10^6
// systems costs one kernel launch instead of minutes of Python object construction.  One thread per system.
// The host generators remain the bit-compatible path for reproducing a reference run; this path is statistically
// equivalent, not bit-equal (different RNG).
#include "common.cuh"

namespace nb {

struct Philox {
  uint32_t key[2], ctr[4], out[4];
  int have;
  __device__ Philox(uint64_t seed, uint64_t stream, uint64_t index) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = (uint32_t)index; ctr[1] = (uint32_t)(index >> 32); ctr[2] = (uint32_t)stream; ctr[3] = 0u;
    have = 0;
  }
  __device__ void round10() {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    ++ctr[3];                       // next block of this (seed, stream, index)
    have = 4;
  }
  __device__ uint32_t next32() {
    if (have == 0) round10();
    return out[4 - have--];
  }
  __device__ double uniform() {     // (0, 1), 53 bits
    const uint64_t hi = next32(), lo = next32();
    const uint64_t b = ((hi << 32) | lo) >> 11;
    return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ double uniform(double a, double b) { return a + (b - a) * uniform(); }
  __device__ void normal2(double& z0, double& z1) {   // Box-Muller
    const double u = uniform(), v = uniform();
    const double r = sqrt(-2.0 * log(u));
    double s, c;
    sincospi(2.0 * v, &s, &c);
    z0 = r * c; z1 = r * s;
  }
};

struct GenArgs {
  int cohort, B;
  uint64_t seed, first;
  double* m; double* q; double* v; double* eps;
};

template <int N>
__device__ __forceinline__ void remove_com(const double* m, double* vx, double* vy) {
  double M = 0.0, px = 0.0, py = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) { M += m[i]; px += m[i] * vx[i]; py += m[i] * vy[i]; }
#pragma unroll
  for (int i = 0; i < N; ++i) { vx[i] -= px / M; vy[i] -= py / M; }
}

// cohorts: 0 random (virial), 1 hierarchical triple (N = 3), 2 equal-mass polygon, 3 close encounter,
//          4 planetary resonant chain, 5 planetary TTV (first planet 10x heavier)
template <int N>
__global__ void __launch_bounds__(128) generate_kernel(GenArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.B) return;
  Philox g(a.seed, (uint64_t)a.cohort * 16 + N, a.first + (uint64_t)t);
  double m[N], x[N], y[N], vx[N], vy[N], eps = 0.0;
  const int cohort = a.cohort;
  if (cohort == 0 || cohort == 3) {
    const bool close = cohort == 3;
    const double scale = close ? 0.1 : g.uniform(0.5, 2.0);
    const double frac = close ? 1.5 : g.uniform(0.8, 1.2);
    const double pert = close ? 0.3 : g.uniform(0.05, 0.2);
    eps = close ? 0.001 : g.uniform(0.001, 0.1);
    const bool logm = !close && (((a.first + (uint64_t)t) & 1ull) == 0ull);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      m[i] = logm ? exp(g.uniform(log(0.1), log(10.0))) : g.uniform(0.1, 10.0);
      double z0, z1;
      g.normal2(z0, z1);
      x[i] = z0 * scale; y[i] = z1 * scale;
    }
    double U = 0.0, M = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      M += m[i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) {
        const double dx = x[i] - x[j], dy = y[i] - y[j];
        U -= m[i] * m[j] / (sqrt(dx * dx + dy * dy) + eps);
      }
    }
    const double K = -U / 2.0 * frac;
    const double vchar = sqrt(2.0 * fmax(K, 1e-300) / M);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double z0, z1;
      g.normal2(z0, z1);
      const double nrm = sqrt(z0 * z0 + z1 * z1);
      vx[i] = z0 / nrm * vchar; vy[i] = z1 / nrm * vchar;
    }
    remove_com<N>(m, vx, vy);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double z0, z1;
      g.normal2(z0, z1);
      vx[i] += z0 * vchar * pert; vy[i] += z1 * vchar * pert;
    }
    remove_com<N>(m, vx, vy);
  } else if (cohort == 1) {
    const double m2 = g.uniform(0.1, 1.0), m3 = g.uniform(0.1, 2.0), sep = g.uniform(3.0, 50.0);
    const double a_out = fmax(sep, 5.0);
#pragma unroll
    for (int i = 0; i < N; ++i) { m[i] = 1.0; x[i] = 0.0; y[i] = 0.0; vx[i] = 0.0; vy[i] = 0.0; }
    if (N >= 3) {
      m[1] = m2; m[2] = m3;
      x[0] = -m2 / (1.0 + m2); x[1] = 1.0 / (1.0 + m2); x[2] = a_out;
      const double v_in = sqrt(1.0 + m2), v_out = sqrt((1.0 + m2 + m3) / a_out);
      vy[0] = -m2 * v_in / (1.0 + m2); vy[1] = v_in / (1.0 + m2); vy[2] = v_out;
    }
    remove_com<N>(m, vx, vy);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double z0, z1;
      g.normal2(z0, z1);
      vx[i] += 0.05 * z0; vy[i] += 0.05 * z1;
    }
    eps = 0.01;
  } else if (cohort == 2) {
    const double radius = g.uniform(0.5, 3.0), rot = g.uniform(0.0, 1.0);
    const double vs = sqrt((double)N / radius) * rot;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s, c;
      sincospi(2.0 * (double)i / (double)N, &s, &c);
      m[i] = 1.0; x[i] = radius * c; y[i] = radius * s; vx[i] = -vs * s; vy[i] = vs * c;
    }
    eps = 0.05;
  } else {   // planetary: star + N-1 planets near 3:2 / 2:1 / 5:3 chains, circular, coplanar
    m[0] = 1.0; x[0] = 0.0; y[0] = 0.0; vx[0] = 0.0; vy[0] = 0.0;
    double P = 1.0;
#pragma unroll
    for (int i = 1; i < N; ++i) {
      m[i] = exp(g.uniform(-6.0, -3.0) * 2.302585092994046);
      if (cohort == 5 && i == 1) m[i] *= 10.0;
      if (i > 1) {
        const int pick = (int)(g.next32() % 3u);
        const double ratio = (pick == 0 ? 1.5 : (pick == 1 ? 2.0 : 5.0 / 3.0)) * (1.0 + g.uniform(-0.02, 0.02));
        P *= ratio;
      }
      const double sma = cbrt(P * P);
      double s, c;
      sincospi(2.0 * g.uniform(), &s, &c);
      const double vc = sqrt((1.0 + m[i]) / sma);
      x[i] = sma * c; y[i] = sma * s; vx[i] = -vc * s; vy[i] = vc * c;
    }
    eps = 0.0;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a.m[(size_t)t * N + i] = m[i];
    a.q[((size_t)t * N + i) * 2 + 0] = x[i];
    a.q[((size_t)t * N + i) * 2 + 1] = y[i];
    a.v[((size_t)t * N + i) * 2 + 0] = vx[i];
    a.v[((size_t)t * N + i) * 2 + 1] = vy[i];
  }
  a.eps[t] = eps;
}

__global__ void __launch_bounds__(128) tangent_kernel(int N, int B, uint64_t seed, uint64_t first, double* dr, double* dv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B) return;
  Philox g(seed, 0x7a6e67ull, first + (uint64_t)t);       // its own stream: independent of the initial conditions
  double* r = dr + (size_t)t * N * 2;
  double* w = dv + (size_t)t * N * 2;
  for (int i = 0; i < N; ++i) g.normal2(r[2 * i], r[2 * i + 1]);
  for (int i = 0; i < N; ++i) g.normal2(w[2 * i], w[2 * i + 1]);
}

int generate_tangent(int N, int B, uint64_t seed, uint64_t first, double* dr, double* dv, cudaStream_t st) {
  if (!dr || !dv || B < 0 || N < 1) { set_error("nb_generate_tangent_f64: bad arguments"); return NB_ERR_ARG; }
  if (B == 0) return NB_OK;
  tangent_kernel<<<(B + 127) / 128, 128, 0, st>>>(N, B, seed, first, dr, dv);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int generate_ensemble(int cohort, int N, int B, uint64_t seed, uint64_t first, double* m, double* q, double* v, double* eps,
                      cudaStream_t st) {
  if (!m || !q || !v || !eps || B < 0 || cohort < 0 || cohort > 5 || (cohort == 1 && N != 3)) {
    set_error("nb_generate_ensemble_f64: bad arguments (cohort 0..5; the hierarchical cohort has N = 3)");
    return NB_ERR_ARG;
  }
  if (B == 0) return NB_OK;
  GenArgs a{cohort, B, seed, first, m, q, v, eps};
  const int threads = 128, blocks = (B + threads - 1) / threads;
  switch (N) {
    case 2: generate_kernel<2><<<blocks, threads, 0, st>>>(a); break;
    case 3: generate_kernel<3><<<blocks, threads, 0, st>>>(a); break;
    case 4: generate_kernel<4><<<blocks, threads, 0, st>>>(a); break;
    case 5: generate_kernel<5><<<blocks, threads, 0, st>>>(a); break;
    case 6: generate_kernel<6><<<blocks, threads, 0, st>>>(a); break;
    case 7: generate_kernel<7><<<blocks, threads, 0, st>>>(a); break;
    case 8: generate_kernel<8><<<blocks, threads, 0, st>>>(a); break;
    default: set_error("N must be in 2..8"); return NB_ERR_ARG;
  }
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
