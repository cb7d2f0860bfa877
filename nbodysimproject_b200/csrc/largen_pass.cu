// largen_pass.cu -- the O(N^2) reductions of the ham_soft epsilon flow for ONE large-N system (C5).
//
// The reference evaluates the eps* model with Python double loops over pairs (hamsoft_eps_model.py:316-400
// `_solve_hi`, :451-556 `_production_grad`, softening.py:86-131 `grad_eps_target`,
// hamiltonian_softening_integrator.py:251-296 tau_grav) -- O(N^2) per call, impossible beyond N ~ 1e3.
// Here every such double loop is one pass of the same tile pipeline as the force kernel: a TMA bulk copy
// lands the AoS (x, y, m, *) tile (plus an optional per-j float2 tile) in shared memory, the CTA transposes
// it into rows, and the pair arithmetic runs as packed f32x2 over pairs of j-particles with fp32
// accumulation per tile promoted to fp64 across tiles.  Rank-local i-range [i0, i0+ni) like the force kernel.
//
//   DENSITY : out[i] = { S0_i, S1_i } = sum_{j != i} m_j e^{-r^2/h_i^2} { 1, r^2 }          (_solve_hi sweep;
//             Sigma_i = S0/(pi h^2), dSigma/dh = (-2 S0/h + 2 S1/h^3)/(pi h^2) for _production_grad)
//   EPSGRAD : out[i] = sum_{j != i} (q_i - q_j) [ A_i m_j e^{-r^2/h_i^2} + A_j m_i e^{-r^2/h_j^2} ]   (_production_grad,
//             gather form of its scatter loop; A = s (-2/(pi h^4)))
//   UNITGRAD: out[i] = sum_{j != i} (q_i - q_j) / r^3                                       (legacy grad direction)
//   TAUMIN  : out[i] = min_{j != i} (r^2 + eps^2)^{3/2} / (m_i + m_j)                         (tau_grav^2 G)
#include "largen_tile.cuh"

namespace nb {

constexpr int LP_TILE = 512;
constexpr int LP_IPT = 2;

enum { LP_DENSITY = 0, LP_EPSGRAD = 1, LP_UNITGRAD = 2, LP_TAUMIN = 3 };

// ---- locality culling ------------------------------------------------------------------------------------------
// exp(-r^2/h^2) is evaluated as ex2.approx.ftz(r^2 nk) with nk = -log2(e)/h^2: it is EXACTLY +0 once r^2 |nk| > 126
// (the result would be subnormal and .ftz flushes it), i.e. beyond r > 9.35 h.  Smoothing lengths are a few mean
// inter-particle distances, so for a spatially ordered particle array (LargeNHamSoftSimulation sorts along a Morton
// curve) almost every (i-block, j-tile) pair of the DENSITY and EPSGRAD passes contributes exact zeros.  A tile is
// skipped -- no TMA load, no arithmetic -- when the distance between the bounding boxes of the i-block and of the tile
// puts every pair beyond that radius for the largest h involved; the result is bit-identical to the unculled pass
// (only exact zeros are dropped; tests/test_gpu_largen_hamsoft.py) and the cost drops from O(N^2) to O(N x neighbours).
// Correctness never depends on the ordering, only the hit rate does.  UNITGRAD / TAUMIN (long range) are never culled.
struct TileBox {
  float xmin, ymin, xmax, ymax;
  float kmin;      // min_j |jaux.x| of the tile (EPSGRAD: the j-side exponent scale), +inf when no jaux
  float pad[3];
};
constexpr float LP_CULL = 127.0f;      // > 126 with a margin for the fp32 rounding of r^2 vs the box distance
constexpr int LP_MAX_TILES = 1024;     // tiles per (i-block, j-chunk) unit (the launcher caps the j-chunk accordingly)

__global__ void __launch_bounds__(256) largeN_tile_box_kernel(const float4* __restrict__ xym, const float2* __restrict__ jaux,
                                                              int n_total, int tile, TileBox* boxes) {
  const int t = blockIdx.x;
  const int j0 = t * tile, j1 = min(n_total, j0 + tile);
  float xmin = 3.0e38f, ymin = 3.0e38f, xmax = -3.0e38f, ymax = -3.0e38f, kmin = 3.0e38f;
  for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
    const float4 p = xym[j];
    xmin = fminf(xmin, p.x); xmax = fmaxf(xmax, p.x);
    ymin = fminf(ymin, p.y); ymax = fmaxf(ymax, p.y);
    if (jaux) kmin = fminf(kmin, fabsf(jaux[j].x));
  }
  __shared__ float red[5][8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, off));
    ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, off));
    xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, off));
    ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, off));
    kmin = fminf(kmin, __shfl_xor_sync(0xffffffffu, kmin, off));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = xmin; red[1][w] = ymin; red[2][w] = xmax; red[3][w] = ymax; red[4][w] = kmin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      xmin = fminf(xmin, red[0][k]); ymin = fminf(ymin, red[1][k]);
      xmax = fmaxf(xmax, red[2][k]); ymax = fmaxf(ymax, red[3][k]); kmin = fminf(kmin, red[4][k]);
    }
    TileBox b;
    b.xmin = xmin; b.ymin = ymin; b.xmax = xmax; b.ymax = ymax; b.kmin = kmin;
    b.pad[0] = b.pad[1] = b.pad[2] = 0.f;
    boxes[t] = b;
  }
}

struct PassArgs {
  const TileBox* boxes;   // per LP_TILE-aligned j-tile, or null (no culling)
  const float4* xym;
  const float2* jaux;     // EPSGRAD: (nk_j, A_j) per particle, nk = -log2(e)/h^2
  const float* iparam;    // DENSITY: h_i for the local particles
  int n_total, i0, ni;
  float eps2;
  double* out;            // [ni][2] (TAUMIN: [ni][1]); zeroed by the launcher for the additive kinds
  int n_ichunks, n_jchunks, jchunk;
};

template <int KIND>
__global__ void __launch_bounds__(LN_TPB, 4) largeN_pass_kernel(PassArgs a) {
  constexpr bool AUX = (KIND == LP_EPSGRAD);
  constexpr int NROW = AUX ? 5 : 3;
  __shared__ __align__(128) float4 raw[2][LP_TILE];
  __shared__ __align__(128) float2 rawaux[AUX ? 2 : 1][AUX ? LP_TILE : 1];
  __shared__ __align__(16) float rows[2][NROW][LP_TILE];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ unsigned short act[LP_MAX_TILES];       // tiles of this unit that survive the culling test, ascending
  __shared__ float bred[5][LN_TPB / 32];
  __shared__ int wcount[LN_TPB / 32 + 1];
  constexpr bool CULL = (KIND == LP_DENSITY || KIND == LP_EPSGRAD);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase[2] = {0u, 0u};
  const int n_units = a.n_ichunks * a.n_jchunks;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int ic = unit / a.n_jchunks;
    const int jc = unit - ic * a.n_jchunks;
    const int j_begin = jc * a.jchunk;
    const int j_end = min(a.n_total, j_begin + a.jchunk);
    float2 nxi[LP_IPT], nyi[LP_IPT], nk[LP_IPT];
    float mi[LP_IPT], Ai[LP_IPT];
    int ii[LP_IPT];
#pragma unroll
    for (int k = 0; k < LP_IPT; ++k) {
      ii[k] = ic * (LN_TPB * LP_IPT) + k * LN_TPB + tid;
      const int li = min(ii[k], a.ni - 1);
      const int gi = a.i0 + li;
      const float4 p = a.xym[gi];
      nxi[k] = make_float2(-p.x, -p.x);
      nyi[k] = make_float2(-p.y, -p.y);
      mi[k] = p.z;
      Ai[k] = 0.f;
      nk[k] = make_float2(0.f, 0.f);
      if (KIND == LP_DENSITY) {
        const float h = fmaxf(a.iparam[li], 1e-12f);
        const float v = -1.4426950408889634f / (h * h);
        nk[k] = make_float2(v, v);
      }
      if (KIND == LP_EPSGRAD) {
        const float2 ja = a.jaux[gi];
        nk[k] = make_float2(ja.x, ja.x);
        Ai[k] = ja.y;
      }
      if (KIND == LP_TAUMIN) nk[k] = make_float2(p.z, p.z);     // m_i, packed
    }
    double o0[LP_IPT], o1[LP_IPT], o2[LP_IPT], o3[LP_IPT];
#pragma unroll
    for (int k = 0; k < LP_IPT; ++k) {
      o0[k] = (KIND == LP_TAUMIN) ? 1e300 : 0.0;
      o1[k] = 0.0; o2[k] = 0.0; o3[k] = 0.0;
    }
    const int n_tiles_all = (j_end - j_begin + LP_TILE - 1) / LP_TILE;
    // ---- culling: bounding box and largest cut-off radius of this i-block, then the ordered list of live tiles
    int n_tiles = n_tiles_all;
    if (CULL && a.boxes != nullptr) {
      float xmin = 3.0e38f, ymin = 3.0e38f, xmax = -3.0e38f, ymax = -3.0e38f, kmin = 3.0e38f;
#pragma unroll
      for (int k = 0; k < LP_IPT; ++k) {
        xmin = fminf(xmin, -nxi[k].x); xmax = fmaxf(xmax, -nxi[k].x);
        ymin = fminf(ymin, -nyi[k].x); ymax = fmaxf(ymax, -nyi[k].x);
        kmin = fminf(kmin, fabsf(nk[k].x));
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, off));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, off));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, off));
        ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, off));
        kmin = fminf(kmin, __shfl_xor_sync(0xffffffffu, kmin, off));
      }
      if ((tid & 31) == 0) { bred[0][tid >> 5] = xmin; bred[1][tid >> 5] = ymin; bred[2][tid >> 5] = xmax; bred[3][tid >> 5] = ymax; bred[4][tid >> 5] = kmin; }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < LN_TPB / 32; ++k) {
        xmin = fminf(xmin, bred[0][k]); ymin = fminf(ymin, bred[1][k]);
        xmax = fmaxf(xmax, bred[2][k]); ymax = fmaxf(ymax, bred[3][k]); kmin = fminf(kmin, bred[4][k]);
      }
      int n_act = 0;
      for (int t0 = 0; t0 < n_tiles_all; t0 += LN_TPB) {
        const int t = t0 + tid;
        bool live = false;
        if (t < n_tiles_all) {
          const TileBox b = a.boxes[j_begin / LP_TILE + t];
          const float gx = fmaxf(0.f, fmaxf(b.xmin - xmax, xmin - b.xmax));
          const float gy = fmaxf(0.f, fmaxf(b.ymin - ymax, ymin - b.ymax));
          const float d2 = gx * gx + gy * gy;
          const float k_eff = (KIND == LP_EPSGRAD) ? fminf(kmin, b.kmin) : kmin;
          live = !(d2 * k_eff > LP_CULL);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        if ((tid & 31) == 0) wcount[tid >> 5] = __popc(bal);
        __syncthreads();
        int before = n_act;
        for (int w = 0; w < (tid >> 5); ++w) before += wcount[w];
        if (live) act[before + __popc(bal & ((1u << (tid & 31)) - 1u))] = (unsigned short)t;
        int tot = 0;
#pragma unroll
        for (int w = 0; w < LN_TPB / 32; ++w) tot += wcount[w];
        n_act += tot;
        __syncthreads();
      }
      n_tiles = n_act;
    } else {
      for (int t = tid; t < n_tiles_all; t += LN_TPB) act[t] = (unsigned short)t;
      __syncthreads();
    }
    auto issue = [&](int idx, int buf) {
      if (tid == 0) {
        const int j0 = j_begin + (int)act[idx] * LP_TILE;
        const int cnt = min(LP_TILE, j_end - j0);
        // bulk copies move multiples of 16 bytes: the float2 tile is rounded up to an even count (the caller
        // pads jaux to an even length)
        const uint32_t aux_bytes = AUX ? (uint32_t)((cnt + 1) & ~1) * 8u : 0u;
        mbar_expect_tx(&bars[buf], (uint32_t)cnt * 16u + aux_bytes);
        tma_load_1d(&raw[buf][0], a.xym + j0, (uint32_t)cnt * 16u, &bars[buf]);
        if (AUX) tma_load_1d(&rawaux[AUX ? buf : 0][0], a.jaux + j0, aux_bytes, &bars[buf]);
      }
    };
    if (n_tiles > 0) issue(0, 0);
    if (n_tiles > 1) issue(1, 1);
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int jt0 = j_begin + (int)act[t] * LP_TILE;
      const int cnt = min(LP_TILE, j_end - jt0);
      mbar_wait(&bars[buf], phase[buf]);
      phase[buf] ^= 1u;
#pragma unroll
      for (int r = 0; r < LP_TILE / LN_TPB; ++r) {
        const int j = r * LN_TPB + tid;
        float4 p = raw[buf][j];
        if (j >= cnt) p = make_float4(0.f, 0.f, 0.f, 0.f);
        rows[buf][0][j] = p.x;
        rows[buf][1][j] = p.y;
        // UNITGRAD ignores the masses (softening.py:86-131); padding must still contribute nothing.
        // TAUMIN: padding gets an infinite mass so that its value (rho^3 / inf = 0 ... ) is masked below instead.
        rows[buf][2][j] = (KIND == LP_UNITGRAD) ? (j < cnt ? 1.f : 0.f) : p.z;
        if (AUX) {
          float2 q = rawaux[AUX ? buf : 0][j];
          if (j >= cnt) q = make_float2(0.f, 0.f);
          rows[buf][NROW - 2][j] = q.x;
          rows[buf][NROW - 1][j] = q.y;
        }
      }
      __syncthreads();
      if (t + 2 < n_tiles) issue(t + 2, buf);

      float2 s0[LP_IPT], s1[LP_IPT], s2[LP_IPT], s3[LP_IPT];
#pragma unroll
      for (int k = 0; k < LP_IPT; ++k) {
        const float init = (KIND == LP_TAUMIN) ? 3.0e38f : 0.f;
        s0[k] = make_float2(init, init);
        s1[k] = make_float2(0.f, 0.f); s2[k] = s1[k]; s3[k] = s1[k];
      }
      const int cnt4 = (cnt + 3) & ~3;
      const float4* px = reinterpret_cast<const float4*>(rows[buf][0]);
      const float4* py = reinterpret_cast<const float4*>(rows[buf][1]);
      const float4* pm = reinterpret_cast<const float4*>(rows[buf][2]);
      const float4* pk = reinterpret_cast<const float4*>(rows[buf][NROW - 2]);
      const float4* pa = reinterpret_cast<const float4*>(rows[buf][NROW - 1]);
      const float2 eps2v = make_float2(a.eps2, a.eps2);
#pragma unroll 4
      for (int j4 = 0; j4 < cnt4 / 4; ++j4) {
        const float4 X = px[j4], Y = py[j4], M = pm[j4];
        float4 K = make_float4(0.f, 0.f, 0.f, 0.f), A = K;
        if (AUX) { K = pk[j4]; A = pa[j4]; }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 xj = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
          const float2 yj = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
          const float2 mj = h ? make_float2(M.z, M.w) : make_float2(M.x, M.y);
          const float2 kj = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
          const float2 aj = h ? make_float2(A.z, A.w) : make_float2(A.x, A.y);
#pragma unroll
          for (int k = 0; k < LP_IPT; ++k) {
            const float2 dx = __fadd2_rn(xj, nxi[k]);          // x_j - x_i
            const float2 dy = __fadd2_rn(yj, nyi[k]);
            if (KIND == LP_DENSITY) {
              const float2 r2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
              const float2 arg = __fmul2_rn(r2, nk[k]);
              float2 e;
              e.x = ex2_ftz(arg.x); e.y = ex2_ftz(arg.y);
              const float2 me = __fmul2_rn(mj, e);
              s0[k] = __fadd2_rn(s0[k], me);
              s1[k] = __ffma2_rn(me, r2, s1[k]);
            } else if (KIND == LP_EPSGRAD) {
              const float2 r2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
              const float2 ai = __fmul2_rn(r2, nk[k]);
              const float2 bj = __fmul2_rn(r2, kj);
              float2 ei, ej;
              ei.x = ex2_ftz(ai.x); ei.y = ex2_ftz(ai.y);
              ej.x = ex2_ftz(bj.x); ej.y = ex2_ftz(bj.y);
              const float2 tt = __fmul2_rn(mj, ei);
              const float2 uu = __fmul2_rn(aj, ej);
              s0[k] = __ffma2_rn(tt, dx, s0[k]);
              s1[k] = __ffma2_rn(tt, dy, s1[k]);
              s2[k] = __ffma2_rn(uu, dx, s2[k]);
              s3[k] = __ffma2_rn(uu, dy, s3[k]);
            } else if (KIND == LP_UNITGRAD) {
              const float2 r2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
              float2 w;
              w.x = (r2.x > 0.f) ? rsqrtf(r2.x) : 0.f;
              w.y = (r2.y > 0.f) ? rsqrtf(r2.y) : 0.f;
              const float2 c = __fmul2_rn(__fmul2_rn(mj, w), __fmul2_rn(w, w));
              s0[k] = __ffma2_rn(c, dx, s0[k]);
              s1[k] = __ffma2_rn(c, dy, s1[k]);
            } else {   // LP_TAUMIN
              const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
              const float2 r2 = __fadd2_rn(d2, eps2v);
              const float2 msum = __fadd2_rn(mj, nk[k]);
              float2 v;
              v.x = r2.x * sqrtf(r2.x) / msum.x;
              v.y = r2.y * sqrtf(r2.y) / msum.y;
              // the self pair (and tile padding, m_j = 0 at the origin) must not win the minimum
              const int jj = jt0 + 4 * j4 + 2 * h;
              const int gi = a.i0 + ii[k];
              if (jj == gi) v.x = 3.0e38f;
              if (jj + 1 == gi) v.y = 3.0e38f;
              if (jj >= jt0 + cnt) v.x = 3.0e38f;
              if (jj + 1 >= jt0 + cnt) v.y = 3.0e38f;
              s0[k].x = fminf(s0[k].x, v.x);
              s0[k].y = fminf(s0[k].y, v.y);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < LP_IPT; ++k) {
        if (KIND == LP_TAUMIN) {
          o0[k] = fmin(o0[k], (double)fminf(s0[k].x, s0[k].y));
        } else {
          o0[k] += (double)(s0[k].x + s0[k].y);
          o1[k] += (double)(s1[k].x + s1[k].y);
          if (KIND == LP_EPSGRAD) {
            o2[k] += (double)(s2[k].x + s2[k].y);
            o3[k] += (double)(s3[k].x + s3[k].y);
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LP_IPT; ++k) {
      if (ii[k] >= a.ni) continue;
      const int gi = a.i0 + ii[k];
      if (KIND == LP_DENSITY) {
        // the self pair contributed exactly m_i e^0 to S0 (and 0 to S1): remove it in the chunk that saw it
        if (gi >= j_begin && gi < j_end) o0[k] -= (double)mi[k];
        atomicAdd(&a.out[2 * (size_t)ii[k] + 0], o0[k]);
        atomicAdd(&a.out[2 * (size_t)ii[k] + 1], o1[k]);
      } else if (KIND == LP_EPSGRAD) {
        // accumulated with (q_j - q_i): flip the sign
        atomicAdd(&a.out[2 * (size_t)ii[k] + 0], -((double)Ai[k] * o0[k] + (double)mi[k] * o2[k]));
        atomicAdd(&a.out[2 * (size_t)ii[k] + 1], -((double)Ai[k] * o1[k] + (double)mi[k] * o3[k]));
      } else if (KIND == LP_UNITGRAD) {
        atomicAdd(&a.out[2 * (size_t)ii[k] + 0], -o0[k]);
        atomicAdd(&a.out[2 * (size_t)ii[k] + 1], -o1[k]);
      } else {
        // positive doubles order like their bit patterns: minimum across j-chunks with an integer atomic
        atomicMin(reinterpret_cast<unsigned long long*>(&a.out[ii[k]]), (unsigned long long)__double_as_longlong(o0[k]));
      }
    }
  }
}

__global__ void fill_f64_kernel(double* p, size_t n, double v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int largeN_tile_boxes(const float* xym, const float* jaux, int n_total, float* boxes, cudaStream_t st) {
  if (!xym || !boxes || n_total <= 0) { set_error("nb_largeN_tile_boxes_f32: bad arguments"); return NB_ERR_ARG; }
  const int n_tiles = (n_total + LP_TILE - 1) / LP_TILE;
  largeN_tile_box_kernel<<<n_tiles, 256, 0, st>>>(reinterpret_cast<const float4*>(xym), reinterpret_cast<const float2*>(jaux),
                                                  n_total, LP_TILE, reinterpret_cast<TileBox*>(boxes));
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int largeN_pass(int kind, const float* xym, const float* jaux, int n_total, int i0, int ni, const float* iparam,
                float eps, double* out, const float* boxes, cudaStream_t st) {
  if (!xym || !out || n_total <= 0 || ni <= 0 || i0 < 0 || i0 + ni > n_total || kind < 0 || kind > 3 ||
      (kind == LP_DENSITY && !iparam) || (kind == LP_EPSGRAD && !jaux)) {
    set_error("nb_largeN_pass_f32: bad arguments");
    return NB_ERR_ARG;
  }
  static const char* const names[4] = {"nb_largeN_pass: DENSITY", "nb_largeN_pass: EPSGRAD", "nb_largeN_pass: UNITGRAD", "nb_largeN_pass: TAUMIN"};
  NvtxRange r(names[kind]);
  int dev = 0, sm = 0;
  NB_CUDA_CHECK(cudaGetDevice(&dev));
  NB_CUDA_CHECK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  LargeNChunks c = largeN_chunks(n_total, ni, LP_IPT, LP_TILE, sm * 4, false);
  if ((c.jchunk + LP_TILE - 1) / LP_TILE > LP_MAX_TILES) {
    c.jchunk = LP_MAX_TILES * LP_TILE;
    c.n_jchunks = (n_total + c.jchunk - 1) / c.jchunk;
    const int n_units = c.n_ichunks * c.n_jchunks;
    c.grid = n_units < sm * 4 ? n_units : sm * 4;
  }
  PassArgs a;
  a.boxes = reinterpret_cast<const TileBox*>(boxes);
  a.xym = reinterpret_cast<const float4*>(xym);
  a.jaux = reinterpret_cast<const float2*>(jaux);
  a.iparam = iparam;
  a.n_total = n_total; a.i0 = i0; a.ni = ni;
  a.eps2 = eps * eps;
  a.out = out;
  a.n_ichunks = c.n_ichunks; a.n_jchunks = c.n_jchunks; a.jchunk = c.jchunk;
  if (kind != LP_TAUMIN) NB_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * 2 * (size_t)ni, st));
  else fill_f64_kernel<<<(ni + 255) / 256, 256, 0, st>>>(out, (size_t)ni, 1.0e300);
  switch (kind) {
    case LP_DENSITY: largeN_pass_kernel<LP_DENSITY><<<c.grid, LN_TPB, 0, st>>>(a); break;
    case LP_EPSGRAD: largeN_pass_kernel<LP_EPSGRAD><<<c.grid, LN_TPB, 0, st>>>(a); break;
    case LP_UNITGRAD: largeN_pass_kernel<LP_UNITGRAD><<<c.grid, LN_TPB, 0, st>>>(a); break;
    default: largeN_pass_kernel<LP_TAUMIN><<<c.grid, LN_TPB, 0, st>>>(a); break;
  }
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
