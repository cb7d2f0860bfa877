// largen.cu -- single large-N direct sum (C5): fp32 pair arithmetic, fp64 accumulation across j-tiles.
//
// Same formula as forces.py:63-75 (acceleration), potential.py:23-64 (U) and forces.py:77-112 (dV/deps),
// which the reference can only evaluate through dense (N,N,2) arrays (16 N^2 bytes, geometry_cache.py:30).
// Here the j-particles stream through shared memory in tiles, memory is O(N), and a rank owns the
// i-range [i0, i0+ni) so the i-blocks shard over GPUs with one position all-gather per evaluation.
//
// Work decomposition: unit = (i-block of TPB*IPT particles) x (j-chunk).  Units are dealt round-robin to a
// persistent grid sized as a multiple of the SM count, partial sums go to fp64 accumulators with
// atomicAdd, so the tail is at most one unit per CTA.
#include "largen_tile.cuh"

namespace nb {

constexpr int LN_TILE = 1024;      // v1: j-particles per shared-memory tile (16 KB of float4)

struct LargeNArgs {
  const float4* xym;
  int n_total;
  int i0;
  int ni;
  float eps2;
  float G;
  double* acc64;     // [ni][2] fp64 accumulators (zeroed by the caller wrapper)
  double* sums;      // [2]: sum m_i m_j / rho, sum m_i m_j / rho^3 over ordered pairs i != j
  int n_ichunks;
  int n_jchunks;
  int jchunk;        // j-particles per chunk (multiple of the tile size)
};

template <int IPT, bool SCALARS, bool EPS_ZERO, bool USE_TMA>
__global__ void __launch_bounds__(LN_TPB) largeN_accel_kernel(LargeNArgs a) {
  __shared__ __align__(128) float4 tile[2][LN_TILE];
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x;
  if (USE_TMA) {
    if (tid == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  uint32_t phase[2] = {0u, 0u};
  const int n_units = a.n_ichunks * a.n_jchunks;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int ic = unit / a.n_jchunks;
    const int jc = unit - ic * a.n_jchunks;
    const int j_begin = jc * a.jchunk;
    const int j_end = min(a.n_total, j_begin + a.jchunk);
    float xi[IPT], yi[IPT], mi[IPT];
    int ii[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      ii[k] = ic * (LN_TPB * IPT) + k * LN_TPB + tid;     // local index in [0, ni)
      const int gi = min(a.i0 + ii[k], a.n_total - 1);
      const float4 p = a.xym[gi];
      xi[k] = p.x; yi[k] = p.y; mi[k] = p.z;
    }
    double ax64[IPT], ay64[IPT], u64 = 0.0, s364 = 0.0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) { ax64[k] = 0.0; ay64[k] = 0.0; }

    const int n_tiles = (j_end - j_begin + LN_TILE - 1) / LN_TILE;
    // prologue: tile 0 -> buffer 0
    auto issue = [&](int t, int buf) {
      const int j0 = j_begin + t * LN_TILE;
      const int cnt = min(LN_TILE, j_end - j0);
      if (USE_TMA) {
        if (tid == 0) {
          mbar_expect_tx(&bars[buf], (uint32_t)cnt * 16u);
          tma_load_1d(&tile[buf][0], a.xym + j0, (uint32_t)cnt * 16u, &bars[buf]);
        }
      } else {
        for (int j = tid; j < cnt; j += LN_TPB) tile[buf][j] = a.xym[j0 + j];
      }
    };
    issue(0, 0);
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      if (USE_TMA) {
        if (t + 1 < n_tiles) issue(t + 1, buf ^ 1);     // buffer buf^1 was released by the barrier at the end of t-1
        mbar_wait(&bars[buf], phase[buf]);
        phase[buf] ^= 1u;
      } else {
        __syncthreads();                                // tile t visible
        if (t + 1 < n_tiles) issue(t + 1, buf ^ 1);
      }
      const int cnt = min(LN_TILE, j_end - (j_begin + t * LN_TILE));
      float ax[IPT], ay[IPT], us[IPT], s3[IPT];
#pragma unroll
      for (int k = 0; k < IPT; ++k) { ax[k] = 0.f; ay[k] = 0.f; us[k] = 0.f; s3[k] = 0.f; }
      const float4* tp = tile[buf];
      const int jt0 = j_begin + t * LN_TILE;
      // the i == j self pair contributes exactly 0 to the acceleration (dx = dy = 0) but m_i^2/eps to the scalar
      // sums, where it would swamp the physical sum; only the tile(s) that overlap this CTA's i-range check for it
      const int gi_lo = a.i0 + ic * (LN_TPB * IPT);
      const bool diag = SCALARS && (jt0 < gi_lo + LN_TPB * IPT) && (jt0 + cnt > gi_lo);
      if (!diag) {
#pragma unroll 8
        for (int j = 0; j < cnt; ++j) {
          const float4 pj = tp[j];
#pragma unroll
          for (int k = 0; k < IPT; ++k) {
            const float dx = pj.x - xi[k];
            const float dy = pj.y - yi[k];
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, a.eps2));
            float w = rsqrtf(r2);
            if (EPS_ZERO) w = (r2 > 0.f) ? w : 0.f;
            const float mw = pj.z * w;
            const float c = mw * (w * w);
            ax[k] = fmaf(c, dx, ax[k]);
            ay[k] = fmaf(c, dy, ay[k]);
            if (SCALARS) { us[k] += mw; s3[k] += c; }
          }
        }
      } else {
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
          const float4 pj = tp[j];
#pragma unroll
          for (int k = 0; k < IPT; ++k) {
            const float dx = pj.x - xi[k];
            const float dy = pj.y - yi[k];
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, a.eps2));
            float w = rsqrtf(r2);
            if (EPS_ZERO) w = (r2 > 0.f) ? w : 0.f;
            if (jt0 + j == a.i0 + ii[k]) w = 0.f;
            const float mw = pj.z * w;
            const float c = mw * (w * w);
            ax[k] = fmaf(c, dx, ax[k]);
            ay[k] = fmaf(c, dy, ay[k]);
            us[k] += mw; s3[k] += c;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < IPT; ++k) {
        ax64[k] += (double)ax[k];
        ay64[k] += (double)ay[k];
        if (SCALARS) { u64 += (double)(mi[k] * us[k]) * (ii[k] < a.ni ? 1.0 : 0.0); s364 += (double)(mi[k] * s3[k]) * (ii[k] < a.ni ? 1.0 : 0.0); }
      }
      __syncthreads();                                  // everyone done with buffer buf before it is refilled
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      if (ii[k] < a.ni) {
        atomicAdd(&a.acc64[2 * (size_t)ii[k] + 0], (double)a.G * ax64[k]);
        atomicAdd(&a.acc64[2 * (size_t)ii[k] + 1], (double)a.G * ay64[k]);
      }
    }
    if (SCALARS) {
      // block reduction of the two scalars
      __shared__ double red[2][LN_TPB / 32];
      for (int off = 16; off > 0; off >>= 1) {
        u64 += __shfl_down_sync(0xffffffffu, u64, off);
        s364 += __shfl_down_sync(0xffffffffu, s364, off);
      }
      if ((tid & 31) == 0) { red[0][tid >> 5] = u64; red[1][tid >> 5] = s364; }
      __syncthreads();
      if (tid == 0) {
        double su = 0.0, s3s = 0.0;
        for (int w = 0; w < LN_TPB / 32; ++w) { su += red[0][w]; s3s += red[1][w]; }
        atomicAdd(&a.sums[0], su);
        atomicAdd(&a.sums[1], s3s);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// v2: packed f32x2 arithmetic over PAIRS OF j-PARTICLES (Blackwell FFMA2 / FADD2 / FMUL2).
//
// v1 is issue-bound: ~13 issue slots per pair (4 FFMA + 3.5 FMUL + 2 FADD + MUFU + the denormal fix-up
// of rsqrtf).  Here a TMA bulk copy lands the AoS (x, y, m, 0) tile in `raw`, the CTA transposes it
// once into x[] / y[] / m[] rows, and every LDS.128 of a row then yields two aligned register pairs
// (x_j, x_j+1), (x_j+2, x_j+3).  Per TWO pairs: 2 FADD2, 4 FFMA2, 3 FMUL2, 2 MUFU.RSQ (ftz, no fix-up)
// = 5.5 issue slots per pair; the i-particle operands are loop-invariant (x_i, x_i) register pairs.
// Padding entries (m = 0 at the origin) contribute exactly 0.
// ------------------------------------------------------------------------------------------------
constexpr int LN2_TILE = 512;      // j-particles per tile: raw 8 KB + rows 6 KB, double buffered = 28 KB

template <int IPT, int MINB, bool SCALARS, bool EPS_ZERO>
__global__ void __launch_bounds__(LN_TPB, MINB) largeN_accel_x2_kernel(LargeNArgs a) {
  __shared__ __align__(128) float4 raw[2][LN2_TILE];
  __shared__ __align__(16) float sx[2][LN2_TILE];
  __shared__ __align__(16) float sy[2][LN2_TILE];
  __shared__ __align__(16) float sm[2][LN2_TILE];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ double red[2][LN_TPB / 32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase[2] = {0u, 0u};
  const int n_units = a.n_ichunks * a.n_jchunks;
  const float2 eps2v = make_float2(a.eps2, a.eps2);
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int ic = unit / a.n_jchunks;
    const int jc = unit - ic * a.n_jchunks;
    const int j_begin = jc * a.jchunk;
    const int j_end = min(a.n_total, j_begin + a.jchunk);
    float2 nxi[IPT], nyi[IPT];
    float mi[IPT];
    int ii[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      ii[k] = ic * (LN_TPB * IPT) + k * LN_TPB + tid;
      const int gi = min(a.i0 + ii[k], a.n_total - 1);
      const float4 p = a.xym[gi];
      nxi[k] = make_float2(-p.x, -p.x);
      nyi[k] = make_float2(-p.y, -p.y);
      mi[k] = p.z;
    }
    double ax64[IPT], ay64[IPT], u64 = 0.0, s364 = 0.0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) { ax64[k] = 0.0; ay64[k] = 0.0; }

    const int n_tiles = (j_end - j_begin + LN2_TILE - 1) / LN2_TILE;
    auto issue = [&](int t, int buf) {
      if (tid == 0) {
        const int j0 = j_begin + t * LN2_TILE;
        const int cnt = min(LN2_TILE, j_end - j0);
        mbar_expect_tx(&bars[buf], (uint32_t)cnt * 16u);
        tma_load_1d(&raw[buf][0], a.xym + j0, (uint32_t)cnt * 16u, &bars[buf]);
      }
    };
    issue(0, 0);
    if (n_tiles > 1) issue(1, 1);
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int jt0 = j_begin + t * LN2_TILE;
      const int cnt = min(LN2_TILE, j_end - jt0);
      mbar_wait(&bars[buf], phase[buf]);
      phase[buf] ^= 1u;
      // transpose AoS -> rows (conflict-free: consecutive threads, consecutive float4 / floats)
#pragma unroll
      for (int r = 0; r < LN2_TILE / LN_TPB; ++r) {
        const int j = r * LN_TPB + tid;
        float4 p = raw[buf][j];
        if (j >= cnt) p = make_float4(0.f, 0.f, 0.f, 0.f);
        sx[buf][j] = p.x; sy[buf][j] = p.y; sm[buf][j] = p.z;
      }
      __syncthreads();   // rows[buf] visible; raw[buf] consumed by all; everyone finished tile t-1 (rows[buf^1])
      if (t + 2 < n_tiles) issue(t + 2, buf);

      float2 ax[IPT], ay[IPT], us[IPT], s3[IPT];
#pragma unroll
      for (int k = 0; k < IPT; ++k) {
        ax[k] = make_float2(0.f, 0.f); ay[k] = ax[k]; us[k] = ax[k]; s3[k] = ax[k];
      }
      const int gi_lo = a.i0 + ic * (LN_TPB * IPT);
      const bool diag = SCALARS && (jt0 < gi_lo + LN_TPB * IPT) && (jt0 + cnt > gi_lo);
      const int cnt4 = (cnt + 3) & ~3;
      const float4* px = reinterpret_cast<const float4*>(sx[buf]);
      const float4* py = reinterpret_cast<const float4*>(sy[buf]);
      const float4* pm = reinterpret_cast<const float4*>(sm[buf]);
      if (!diag) {
#pragma unroll 4
        for (int j4 = 0; j4 < cnt4 / 4; ++j4) {
          const float4 X = px[j4], Y = py[j4], M = pm[j4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 xj = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
            const float2 yj = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
            const float2 mj = h ? make_float2(M.z, M.w) : make_float2(M.x, M.y);
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
              const float2 dx = __fadd2_rn(xj, nxi[k]);
              const float2 dy = __fadd2_rn(yj, nyi[k]);
              const float2 r2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, eps2v));
              float2 w;
              if (EPS_ZERO) {
                w.x = (r2.x > 0.f) ? rsqrtf(r2.x) : 0.f;
                w.y = (r2.y > 0.f) ? rsqrtf(r2.y) : 0.f;
              } else {
                w.x = rsqrt_ftz(r2.x);
                w.y = rsqrt_ftz(r2.y);
              }
              const float2 mw = __fmul2_rn(mj, w);
              const float2 c = __fmul2_rn(mw, __fmul2_rn(w, w));
              ax[k] = __ffma2_rn(c, dx, ax[k]);
              ay[k] = __ffma2_rn(c, dy, ay[k]);
              if (SCALARS) { us[k] = __fadd2_rn(us[k], mw); s3[k] = __fadd2_rn(s3[k], c); }
            }
          }
        }
      } else {
        // the tile overlaps this CTA's own i-range: scalar loop with the exact i == j exclusion
        for (int j = 0; j < cnt; ++j) {
          const float xj = sx[buf][j], yj = sy[buf][j], mj = sm[buf][j];
#pragma unroll
          for (int k = 0; k < IPT; ++k) {
            const float dx = xj + nxi[k].x;
            const float dy = yj + nyi[k].x;
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, a.eps2));
            float w = (r2 > 0.f) ? rsqrtf(r2) : 0.f;
            if (jt0 + j == a.i0 + ii[k]) w = 0.f;
            const float mw = mj * w;
            const float c = mw * (w * w);
            ax[k].x = fmaf(c, dx, ax[k].x);
            ay[k].x = fmaf(c, dy, ay[k].x);
            us[k].x += mw; s3[k].x += c;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < IPT; ++k) {
        ax64[k] += (double)(ax[k].x + ax[k].y);
        ay64[k] += (double)(ay[k].x + ay[k].y);
        if (SCALARS) {
          const double live = ii[k] < a.ni ? 1.0 : 0.0;
          u64 += (double)(mi[k] * (us[k].x + us[k].y)) * live;
          s364 += (double)(mi[k] * (s3[k].x + s3[k].y)) * live;
        }
      }
    }
    __syncthreads();     // all threads are past the last tile before the next unit re-arms the buffers
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      if (ii[k] < a.ni) {
        atomicAdd(&a.acc64[2 * (size_t)ii[k] + 0], (double)a.G * ax64[k]);
        atomicAdd(&a.acc64[2 * (size_t)ii[k] + 1], (double)a.G * ay64[k]);
      }
    }
    if (SCALARS) {
      for (int off = 16; off > 0; off >>= 1) {
        u64 += __shfl_down_sync(0xffffffffu, u64, off);
        s364 += __shfl_down_sync(0xffffffffu, s364, off);
      }
      if ((tid & 31) == 0) { red[0][tid >> 5] = u64; red[1][tid >> 5] = s364; }
      __syncthreads();
      if (tid == 0) {
        double su = 0.0, s3s = 0.0;
        for (int w = 0; w < LN_TPB / 32; ++w) { su += red[0][w]; s3s += red[1][w]; }
        atomicAdd(&a.sums[0], su);
        atomicAdd(&a.sums[1], s3s);
      }
      __syncthreads();
    }
  }
}

__global__ void largeN_finish_kernel(const double* __restrict__ acc64, int ni, float2* acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ni) acc[i] = make_float2((float)acc64[2 * (size_t)i], (float)acc64[2 * (size_t)i + 1]);
}

// v += kick_h * a ; q += drift_h * v ; writes the updated (x, y, m, 0) back into the packed gather buffer slice
__global__ void largeN_kick_drift_kernel(float4* __restrict__ xym_local, float2* __restrict__ vel,
                                         const float2* __restrict__ acc, int ni, float kick_h, float drift_h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ni) return;
  float2 v = vel[i];
  if (kick_h != 0.f) {
    const float2 a = acc[i];
    v.x = fmaf(kick_h, a.x, v.x);
    v.y = fmaf(kick_h, a.y, v.y);
    vel[i] = v;
  }
  if (drift_h != 0.f) {
    float4 p = xym_local[i];
    p.x = fmaf(drift_h, v.x, p.x);
    p.y = fmaf(drift_h, v.y, p.y);
    xym_local[i] = p;
  }
}

// kernel variants (tuning / A-B tests): 0..7 = v1 scalar kernel (bit0 TMA, bits1-2 IPT selector 4/2/1);
// 8 = v2 packed f32x2, IPT 4, 4 CTAs/SM; 9 = v2 IPT 8, 2 CTAs/SM; 10 = v2 IPT 2, 4 CTAs/SM (default, -1)
int largeN_accel(const float* xym, int n_total, int i0, int ni, float eps, float G, float* acc, double* sums,
                 double* workspace, int variant_in, cudaStream_t st) {
  if (!xym || !acc || !workspace || n_total <= 0 || ni <= 0 || i0 < 0 || i0 + ni > n_total || variant_in < -1 || variant_in > 10) {
    set_error("nb_largeN_accel_f32: bad arguments (workspace = ni x 2 doubles of device memory is required; variant -1..10)");
    return NB_ERR_ARG;
  }
  NvtxRange r("nb_largeN_accel");
  // stateless: the fp64 accumulators live in the CALLER's workspace, so calls on different streams / devices / threads
  // never share anything (round 1 kept one process-global buffer here)
  int dev = 0, sm_count = 0;
  NB_CUDA_CHECK(cudaGetDevice(&dev));
  NB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  double* const acc64 = workspace;
  // default (-1): 2 i-particles per thread, 4 CTAs/SM, 16 j-particles per unrolled iteration: measured 3.25e12 pairs/s
  // (4 or 8 i per thread: 3.13-3.17e12; with the fused scalar sums their extra accumulators spill at 64 registers)
  const int variant = variant_in >= 0 ? variant_in : 10;
  const bool v2 = variant >= 8;
  const int v2_ipt = variant == 9 ? 8 : (variant == 10 ? 2 : 4);
  const int v2_minb = variant == 9 ? 2 : 4;
  NB_CUDA_CHECK(cudaMemsetAsync(acc64, 0, sizeof(double) * 2 * (size_t)ni, st));
  const bool use_tma = (variant & 1) != 0;
  const int ipt = v2 ? v2_ipt : (((variant >> 1) & 3) == 0 ? 4 : (((variant >> 1) & 3) == 1 ? 2 : 1));
  const int tile = v2 ? LN2_TILE : LN_TILE;
  LargeNArgs a;
  a.xym = reinterpret_cast<const float4*>(xym);
  a.n_total = n_total;
  a.i0 = i0;
  a.ni = ni;
  a.eps2 = eps * eps;
  a.G = G;
  a.acc64 = acc64;
  a.sums = sums;
  const int per_block = LN_TPB * ipt;
  a.n_ichunks = (ni + per_block - 1) / per_block;
  // enough j-chunks that the persistent grid gets >= ~12 rounds of units, but each chunk >= 8 tiles
  const int resident = sm_count * (v2 ? v2_minb : 4);
  int n_j = (12 * resident + a.n_ichunks - 1) / a.n_ichunks;
  const int max_j = (n_total + 8 * tile - 1) / (8 * tile);
  n_j = n_j < 1 ? 1 : (n_j > max_j ? max_j : n_j);
  int jchunk = (n_total + n_j - 1) / n_j;
  jchunk = ((jchunk + tile - 1) / tile) * tile;
  a.jchunk = jchunk;
  a.n_jchunks = (n_total + jchunk - 1) / jchunk;
  const int n_units = a.n_ichunks * a.n_jchunks;
  const int grid = n_units < resident ? n_units : resident;
  // decided on eps^2, which is what the kernel adds: an eps whose square flushes to zero must take the guarded path
  const bool eps_zero = !(eps * eps > 1.17549435e-38f);
  const bool scal = sums != nullptr;
#define NB_LN_LAUNCH(IPT, SC, EZ, TMA) largeN_accel_kernel<IPT, SC, EZ, TMA><<<grid, LN_TPB, 0, st>>>(a)
#define NB_LN_SWITCH(IPT)                                                          \
  do {                                                                             \
    if (use_tma) {                                                                 \
      if (scal) { if (eps_zero) NB_LN_LAUNCH(IPT, true, true, true); else NB_LN_LAUNCH(IPT, true, false, true); } \
      else { if (eps_zero) NB_LN_LAUNCH(IPT, false, true, true); else NB_LN_LAUNCH(IPT, false, false, true); }    \
    } else {                                                                       \
      if (scal) { if (eps_zero) NB_LN_LAUNCH(IPT, true, true, false); else NB_LN_LAUNCH(IPT, true, false, false); } \
      else { if (eps_zero) NB_LN_LAUNCH(IPT, false, true, false); else NB_LN_LAUNCH(IPT, false, false, false); }    \
    }                                                                              \
  } while (0)
#define NB_LN2_LAUNCH(IPT, MINB)                                                                              \
  do {                                                                                                       \
    if (scal) { if (eps_zero) largeN_accel_x2_kernel<IPT, MINB, true, true><<<grid, LN_TPB, 0, st>>>(a);     \
                else largeN_accel_x2_kernel<IPT, MINB, true, false><<<grid, LN_TPB, 0, st>>>(a); }           \
    else { if (eps_zero) largeN_accel_x2_kernel<IPT, MINB, false, true><<<grid, LN_TPB, 0, st>>>(a);         \
           else largeN_accel_x2_kernel<IPT, MINB, false, false><<<grid, LN_TPB, 0, st>>>(a); }               \
  } while (0)
  if (v2) {
    if (v2_ipt == 8) NB_LN2_LAUNCH(8, 2);
    else if (v2_ipt == 2) NB_LN2_LAUNCH(2, 4);
    else NB_LN2_LAUNCH(4, 4);
  } else if (ipt == 4) NB_LN_SWITCH(4);
  else if (ipt == 2) NB_LN_SWITCH(2);
  else NB_LN_SWITCH(1);
#undef NB_LN2_LAUNCH
#undef NB_LN_SWITCH
#undef NB_LN_LAUNCH
  NB_CUDA_CHECK(cudaGetLastError());
  largeN_finish_kernel<<<(ni + 255) / 256, 256, 0, st>>>(acc64, ni, reinterpret_cast<float2*>(acc));
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

int largeN_kick_drift(float* xym_local, float* vel, const float* acc, int ni, float kick_h, float drift_h,
                      cudaStream_t st) {
  if (!xym_local || !vel || ni <= 0 || (kick_h != 0.f && !acc)) {
    set_error("nb_largeN_kick_drift_f32: bad arguments");
    return NB_ERR_ARG;
  }
  largeN_kick_drift_kernel<<<(ni + 255) / 256, 256, 0, st>>>(reinterpret_cast<float4*>(xym_local),
                                                            reinterpret_cast<float2*>(vel),
                                                            reinterpret_cast<const float2*>(acc), ni, kick_h, drift_h);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
