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
#include "common.cuh"

namespace nb {

constexpr int LN_TPB = 256;        // threads per CTA
constexpr int LN_TILE = 1024;      // j-particles per shared-memory tile (16 KB of float4)

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
  int jchunk;        // j-particles per chunk (multiple of LN_TILE)
};

// ---- TMA 1-D bulk copy helpers (cp.async.bulk + mbarrier) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

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

static int g_ln_variant = -1;   // -1 default; set via NB_LARGEN_VARIANT env: bit0 TMA, bits1-2 IPT selector
static int g_sm_count = 0;
static double* g_acc64 = nullptr;
static size_t g_acc64_cap = 0;
static int g_acc64_dev = -1;

int largeN_accel(const float* xym, int n_total, int i0, int ni, float eps, float G, float* acc, double* sums,
                 cudaStream_t st) {
  if (!xym || !acc || n_total <= 0 || ni <= 0 || i0 < 0 || i0 + ni > n_total) {
    set_error("nb_largeN_accel_f32: bad arguments");
    return NB_ERR_ARG;
  }
  int dev = 0;
  NB_CUDA_CHECK(cudaGetDevice(&dev));
  if (g_sm_count == 0 || g_acc64_dev != dev) {
    NB_CUDA_CHECK(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  if (g_acc64_dev != dev || g_acc64_cap < (size_t)ni * 2) {
    if (g_acc64) cudaFree(g_acc64);
    g_acc64 = nullptr;
    NB_CUDA_CHECK(cudaMalloc(&g_acc64, sizeof(double) * 2 * (size_t)ni));
    g_acc64_cap = (size_t)ni * 2;
    g_acc64_dev = dev;
  }
  if (g_ln_variant < 0) {
    const char* e = getenv("NB_LARGEN_VARIANT");
    g_ln_variant = e ? atoi(e) : 1;
  }
  NB_CUDA_CHECK(cudaMemsetAsync(g_acc64, 0, sizeof(double) * 2 * (size_t)ni, st));
  const bool use_tma = (g_ln_variant & 1) != 0;
  const int ipt = ((g_ln_variant >> 1) & 3) == 0 ? 4 : (((g_ln_variant >> 1) & 3) == 1 ? 2 : 1);
  LargeNArgs a;
  a.xym = reinterpret_cast<const float4*>(xym);
  a.n_total = n_total;
  a.i0 = i0;
  a.ni = ni;
  a.eps2 = eps * eps;
  a.G = G;
  a.acc64 = g_acc64;
  a.sums = sums;
  const int per_block = LN_TPB * ipt;
  a.n_ichunks = (ni + per_block - 1) / per_block;
  // enough j-chunks that the persistent grid gets >= ~12 rounds of units, but each chunk >= 8 tiles
  const int resident = g_sm_count * 4;
  int n_j = (12 * resident + a.n_ichunks - 1) / a.n_ichunks;
  const int max_j = (n_total + 8 * LN_TILE - 1) / (8 * LN_TILE);
  n_j = n_j < 1 ? 1 : (n_j > max_j ? max_j : n_j);
  int jchunk = (n_total + n_j - 1) / n_j;
  jchunk = ((jchunk + LN_TILE - 1) / LN_TILE) * LN_TILE;
  a.jchunk = jchunk;
  a.n_jchunks = (n_total + jchunk - 1) / jchunk;
  const int n_units = a.n_ichunks * a.n_jchunks;
  const int grid = n_units < resident ? n_units : resident;
  const bool eps_zero = !(eps > 0.f);
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
  if (ipt == 4) NB_LN_SWITCH(4);
  else if (ipt == 2) NB_LN_SWITCH(2);
  else NB_LN_SWITCH(1);
#undef NB_LN_SWITCH
#undef NB_LN_LAUNCH
  NB_CUDA_CHECK(cudaGetLastError());
  largeN_finish_kernel<<<(ni + 255) / 256, 256, 0, st>>>(g_acc64, ni, reinterpret_cast<float2*>(acc));
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
