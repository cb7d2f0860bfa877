// classifier.cu -- stability-classifier inference fused onto the feature tensors.
//
// The reference trains an MLP (model_zoo.py:18-33: F -> 128 -> 64 -> 1, ReLU, dropout inactive at inference) on the
// StandardScaler-ed feature table (train_mlp.py:44-60, 141-217) and thresholds sigmoid(logit).  This kernel reads the
// fp64 feature tensors the ensemble kernels wrote (dyn_features[B][22], static_features[B][25]) through a gather map,
// applies nan_to_num + (x - mean) / scale (stability_dataset.py:83-85), the three layers, the sigmoid and the
// threshold in one launch: the feature table never leaves HBM and no per-row Python object is built.
// fp32 CUDA-core arithmetic like the reference's torch model; a 64-row tile per CTA, weights staged once in shared
// memory, register-tiled 4x8 / 4x4 outer products.  ~14 kFMA per row: 1e6 systems cost a few milliseconds, three
// orders of magnitude below the integration that produced the features, so tensor cores are not warranted.
#include "common.cuh"

namespace nb {

constexpr int MLP_H1 = 128, MLP_H2 = 64, MLP_ROWS = 64, MLP_THREADS = 256, MLP_MAXF = 64;

struct MlpArgs {
  const double* dyn;      // [B][NB_N_DYN]
  const double* stat;     // [B][NB_N_STATIC] or null
  const int32_t* idx;     // [F]: < 63 -> dyn column, 63 -> pathological_energy (derived), >= 64 -> static column idx - 64
  const float* mean;      // [F]
  const float* inv_scale; // [F]
  const float* w1;        // [F][128]  (input-major)
  const float* b1;        // [128]
  const float* w2;        // [128][64] (input-major)
  const float* b2;        // [64]
  const float* w3;        // [64]
  float b3, threshold;
  int B, F;
  float* prob;            // [B]
  int32_t* label;         // [B]
};

__global__ void __launch_bounds__(MLP_THREADS) mlp_classify_kernel(MlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* Xs = sm;                                  // [64][F]
  float* W1s = Xs + MLP_ROWS * MLP_MAXF;           // [F][128]
  float* H1s = W1s + MLP_MAXF * MLP_H1;            // [64][128]
  float* W2s = H1s + MLP_ROWS * MLP_H1;            // [128][64]
  const int tid = threadIdx.x;
  const int F = a.F;
  for (int i = tid; i < F * MLP_H1; i += MLP_THREADS) W1s[i] = a.w1[i];
  for (int i = tid; i < MLP_H1 * MLP_H2; i += MLP_THREADS) W2s[i] = a.w2[i];
  const int ty = tid >> 4, tx = tid & 15;          // 16 x 16 thread grid: rows 4 ty .. 4 ty + 3
  for (int tile = blockIdx.x; tile * MLP_ROWS < a.B; tile += gridDim.x) {
    const int row0 = tile * MLP_ROWS;
    __syncthreads();
    for (int i = tid; i < MLP_ROWS * F; i += MLP_THREADS) {
      const int r = i / F, k = i - r * F;
      const int row = min(row0 + r, a.B - 1);
      const int c = a.idx[k];
      double v;
      if (c == NB_MLP_COL_PATHOLOGICAL) {
        // the dataset writer's derived column: pathological_energy = |energy_drift| > 10 (batch_stability_analyzer.py:45-52)
        v = fabs(a.dyn[(size_t)row * NB_N_DYN + NB_F_ENERGY_DRIFT]) > 10.0 ? 1.0 : 0.0;
      } else {
        v = c < 64 ? a.dyn[(size_t)row * NB_N_DYN + c] : a.stat[(size_t)row * NB_N_STATIC + (c - 64)];
      }
      if (v != v) v = 0.0;                                         // nan_to_num(nan=0.0); +-inf stay as in numpy
      Xs[r * F + k] = ((float)v - a.mean[k]) * a.inv_scale[k];
    }
    __syncthreads();
    // ---- fc1 + ReLU: 4 rows x 8 columns per thread
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = a.b1[tx + 16 * c];
    for (int k = 0; k < F; ++k) {
      float xa[4], wb[8];
#pragma unroll
      for (int r = 0; r < 4; ++r) xa[r] = Xs[(4 * ty + r) * F + k];
#pragma unroll
      for (int c = 0; c < 8; ++c) wb[c] = W1s[k * MLP_H1 + tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(xa[r], wb[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) H1s[(4 * ty + r) * MLP_H1 + tx + 16 * c] = fmaxf(acc[r][c], 0.f);
    __syncthreads();
    // ---- fc2 + ReLU: 4 rows x 4 columns per thread
    float h2[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) h2[r][c] = a.b2[tx + 16 * c];
#pragma unroll 4
    for (int k = 0; k < MLP_H1; ++k) {
      float ha[4], wb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) ha[r] = H1s[(4 * ty + r) * MLP_H1 + k];
#pragma unroll
      for (int c = 0; c < 4; ++c) wb[c] = W2s[k * MLP_H2 + tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) h2[r][c] = fmaf(ha[r], wb[c], h2[r][c]);
    }
    // ---- fc3 + sigmoid + threshold: partial dot over this thread's 4 columns, reduced over the 16 tx lanes
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) s = fmaf(fmaxf(h2[r][c], 0.f), a.w3[tx + 16 * c], s);
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      const int row = row0 + 4 * ty + r;
      if (tx == 0 && row < a.B) {
        const float logit = s + a.b3;
        const float p = 1.f / (1.f + expf(-logit));
        a.prob[row] = p;
        if (a.label) a.label[row] = p > a.threshold ? 1 : 0;
      }
    }
  }
}

int mlp_classify(const double* dyn, const double* stat, const int32_t* idx, int F, const float* mean,
                 const float* inv_scale, const float* w1, const float* b1, const float* w2, const float* b2,
                 const float* w3, float b3, float threshold, int B, float* prob, int32_t* label, cudaStream_t st) {
  if (!dyn || !idx || !mean || !inv_scale || !w1 || !b1 || !w2 || !b2 || !w3 || !prob || F < 1 || F > MLP_MAXF || B < 0) {
    set_error("nb_mlp_classify_f32: bad arguments (1 <= F <= 64)");
    return NB_ERR_ARG;
  }
  if (B == 0) return NB_OK;
  const size_t smem = sizeof(float) * (MLP_ROWS * MLP_MAXF + MLP_MAXF * MLP_H1 + MLP_ROWS * MLP_H1 + MLP_H1 * MLP_H2);
  // the opt-in above 48 KB is per DEVICE: set it on every call (a cheap host-side call) instead of once per process
  NB_CUDA_CHECK(cudaFuncSetAttribute(mlp_classify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MlpArgs a{dyn, stat, idx, mean, inv_scale, w1, b1, w2, b2, w3, b3, threshold, B, F, prob, label};
  const int tiles = (B + MLP_ROWS - 1) / MLP_ROWS;
  const int grid = tiles < 148 * 2 ? tiles : 148 * 2;
  mlp_classify_kernel<<<grid, MLP_THREADS, smem, st>>>(a);
  NB_CUDA_CHECK(cudaGetLastError());
  return NB_OK;
}

}  // namespace nb
