// largen_tile.cuh -- TMA 1-D bulk copy + mbarrier helpers shared by the large-N kernels.
#pragma once
#include "common.cuh"

namespace nb {

constexpr int LN_TPB = 256;        // threads per CTA (all large-N kernels)

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

__device__ __forceinline__ float rsqrt_ftz(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// work decomposition shared by the launchers: unit = (i-block of LN_TPB*ipt particles) x (j-chunk)
struct LargeNChunks {
  int n_ichunks, n_jchunks, jchunk, grid;
};
inline LargeNChunks largeN_chunks(int n_total, int ni, int ipt, int tile, int resident, bool single_jchunk) {
  LargeNChunks c;
  const int per_block = LN_TPB * ipt;
  c.n_ichunks = (ni + per_block - 1) / per_block;
  // enough j-chunks that the persistent grid gets >= ~12 rounds of units, but each chunk >= 8 tiles
  int n_j = (12 * resident + c.n_ichunks - 1) / c.n_ichunks;
  const int max_j = (n_total + 8 * tile - 1) / (8 * tile);
  n_j = n_j < 1 ? 1 : (n_j > max_j ? max_j : n_j);
  if (single_jchunk) n_j = 1;
  int jchunk = (n_total + n_j - 1) / n_j;
  jchunk = ((jchunk + tile - 1) / tile) * tile;
  c.jchunk = jchunk;
  c.n_jchunks = (n_total + jchunk - 1) / jchunk;
  const int n_units = c.n_ichunks * c.n_jchunks;
  c.grid = n_units < resident ? n_units : resident;
  return c;
}

}  // namespace nb
