// ensemble_yoshida4.cu -- instantiates the persistent ensemble kernels for integrator_mode="yoshida4", N = 2..8.
#include "ensemble_run.cuh"
namespace nb {
int ensemble_run_yoshida4(const RunArgs& a, int N, int phase, int write_state, cudaStream_t st) {
  return launch_run_n<NB_MODE_YOSHIDA4>(a, N, phase, write_state, st);
}
}  // namespace nb
