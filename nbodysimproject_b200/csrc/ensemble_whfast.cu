// ensemble_whfast.cu -- instantiates the persistent ensemble kernels for integrator_mode="whfast", N = 2..8.
#include "ensemble_run.cuh"
namespace nb {
int ensemble_run_whfast(const RunArgs& a, int N, int phase, int write_state, cudaStream_t st) {
  return launch_run_n<NB_MODE_WHFAST>(a, N, phase, write_state, st);
}
}  // namespace nb
