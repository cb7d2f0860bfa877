// ensemble_verlet.cu -- instantiates the persistent ensemble kernels for integrator_mode="verlet", N = 2..8.
#include "ensemble_run.cuh"
namespace nb {
int ensemble_run_verlet(const RunArgs& a, int N, int phase, int write_state, cudaStream_t st) {
  return launch_run_n<NB_MODE_VERLET>(a, N, phase, write_state, st);
}
}  // namespace nb
