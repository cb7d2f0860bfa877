// ensemble_whfast.cu -- instantiates the persistent ensemble kernel for integrator_mode="whfast", N = 2..8.
#include "ensemble_run.cuh"
namespace nb {
int ensemble_run_whfast(const RunArgs& a, int N, cudaStream_t st) { return launch_run_n<NB_MODE_WHFAST>(a, N, st); }
}  // namespace nb
