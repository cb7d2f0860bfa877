// hamsoft.cu -- ham_soft Strang-split ensemble kernel (placeholder until the kernel lands).
#include "common.cuh"
namespace nb {
int hamsoft_run(const double*, double*, double*, double, int, int, unsigned, double, int, int, int, const int32_t*,
                const int32_t*, const double*, const double*, double*, const double*, double*, int32_t*, cudaStream_t) {
  set_error("ham_soft ensemble kernel not built");
  return NB_ERR_UNSUPPORTED;
}
}  // namespace nb
