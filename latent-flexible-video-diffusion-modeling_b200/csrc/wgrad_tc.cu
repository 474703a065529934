// tcgen05 weight-gradient kernel (placeholder translation unit: the CUDA-core engine in bwd_conv.cu takes every shape until
// the tensor-core kernel lands).
#include "common.cuh"

namespace fdm {
int conv_wgrad_tc(const fdm_conv_wgrad_args* a, cudaStream_t st) {
  (void)a;
  (void)st;
  return FDM_ERR_UNSUPPORTED;
}
}  // namespace fdm
