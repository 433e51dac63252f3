// tcgen05 implicit-GEMM convolution (placeholder until the kernel lands in this file).
#include "conv.h"

namespace hesic {
bool conv_tc_supported(const hesic_conv *, const hesic_tensor *, const hesic_tensor *) { return false; }
int conv_forward_tc(hesic_conv *, const hesic_tensor *, const hesic_tensor *, int, cudaStream_t) {
  set_error("tcgen05 path not built");
  return HESIC_E_UNSUPPORTED;
}
}  // namespace hesic
