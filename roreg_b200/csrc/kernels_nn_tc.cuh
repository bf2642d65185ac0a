// kernels_nn_tc.cuh - tensor-core (tcgen05, 3xTF32) mutual-NN Gram path.  Placeholder until the
// tcgen05 kernel lands: mode 1 reports ROREG_ERR_UNSUPPORTED instead of silently falling back.
#pragma once
#include "common.cuh"
#include "kernels_match.cuh"
namespace roreg {
static inline int nn_tc_launch(roreg_ctx* c, const NNArgs&, int, cudaStream_t) {
  snprintf(c->err, sizeof(c->err), "nn mode 1 (tcgen05 Gram) not built in this library");
  return ROREG_ERR_UNSUPPORTED;
}
}  // namespace roreg
