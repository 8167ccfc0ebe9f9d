// G = X^T X on the 5th-generation tensor cores (tcgen05, kind::tf32, 3xTF32 split).
// Placeholder until the tcgen05 kernel lands: reports "unsupported" so MMB_GRAM_AUTO uses
// the FP32 CUDA-core kernel.
#include "common.cuh"

namespace mmb {

bool gram_tc_supported(int64_t, int) { return false; }
size_t gram_tc_workspace_bytes(int64_t, int) { return 0; }
int gram_tc(const float*, int64_t, int, float*, void*, size_t, cudaStream_t) {
  set_error("gram_tc: not built");
  return MMB_E_UNSUPPORTED;
}

}  // namespace mmb
