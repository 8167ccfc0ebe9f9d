// MMB likelihood step kernels (heads, masked Gaussian log-likelihood, angular word term).
// Placeholders: the entry points exist so that the ABI is complete; they report
// MMB_E_UNSUPPORTED until the kernels land.
#include "common.cuh"

using namespace mmb;

#define MMB_TODO(name)                               \
  do {                                               \
    set_error(name ": not implemented in this build"); \
    return MMB_E_UNSUPPORTED;                        \
  } while (0)

extern "C" int mmb_heads_forward(const float*, int, int, int, const float* const*, const float* const*,
                                 const int*, const int*, float* const*, int, mmb_stream_t) {
  MMB_TODO("mmb_heads_forward");
}
extern "C" int mmb_gauss_ll(const float* const*, const float* const*, const int*, int, int, int, const int*,
                            const float* const*, const float* const*, float*, float* const*, float* const*,
                            int*, mmb_stream_t) {
  MMB_TODO("mmb_gauss_ll");
}
extern "C" int mmb_row_inv_norm(const float*, int64_t, int, float*, mmb_stream_t) {
  MMB_TODO("mmb_row_inv_norm");
}
extern "C" size_t mmb_word_ll_workspace_bytes(int, int64_t, int) { return 0; }
extern "C" int mmb_word_ll(const float*, int, int, const float*, const float*, int64_t, const float*, int64_t,
                           int64_t, const float*, const float*, int64_t, int64_t, int, float, float*, float*,
                           void*, size_t, int*, mmb_stream_t) {
  MMB_TODO("mmb_word_ll");
}
