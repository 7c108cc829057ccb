// tcgen05 tensor-core path of the word-region loss (TGFR_PREC_TC) -- placeholder until the
// kernel lands: fails loudly instead of silently using another path.
#include "common.cuh"
namespace tgfr {
size_t wordregion_tc_workspace_bytes(int, int, int, int, int) { return 0; }
int wordregion_fwd_tc(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t,
                      const int32_t*, int, int, int, int, int, float, float, float, float, float*, void*, size_t,
                      cudaStream_t) {
  set_error("wordregion: TGFR_PREC_TC is not built into this library");
  return TGFR_E_INVALID;
}
}  // namespace tgfr
