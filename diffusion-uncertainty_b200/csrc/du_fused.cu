// du_fused.cu — the fused uncertainty step (placeholder until the cluster kernel lands).
#include "du_common.cuh"

using namespace du;

extern "C" int du_fused_supported(int64_t n, int score_dtype) {
  (void)n; (void)score_dtype;
  return 0;
}

extern "C" int du_fused_uncertainty_step(const du_fused_params* p, du_stream_t stream) {
  (void)p; (void)stream;
  return set_error(DU_ERR_TOO_LARGE, "du_fused_uncertainty_step: not available for this shape; use the unfused calls");
}
