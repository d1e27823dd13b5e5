// score_tc.cu -- stage 1 of agcf_score_topk on the 5th-gen tensor cores (tcgen05,
// TF32 inputs straight from the fp32 tables, fp32 accumulators in TMEM).
// Placeholder until the tcgen05 kernel lands: reports "unsupported" so callers
// fail loudly instead of silently falling back.
#include "common.cuh"

namespace agcf {
int launch_group_max_tc(const float*, const int32_t*, int, const float*, int, int, const uint32_t*, int, float*,
                        cudaStream_t) {
  return AGCF_EUNSUPPORTED;
}
}  // namespace agcf
