// placeholder for the tcgen05 ray-score kernel (filled in by the next milestone)
#include "common.cuh"
namespace sixdgs {
int score_tc_parts() { return kNumSMs; }
int score_tc_pass1(const void*, int64_t, const float*, int, float*, float*, cudaStream_t) {
  set_error("score_tc: not built yet");
  return SIXDGS_EUNSUPPORTED;
}
int score_tc_pass2(const void*, int64_t, const float*, int, const float*, const float*, float*, cudaStream_t) {
  set_error("score_tc: not built yet");
  return SIXDGS_EUNSUPPORTED;
}
}  // namespace sixdgs
