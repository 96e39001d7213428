// Library-level entry points: version, thread-local error string, device check, score dispatch.
#include <stdarg.h>
#include "common.cuh"

namespace sixdgs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int score_simt_pass1(const void*, int, int64_t, const float*, int, float*, float*, cudaStream_t);
int score_simt_pass2(const void*, int, int64_t, const float*, int, const float*, const float*, float*, float*,
                     cudaStream_t);
int score_simt_parts();
int score_tc_pass1(const void*, int64_t, const float*, int, float*, float*, void*, size_t, cudaStream_t);
int score_tc_pass2(const void*, int64_t, const float*, int, const float*, const float*, float*, void*, size_t,
                   cudaStream_t);
int score_tc_parts();
size_t score_tc_workspace();
// exact tensor-core mode (f16x2 key cache), score_tc_mq.cu
int score_tc_split_pass1(const void*, int, int64_t, const float*, int, float*, float*, void*, size_t, cudaStream_t);
int score_tc_split_pass2(const void*, int, int64_t, const float*, int, const float*, const float*, float*, void*, size_t,
                         cudaStream_t);
size_t score_tc_split_workspace();

}  // namespace sixdgs

using namespace sixdgs;

extern "C" int sixdgs_version(void) { return 200; }  // 0.2.0
extern "C" const char* sixdgs_last_error(void) { return g_err; }

extern "C" int sixdgs_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int sixdgs_score_parts(int impl) { return impl == 1 ? score_tc_parts() : score_simt_parts(); }
extern "C" size_t sixdgs_score_workspace(int impl) {
  if (impl != 1) return 0;
  const size_t a = score_tc_workspace(), b = score_tc_split_workspace();
  return a > b ? a : b;
}

static int check_score_args(const void* k, int k_dtype, int64_t n_rays, const float* q, int n_img, int impl) {
  SIXDGS_REQUIRE(k && q, "null pointer");
  SIXDGS_REQUIRE(k_dtype == SIXDGS_F32 || k_dtype == SIXDGS_BF16 || k_dtype == SIXDGS_F16X2 || k_dtype == SIXDGS_F16F8,
                 "unsupported k_dtype");
  SIXDGS_REQUIRE(n_rays > 0, "n_rays must be positive");
  SIXDGS_REQUIRE(n_img > 0 && n_img <= kMaxTokens, "n_img must be in [1, 256]");
  SIXDGS_REQUIRE(impl == 0 || impl == 1, "impl must be 0 (SIMT fp32) or 1 (tcgen05 tensor cores)");
  SIXDGS_REQUIRE(impl == 0 || k_dtype != SIXDGS_F32, "impl 1 needs a bf16 or f16x2 key cache");
  SIXDGS_REQUIRE(impl == 1 || (k_dtype != SIXDGS_F16X2 && k_dtype != SIXDGS_F16F8), "an f16x2 / f16f8 key cache needs impl 1");
  return SIXDGS_OK;
}

extern "C" int sixdgs_score_pass1(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_img,
                                  float* part_m, float* part_z, int impl, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  int rc = check_score_args(k_cache, k_dtype, n_rays, q, n_img, impl);
  if (rc) return rc;
  SIXDGS_REQUIRE(part_m && part_z, "null pointer");
  if (impl == 1 && (k_dtype == SIXDGS_F16X2 || k_dtype == SIXDGS_F16F8))
    return score_tc_split_pass1(k_cache, k_dtype, n_rays, q, n_img, part_m, part_z, workspace, workspace_bytes,
                                (cudaStream_t)stream);
  if (impl == 1)
    return score_tc_pass1(k_cache, n_rays, q, n_img, part_m, part_z, workspace, workspace_bytes, (cudaStream_t)stream);
  return score_simt_pass1(k_cache, k_dtype, n_rays, q, n_img, part_m, part_z, (cudaStream_t)stream);
}

extern "C" int sixdgs_score_pass2(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_img,
                                  const float* m, const float* z, float* scores, float* attn_map, int impl,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_score_args(k_cache, k_dtype, n_rays, q, n_img, impl);
  if (rc) return rc;
  SIXDGS_REQUIRE(m && z && scores, "null pointer");
  if (impl == 1) {
    SIXDGS_REQUIRE(attn_map == nullptr, "impl 1 does not materialise the attention map");
    if (k_dtype == SIXDGS_F16X2 || k_dtype == SIXDGS_F16F8)
      return score_tc_split_pass2(k_cache, k_dtype, n_rays, q, n_img, m, z, scores, workspace, workspace_bytes,
                                  (cudaStream_t)stream);
    return score_tc_pass2(k_cache, n_rays, q, n_img, m, z, scores, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  return score_simt_pass2(k_cache, k_dtype, n_rays, q, n_img, m, z, scores, attn_map, (cudaStream_t)stream);
}
