// a10 (+ k_proj of a11): ray (ori, dir, rgb) -> positional encoding -> 4-layer MLP -> key projection.
// Reference: pose_estimation/ray_preprocessor.py:3-46, our_multihead_attention.py:75.
//
// The reference recomputes this for every query although it does not depend on the image
// (identification_module.py:79-80); here it runs once per scene (or weight update) and its output
// -- the key cache K[n_rays, 384] in fp32 or bf16 -- is what the per-query score kernels stream.
//
// Layout of the per-chunk workspace (fp32, row-major, chunk = kChunk rays):
//   X  [chunk, 672] : cols 0..511 = second hidden layer h, cols 512..652 = the 141-wide MLP input
//                     (9 raw + PE), cols 653..671 = 0  -> the concat [h, x] of mlp2 is free
//                     (672 = 512 + 160: every reduction dim is a multiple of 32 for the TF32 UMMA k-blocks)
//   H  [chunk, 512] : first hidden layer, later the third hidden layer
//   F  [chunk, 384] : ray feature (output of mlp2) when a projection follows
#include "common.cuh"
#include "gemm_simt.cuh"

namespace sixdgs {

constexpr int kChunk = 1 << 17;  // rays per workspace chunk
constexpr int kXW = 672;         // padded concat width (512 + 141 -> 672)
constexpr int kInPad = 160;      // padded MLP input width (141 -> 160)
constexpr int kIn = 141;

// x = [ori3, dir3, rgb3, sin(ori*2^f) (coordinate-major, f=0..7), cos(..), sin(dir*2^f), cos(..),
//      sin(rgb*2^f) f=0..5, cos(..)]                                  ray_preprocessor.py:3-9,36-44
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
  return __uint_as_float(t);
}

__global__ void pe_kernel(const float* __restrict__ ori, const float* __restrict__ dir,
                          const float* __restrict__ rgb, int64_t n, float* __restrict__ X, int round_tf32) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float* x = X + r * kXW + 512;
  float v[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) { v[i] = ori[r * 3 + i]; v[3 + i] = dir[r * 3 + i]; v[6 + i] = rgb[r * 3 + i]; }
#pragma unroll
  for (int i = 0; i < 9; ++i) x[i] = v[i];
  int o = 9;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int nf = (g == 2) ? 6 : 8;
    for (int c = 0; c < 3; ++c)
      for (int f = 0; f < nf; ++f) {
        const float a = v[g * 3 + c] * (float)(1 << f);
        x[o + c * nf + f] = sinf(a);
        x[o + 3 * nf + c * nf + f] = cosf(a);
      }
    o += 6 * nf;
  }
#pragma unroll
  for (int i = kIn; i < kInPad; ++i) x[i] = 0.f;
  if (round_tf32) {  // tensor-core path: make the TF32 truncation of the MMA exact (see features_tc.cu)
    for (int i = 0; i < kIn; ++i) x[i] = rna_tf32(x[i]);
  }
}

template <typename TO>
__device__ __forceinline__ void store_out(TO* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// y[m, n] = act(x[m, k] w[n, k]^T + b[n])
template <typename TO>
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ x, int64_t m, int k, int64_t lda, const float* __restrict__ w,
              const float* __restrict__ b, int n, TO* __restrict__ y, int64_t ldc, int relu) {
  constexpr int BM = 128, BN = 128;
  __shared__ GemmSmem<BM, BN> sm;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  float acc[8][8];
  gemm_nt_mainloop<BM, BN, float>(x + row0 * lda, lda, min((int64_t)BM, m - row0), w + (int64_t)col0 * k, k,
                                  min(BN, n - col0), k, sm, acc);
  const int tx = threadIdx.x % (BN / 8), ty = threadIdx.x / (BN / 8);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + tile_row<BM>(ty, i);
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = col0 + tile_col<BN>(tx, j);
      if (col >= n) continue;
      float v = acc[i][j] + (b ? b[col] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      store_out<TO>(y + row * ldc + col, v);
    }
  }
}

template <typename TO>
static int launch_linear(const float* x, int64_t m, int k, int64_t lda, const float* w, const float* b, int n,
                         TO* y, int64_t ldc, int relu, cudaStream_t s) {
  dim3 grid((unsigned)((m + 127) / 128), (unsigned)((n + 127) / 128));
  linear_kernel<TO><<<grid, 256, 0, s>>>(x, m, k, lda, w, b, n, y, ldc, relu);
  return check_launch("linear");
}

// tensor-core (TF32 tcgen05) GEMM, features_tc.cu
template <typename TO>
int launch_linear_tc(const float* x, int64_t m, int k, int64_t lda, const float* w, const float* b, int n, TO* y,
                     int64_t ldc, int relu, int round_tf32, cudaStream_t s);

// the same GEMM with a shared-memory staged TMA-store epilogue (features_tc.cu): bit-identical, 8 % faster -> the default
template <typename TO>
int launch_linear_tc_staged(const float* x, int64_t m, int k, int64_t lda, const float* w, const float* b, int n, TO* y,
                            int64_t ldc, int relu, int round_tf32, cudaStream_t s);

// `feeds_gemm`: the output is the A operand of another tensor-core layer -> round it to TF32 on store
template <typename TO>
static int launch_any(int impl, const float* x, int64_t m, int k, int64_t lda, const float* w, const float* b, int n, TO* y,
                      int64_t ldc, int relu, cudaStream_t s, int feeds_gemm = 1) {
  if (impl == 1) return launch_linear_tc_staged<TO>(x, m, k, lda, w, b, n, y, ldc, relu, feeds_gemm, s);
  if (impl == 3) return launch_linear_tc<TO>(x, m, k, lda, w, b, n, y, ldc, relu, feeds_gemm, s);
  return launch_linear<TO>(x, m, k, lda, w, b, n, y, ldc, relu, s);
}

}  // namespace sixdgs

using namespace sixdgs;

extern "C" size_t sixdgs_ray_features_workspace(int64_t n) {
  const int64_t c = n < kChunk ? (n > 0 ? n : 1) : kChunk;
  return (size_t)c * (kXW + 512 + kFeat) * sizeof(float);
}

extern "C" int sixdgs_linear(const float* x, int64_t m, int k, int lda, const float* w, const float* b, int n,
                             float* y, int ldc, int relu, void* stream) {
  SIXDGS_REQUIRE(x && w && y, "null pointer");
  SIXDGS_REQUIRE(k > 0 && k % 16 == 0 && lda % 4 == 0 && lda >= k, "k must be a multiple of 16, lda of 4");
  SIXDGS_REQUIRE(m >= 0 && n > 0 && ldc >= n, "bad size");
  if (m == 0) return SIXDGS_OK;
  return launch_linear<float>(x, m, k, lda, w, b, n, y, ldc, relu, (cudaStream_t)stream);
}

extern "C" int sixdgs_ray_features(const float* ori, const float* dir, const float* rgb, int64_t n,
                                   const float* w1p, const float* b1, const float* w2, const float* b2,
                                   const float* w3p, const float* b3, const float* w4, const float* b4,
                                   const float* wk, const float* bk, void* k_out, int k_dtype, float* feat_out,
                                   int impl, void* workspace, size_t workspace_bytes, void* stream) {
  SIXDGS_REQUIRE(ori && dir && rgb && w1p && b1 && w2 && b2 && w3p && b3 && w4 && b4, "null pointer");
  SIXDGS_REQUIRE(k_out || feat_out, "no output requested");
  SIXDGS_REQUIRE(!k_out || k_dtype == SIXDGS_F32 || k_dtype == SIXDGS_BF16, "unsupported k_dtype");
  SIXDGS_REQUIRE(!k_out || !wk || bk, "wk without bk");
  SIXDGS_REQUIRE(n >= 0, "negative size");
  SIXDGS_REQUIRE(impl == 0 || impl == 1 || impl == 3, "impl must be 0 (fp32 SIMT), 1 (TF32 tcgen05) or 3 (TF32, direct-store epilogue)");
  if (n == 0) return SIXDGS_OK;
  if (workspace == nullptr || workspace_bytes < sixdgs_ray_features_workspace(n)) {
    set_error("ray_features: workspace too small (%zu < %zu)", workspace_bytes, sixdgs_ray_features_workspace(n));
    return SIXDGS_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t cap = n < kChunk ? n : kChunk;
  float* X = (float*)workspace;
  float* H = X + cap * kXW;
  float* F = H + cap * 512;
  for (int64_t r0 = 0; r0 < n; r0 += cap) {
    const int64_t c = (n - r0) < cap ? (n - r0) : cap;
    pe_kernel<<<(unsigned)((c + 127) / 128), 128, 0, s>>>(ori + r0 * 3, dir + r0 * 3, rgb + r0 * 3, c, X, impl >= 1);
    int rc = check_launch("pe");
    if (rc) return rc;
    // mlp.0: x(160) -> H(512), relu ; mlp.2: H -> X[:, :512], relu
    if ((rc = launch_any<float>(impl, X + 512, c, kInPad, kXW, w1p, b1, 512, H, 512, 1, s))) return rc;
    if ((rc = launch_any<float>(impl, H, c, 512, 512, w2, b2, 512, X, kXW, 1, s))) return rc;
    // mlp2.0: [h, x](672) -> H(512), relu
    if ((rc = launch_any<float>(impl, X, c, kXW, kXW, w3p, b3, 512, H, 512, 1, s))) return rc;
    const bool project = (k_out != nullptr) && (wk != nullptr);
    if (k_out && !project) {
      // no projection requested: the MLP feature itself is the key-shaped output (feat_out must then be NULL)
      if (feat_out) { set_error("ray_features: k_out without wk requires feat_out == NULL"); return SIXDGS_EINVAL; }
      if (k_dtype == SIXDGS_F32)
        rc = launch_any<float>(impl, H, c, 512, 512, w4, b4, kFeat, (float*)k_out + r0 * kFeat, kFeat, 0, s, 0);
      else
        rc = launch_any<__nv_bfloat16>(impl, H, c, 512, 512, w4, b4, kFeat, (__nv_bfloat16*)k_out + r0 * kFeat, kFeat, 0, s, 0);
      if (rc) return rc;
      continue;
    }
    // mlp2.2: H -> feature(384).  It feeds k_proj when a projection follows; if the caller also wants the
    // feature back it stays unrounded (only the last GEMM then sees TF32 truncation on its A operand).
    float* fdst = feat_out ? feat_out + r0 * kFeat : F;
    if ((rc = launch_any<float>(impl, H, c, 512, 512, w4, b4, kFeat, fdst, kFeat, 0, s, (project && !feat_out) ? 1 : 0))) return rc;
    if (project) {
      if (k_dtype == SIXDGS_F32)
        rc = launch_any<float>(impl, fdst, c, kFeat, kFeat, wk, bk, kFeat, (float*)k_out + r0 * kFeat, kFeat, 0, s, 0);
      else
        rc = launch_any<__nv_bfloat16>(impl, fdst, c, kFeat, kFeat, wk, bk, kFeat, (__nv_bfloat16*)k_out + r0 * kFeat, kFeat, 0, s, 0);
      if (rc) return rc;
    }
  }
  return SIXDGS_OK;
}
