// fp32 SIMT "NT" GEMM main loop shared by the ray-feature MLP (features.cu) and the exact-precision
// ray-score kernels (score_simt.cu):  acc[i][j] = sum_k A[row_i][k] * B[col_j][k]
// with A = activations / key-cache rows (fp32 or bf16) and B = weight / query rows (fp32), both
// K-major.  256 threads, 8x8 register micro-tile per thread (as 2x2 blocks of 4x4 so that shared
// memory reads are conflict-free LDS.128), BK = 16, double-buffered shared tiles.
#pragma once
#include "common.cuh"

namespace sixdgs {

constexpr int kBK = 16;

template <int BM, int BN>
struct GemmSmem {
  static constexpr int LDA = BM + 4;
  static constexpr int LDB = BN + 4;
  float a[2][kBK][LDA];
  float b[2][kBK][LDB];
};

template <typename T>
struct RowLoader;  // loads 4 consecutive k values of one row

template <>
struct RowLoader<float> {
  __device__ static __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
};
template <>
struct RowLoader<__nv_bfloat16> {
  __device__ static __forceinline__ float4 load(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

// Thread layout: tx = tid % (BN/8) owns columns {4tx..4tx+3} and {BN/2+4tx..}; ty = tid / (BN/8)
// owns rows {4ty..4ty+3} and {BM/2+4ty..}.
template <int BM, int BN, typename TA>
__device__ __forceinline__ void gemm_nt_mainloop(const TA* __restrict__ A, int64_t lda, int64_t a_rows_valid,
                                                 const float* __restrict__ B, int64_t ldb, int b_rows_valid,
                                                 int K, GemmSmem<BM, BN>& sm, float (&acc)[8][8]) {
  static_assert((BM / 8) * (BN / 8) == 256, "256 threads expected");
  constexpr int A_F4 = BM * kBK / 4;  // float4 loads for the A tile
  constexpr int B_F4 = BN * kBK / 4;
  constexpr int A_PER = (A_F4 + 255) / 256;
  constexpr int B_PER = (B_F4 + 255) / 256;
  const int tid = threadIdx.x;
  const int tx = tid % (BN / 8), ty = tid / (BN / 8);

  float4 ra[A_PER], rb[B_PER];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int f = tid + i * 256;
      const int row = f >> 2, c = f & 3;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < A_F4 && row < a_rows_valid) ra[i] = RowLoader<TA>::load(A + (int64_t)row * lda + k0 + c * 4);
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int f = tid + i * 256;
      const int row = f >> 2, c = f & 3;
      rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < B_F4 && row < b_rows_valid) rb[i] = RowLoader<float>::load(B + (int64_t)row * ldb + k0 + c * 4);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int f = tid + i * 256;
      if (f < A_F4) {
        const int row = f >> 2, c = f & 3;
        sm.a[buf][c * 4 + 0][row] = ra[i].x; sm.a[buf][c * 4 + 1][row] = ra[i].y;
        sm.a[buf][c * 4 + 2][row] = ra[i].z; sm.a[buf][c * 4 + 3][row] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int f = tid + i * 256;
      if (f < B_F4) {
        const int row = f >> 2, c = f & 3;
        sm.b[buf][c * 4 + 0][row] = rb[i].x; sm.b[buf][c * 4 + 1][row] = rb[i].y;
        sm.b[buf][c * 4 + 2][row] = rb[i].z; sm.b[buf][c * 4 + 3][row] = rb[i].w;
      }
    }
  };

#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / kBK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * kBK);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[buf][k][BM / 2 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[buf][k][BN / 2 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
    }
    __syncthreads();
  }
}

// micro-tile index -> tile-local row / column
template <int BM>
__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4); }
template <int BN>
__device__ __forceinline__ int tile_col(int tx, int j) { return (j < 4) ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4); }

}  // namespace sixdgs
