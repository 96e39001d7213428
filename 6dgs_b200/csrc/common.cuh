// Shared helpers for libsixdgs (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sixdgs.h"

namespace sixdgs {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SIXDGS_ECUDA;
  }
  return SIXDGS_OK;
}

#define SIXDGS_REQUIRE(cond, msg)            \
  do {                                       \
    if (!(cond)) {                           \
      ::sixdgs::set_error("%s: %s", __func__, msg); \
      return SIXDGS_EINVAL;                  \
    }                                        \
  } while (0)

constexpr int kFeat = SIXDGS_FEAT;
constexpr int kMaxTokens = SIXDGS_MAX_TOKENS;
constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace sixdgs
