#include <math.h>
#include <stdint.h>
#define __device__
#define __forceinline__ inline
#define __host__
#include "eig3.cuh"
extern "C" void host_sym_eig3(const float* A, int64_t n, float eps, float* vals, float* vecs) {
  for (int64_t i = 0; i < n; ++i) sixdgs::sym_eig3(A + i * 9, eps, vals + i * 3, vecs + i * 9);
}
