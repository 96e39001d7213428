// Normal of one point from its k nearest neighbours (shared by the brute-force and the grid k-NN kernels).
#pragma once
#include "eig3.cuh"

namespace sixdgs {

// normal of one point from its k nearest neighbours (indices bi[0..k), ascending distance):
// centred neighbourhood, scatter matrix X^T X (sampling.py:90-93), smallest-eigenvalue eigenvector,
// majority-sign disambiguation (sampling.py:37-59), normalisation.
__device__ __forceinline__ void normal_from_neighbours(const float* __restrict__ cloud, const int* bi, int k,
                                                       float* __restrict__ out3) {
  float mx = 0.f, my = 0.f, mz = 0.f;
  for (int i = 0; i < k; ++i) {
    const float* p = cloud + (int64_t)bi[i] * 3;
    mx += p[0]; my += p[1]; mz += p[2];
  }
  mx /= (float)k; my /= (float)k; mz /= (float)k;
  float A[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < k; ++i) {
    const float* p = cloud + (int64_t)bi[i] * 3;
    const float cx = p[0] - mx, cy = p[1] - my, cz = p[2] - mz;
    A[0] += cx * cx; A[1] += cx * cy; A[2] += cx * cz;
    A[4] += cy * cy; A[5] += cy * cz; A[8] += cz * cz;
  }
  A[3] = A[1]; A[6] = A[2]; A[7] = A[5];
  float vals[3], vecs[9];
  sym_eig3(A, 1.1920928955078125e-07f, vals, vecs);
  float nx = vecs[0], ny = vecs[3], nz = vecs[6];
  int npos = 0;
  for (int i = 0; i < k; ++i) {
    const float* p = cloud + (int64_t)bi[i] * 3;
    const float pr = nx * (p[0] - mx) + ny * (p[1] - my) + nz * (p[2] - mz);
    npos += (pr > 0.f) ? 1 : 0;
  }
  if ((float)npos < 0.5f * (float)k) { nx = -nx; ny = -ny; nz = -nz; }
  const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
  out3[0] = nx / nn; out3[1] = ny / nn; out3[2] = nz / nn;
}

}  // namespace sixdgs
