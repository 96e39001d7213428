// Closed-form symmetric 3x3 eigen-decomposition, one matrix per thread.
// Follows the algorithm of reference pose_estimation/sym_eig_3x3.py:246-307 (trigonometric
// eigenvalues, soft diagonal blend) and :38-231 (cross-product eigenvectors with +-eps
// regularisation and argmax/argmin selections), restated as scalar device code.
#pragma once
#include <math.h>

namespace sixdgs {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float sgn_nz(float t) { return t > 0.0f ? 1.0f : -1.0f; }
__device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// eigenvector of m = A - lambda*I (rows r0,r1,r2): largest regularised row cross product
__device__ inline V3 eig_null_vector(const float* A, float lam, float eps) {
  V3 r0 = v3(A[0] - lam, A[1], A[2]);
  V3 r1 = v3(A[3], A[4] - lam, A[5]);
  V3 r2 = v3(A[6], A[7], A[8] - lam);
  V3 c01 = cross3(r0, r1), c12 = cross3(r1, r2), c02 = cross3(r0, r2);
  V3 s = v3(eps * sgn_nz(c01.x), eps * sgn_nz(c01.y), eps * sgn_nz(c01.z));
  c01 = v3(c01.x + s.x, c01.y + s.y, c01.z + s.z);
  c12 = v3(c12.x + s.x, c12.y + s.y, c12.z + s.z);
  c02 = v3(c02.x + s.x, c02.y + s.y, c02.z + s.z);
  float n0 = dot3(c01, c01), n1 = dot3(c12, c12), n2 = dot3(c02, c02);
  V3 best = c01; float nb = n0;            // argmax, first occurrence wins ties
  if (n1 > nb) { best = c12; nb = n1; }
  if (n2 > nb) { best = c02; nb = n2; }
  float inv = sqrtf(nb);
  return v3(best.x / inv, best.y / inv, best.z / inv);
}

// unit u, v with {u, v, w} right handed: quarter turn of w about the axis of its smallest |component|
__device__ inline void eig_perp_pair(V3 w, V3& u, V3& v) {
  float ax = fabsf(w.x), ay = fabsf(w.y), az = fabsf(w.z);
  int mi = 0; float mv = ax;
  if (ay < mv) { mi = 1; mv = ay; }
  if (az < mv) { mi = 2; mv = az; }
  V3 r;
  if (mi == 0) r = v3(0.0f, -w.z, w.y);
  else if (mi == 1) r = v3(-w.z, 0.0f, w.x);
  else r = v3(-w.y, w.x, 0.0f);
  float n = fmaxf(sqrtf(dot3(r, r)), 1e-12f);
  u = v3(r.x / n, r.y / n, r.z / n);
  v = cross3(w, u);
}

__device__ inline V3 mat_vec(const float* A, float lam, V3 p) {
  return v3((A[0] - lam) * p.x + A[1] * p.y + A[2] * p.z,
            A[3] * p.x + (A[4] - lam) * p.y + A[5] * p.z,
            A[6] * p.x + A[7] * p.y + (A[8] - lam) * p.z);
}

__device__ inline V3 eig_second_vector(const float* A, float lam, V3 u, V3 v, float eps) {
  V3 Mu = mat_vec(A, lam, u), Mv = mat_vec(A, lam, v);
  float m00 = dot3(u, Mu), m01 = dot3(u, Mv), m10 = dot3(v, Mu), m11 = dot3(v, Mv);
  float s = sgn_nz(m00 * m10 + m01 * m11);
  float a = m00 + s * m10, b = m01 + s * m11;
  float rs = eps * sgn_nz(a);
  a += rs; b += rs;
  float t0 = b, t1 = -a;
  float n = fmaxf(sqrtf(t0 * t0 + t1 * t1), 1e-12f);
  t0 /= n; t1 /= n;
  return v3(u.x * t0 + v.x * t1, u.y * t0 + v.y * t1, u.z * t0 + v.z * t1);
}

__device__ inline void sort3(float& a, float& b, float& c) {
  float t;
  if (a > b) { t = a; a = b; b = t; }
  if (b > c) { t = b; b = c; c = t; }
  if (a > b) { t = a; a = b; b = t; }
}

// A: row-major 3x3.  vals ascending.  vecs (nullable): row-major 3x3 whose COLUMNS are eigenvectors.
__device__ inline void sym_eig3(const float* A, float eps, float* vals, float* vecs) {
  const float d0 = A[0], d1 = A[4], d2 = A[8];
  const float q = (d0 + d1 + d2) / 3.0f;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) sq += A[i] * A[i];
  const float p1 = (sq - (d0 * d0 + d1 * d1 + d2 * d2)) / 2.0f;
  const float e0 = d0 - q, e1 = d1 - q, e2 = d2 - q;
  const float p2 = (e0 * e0 + e1 * e1 + e2 * e2) + 2.0f * fmaxf(p1, eps);
  const float p = sqrtf(p2 / 6.0f);
  float B[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) B[i] = (A[i] - ((i % 4 == 0) ? q : 0.0f)) / p;
  float det = B[0] * (B[4] * B[8] - B[5] * B[7]) - B[1] * (B[3] * B[8] - B[5] * B[6]) +
              B[2] * (B[3] * B[7] - B[4] * B[6]);
  float r = det / 2.0f;
  r = fminf(fmaxf(r, -1.0f + eps), 1.0f - eps);
  const float phi = acosf(r) / 3.0f;
  const float big = q + 2.0f * p * cosf(phi);
  const float small = q + 2.0f * p * cosf(phi + 2.0943951023931953f);
  const float mid = 3.0f * q - big - small;
  const float t = p1 / (6.0f * eps);
  const float soft = expf(-(t * t));
  float s0 = d0, s1 = d1, s2 = d2;
  sort3(s0, s1, s2);
  const float l0 = soft * s0 + (1.0f - soft) * small;
  const float l1 = soft * s1 + (1.0f - soft) * mid;
  const float l2 = soft * s2 + (1.0f - soft) * big;
  vals[0] = l0; vals[1] = l1; vals[2] = l2;
  if (vecs == nullptr) return;
  V3 c0, c1, c2;
  if ((l1 - l0) > (l2 - l1)) {
    V3 a0 = eig_null_vector(A, l0, eps), u, v;
    eig_perp_pair(a0, u, v);
    V3 a1 = eig_second_vector(A, l1, u, v, eps);
    c0 = a0; c1 = a1; c2 = cross3(a0, a1);
  } else {
    V3 a0 = eig_null_vector(A, l2, eps), u, v;
    eig_perp_pair(a0, u, v);
    V3 a1 = eig_second_vector(A, l1, u, v, eps);
    c2 = a0; c1 = a1; c0 = cross3(a0, a1);
  }
  vecs[0] = c0.x; vecs[1] = c1.x; vecs[2] = c2.x;
  vecs[3] = c0.y; vecs[4] = c1.y; vecs[5] = c2.y;
  vecs[6] = c0.z; vecs[7] = c1.z; vecs[8] = c2.z;
}

}  // namespace sixdgs
