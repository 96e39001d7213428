// a4 / a5: exact brute-force k-NN normals and the batched 3x3 eigen-solver entry point.
// Reference: pose_estimation/sampling.py:37-113 (compute_normals, disambiguate_vector_directions),
//            pose_estimation/sym_eig_3x3.py:246-307.
//
// The reference finds neighbours with torch.cdist + topk over a dense [chunk, M] distance matrix
// (O(M^2) memory-bound, which is one reason it caps the scene at 1000 ellipsoids).  Here one thread
// owns one query point, the cloud streams through shared memory in tiles that every thread of the
// CTA scans, and the running k-best list lives in registers/local memory: O(M^2) FLOPs but no
// distance matrix, 12*M bytes of HBM traffic per CTA sweep (L2 resident for M <= 10M).
#include "common.cuh"
#include "eig3.cuh"
#include "normal_fit.cuh"

namespace sixdgs {

constexpr int kKnnTile = 1024;
constexpr int kKnnThreads = 128;
constexpr int kKnnMaxK = 32;

__global__ void __launch_bounds__(kKnnThreads)
knn_normals_kernel(const float* __restrict__ cloud, int64_t m, int64_t q_begin, int64_t q_count,
                   int k, float* __restrict__ out) {
  __shared__ float sx[kKnnTile], sy[kKnnTile], sz[kKnnTile];
  const int64_t qi = (int64_t)blockIdx.x * kKnnThreads + threadIdx.x;
  const bool live = qi < q_count;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (live) {
    const float* p = cloud + (q_begin + qi) * 3;
    px = p[0]; py = p[1]; pz = p[2];
  }
  float bd[kKnnMaxK];
  int bi[kKnnMaxK];
#pragma unroll
  for (int i = 0; i < kKnnMaxK; ++i) { bd[i] = INFINITY; bi[i] = -1; }
  float worst = INFINITY;

  for (int64_t t0 = 0; t0 < m; t0 += kKnnTile) {
    const int cnt = (int)min((int64_t)kKnnTile, m - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += kKnnThreads) {
      const float* p = cloud + (t0 + i) * 3;
      sx[i] = p[0]; sy[i] = p[1]; sz[i] = p[2];
    }
    __syncthreads();
    if (!live) continue;
    for (int i = 0; i < cnt; ++i) {
      const float dx = px - sx[i], dy = py - sy[i], dz = pz - sz[i];
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < worst) {
        // insertion into the ascending list (ties keep the earlier index first)
        int pos = k - 1;
        while (pos > 0 && bd[pos - 1] > d) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bd[pos] = d; bi[pos] = (int)(t0 + i);
        worst = bd[k - 1];
      }
    }
  }
  if (!live) return;

  normal_from_neighbours(cloud, bi, k, out + qi * 3);
}

__global__ void sym_eig_kernel(const float* __restrict__ A, int64_t n, float eps,
                               float* __restrict__ vals, float* __restrict__ vecs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[9], l[3], v[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) a[j] = A[i * 9 + j];
  sym_eig3(a, eps, l, vecs ? v : nullptr);
  vals[i * 3 + 0] = l[0]; vals[i * 3 + 1] = l[1]; vals[i * 3 + 2] = l[2];
  if (vecs) {
#pragma unroll
    for (int j = 0; j < 9; ++j) vecs[i * 9 + j] = v[j];
  }
}

}  // namespace sixdgs

using namespace sixdgs;

extern "C" int sixdgs_knn_normals(const float* cloud, int64_t m, int64_t q_begin, int64_t q_count,
                                  int k, float* normals_out, void* stream) {
  SIXDGS_REQUIRE(cloud && normals_out, "null pointer");
  SIXDGS_REQUIRE(k >= 1 && k <= kKnnMaxK, "k must be in [1, 32]");
  SIXDGS_REQUIRE(m >= k, "cloud has fewer than k points");
  SIXDGS_REQUIRE(q_begin >= 0 && q_count >= 0 && q_begin + q_count <= m, "query range out of bounds");
  if (q_count == 0) return SIXDGS_OK;
  const unsigned blocks = (unsigned)((q_count + kKnnThreads - 1) / kKnnThreads);
  knn_normals_kernel<<<blocks, kKnnThreads, 0, (cudaStream_t)stream>>>(cloud, m, q_begin, q_count, k,
                                                                       normals_out);
  return check_launch("knn_normals");
}

extern "C" int sixdgs_sym_eig3x3(const float* A, int64_t n, float eps, float* vals, float* vecs,
                                 void* stream) {
  SIXDGS_REQUIRE(A && vals, "null pointer");
  SIXDGS_REQUIRE(n >= 0, "negative count");
  if (n == 0) return SIXDGS_OK;
  if (!(eps > 0.f)) eps = 1.1920928955078125e-07f;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  sym_eig_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(A, n, eps, vals, vecs);
  return check_launch("sym_eig3x3");
}
