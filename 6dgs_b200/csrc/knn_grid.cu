// a4 at scale: exact k-NN normals with a uniform grid (reference semantics: pose_estimation/sampling.py:62-113,
// exact cdist + topk).  Same result as the brute-force kernel in normals.cu -- the search in knn_grid.cuh is
// provably exhaustive -- at O(M) instead of O(M^2) work: 1M points in tens of milliseconds instead of ~1 s.
// Pipeline (all stream ordered, no host sync): cell ids + histogram -> exclusive scan -> scatter -> one thread per
// query in cell order (neighbouring threads walk the same cells, so the point reads hit L1/L2).  The order of the
// points inside a cell depends on atomics, the result does not: candidates are ranked by (distance, index), which
// is independent of the order in which they are visited.  Points outside the grid box (the host may build it from
// robust statistics (quantiles) so that a few far outliers do not inflate the cells) are clamped into the border cells; the
// search treats faces on the grid boundary as open, so it stays exhaustive.
#include "common.cuh"
#include "normal_fit.cuh"
#include "knn_grid.cuh"

namespace sixdgs {

__global__ void grid_count_kernel(const float* __restrict__ cloud, int64_t m, KnnGrid g, int* __restrict__ cell_of,
                                  int* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int x = knn_cell_coord(cloud[i * 3], g.lo[0], g.h, g.dim[0]);
  const int y = knn_cell_coord(cloud[i * 3 + 1], g.lo[1], g.h, g.dim[1]);
  const int z = knn_cell_coord(cloud[i * 3 + 2], g.lo[2], g.h, g.dim[2]);
  const int c = (z * g.dim[1] + y) * g.dim[0] + x;
  cell_of[i] = c;
  atomicAdd(&count[c], 1);
}

__global__ void grid_scatter_kernel(const int* __restrict__ cell_of, int64_t m, const int64_t* __restrict__ cell_start,
                                    int* __restrict__ cursor, int* __restrict__ sorted_idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int c = cell_of[i];
  sorted_idx[cell_start[c] + atomicAdd(&cursor[c], 1)] = (int)i;
}

constexpr int kGridMaxK = 32;

__global__ void __launch_bounds__(128)
grid_query_kernel(const float* __restrict__ cloud, int64_t m, KnnGrid g, const int64_t* __restrict__ cell_start,
                  const int* __restrict__ sorted_idx, int64_t q_begin, int64_t q_count, int k, float* __restrict__ out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // position in cell order
  if (s >= m) return;
  const int64_t q = sorted_idx[s];
  if (q < q_begin || q >= q_begin + q_count) return;
  float bd[kGridMaxK];
  int bi[kGridMaxK];
  knn_grid_query(cloud, g, cell_start, sorted_idx, cloud[q * 3], cloud[q * 3 + 1], cloud[q * 3 + 2], k, bd, bi, -1);
  normal_from_neighbours(cloud, bi, k, out + (q - q_begin) * 3);
}

}  // namespace sixdgs

using namespace sixdgs;

extern "C" size_t sixdgs_knn_grid_workspace(int64_t m, int64_t n_cells) {
  return (size_t)m * 8 + (size_t)n_cells * 8 + (size_t)(n_cells + 1) * 8 + 256;
}

extern "C" int sixdgs_knn_normals_grid(const float* cloud, int64_t m, int64_t q_begin, int64_t q_count, int k,
                                       const float* grid_lo_host, float cell, const int* dims_host, float* normals_out,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  SIXDGS_REQUIRE(cloud && normals_out && grid_lo_host && dims_host && workspace, "null pointer");
  SIXDGS_REQUIRE(k >= 1 && k <= kGridMaxK, "k must be in [1, 32]");
  SIXDGS_REQUIRE(m >= k && m < (int64_t)INT32_MAX, "cloud must have k <= m < 2^31 points");
  SIXDGS_REQUIRE(q_begin >= 0 && q_count >= 0 && q_begin + q_count <= m, "query range out of bounds");
  SIXDGS_REQUIRE(cell > 0.f && dims_host[0] > 0 && dims_host[1] > 0 && dims_host[2] > 0, "bad grid");
  const int64_t n_cells = (int64_t)dims_host[0] * dims_host[1] * dims_host[2];
  SIXDGS_REQUIRE(n_cells < (int64_t)INT32_MAX, "too many cells");
  if (workspace_bytes < sixdgs_knn_grid_workspace(m, n_cells)) { set_error("knn_grid: workspace too small"); return SIXDGS_EWORKSPACE; }
  if (q_count == 0) return SIXDGS_OK;
  KnnGrid g;
  for (int a = 0; a < 3; ++a) { g.lo[a] = grid_lo_host[a]; g.dim[a] = dims_host[a]; }
  g.h = cell;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* w = (unsigned char*)workspace;
  int64_t* cell_start = (int64_t*)w; w += (size_t)(n_cells + 1) * 8;
  int* cell_of = (int*)w; w += (size_t)m * 4;
  int* sorted_idx = (int*)w; w += (size_t)m * 4;
  int* count = (int*)w; w += (size_t)n_cells * 4;
  int* cursor = (int*)w;
  cudaError_t e = cudaMemsetAsync(count, 0, (size_t)n_cells * 8, s);  // count + cursor are adjacent
  if (e != cudaSuccess) { set_error("knn_grid memset: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  const unsigned pb = (unsigned)((m + 255) / 256);
  grid_count_kernel<<<pb, 256, 0, s>>>(cloud, m, g, cell_of, count);
  int rc = sixdgs_exclusive_scan(count, n_cells, cell_start, stream);
  if (rc) return rc;
  grid_scatter_kernel<<<pb, 256, 0, s>>>(cell_of, m, cell_start, cursor, sorted_idx);
  grid_query_kernel<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(cloud, m, g, cell_start, sorted_idx, q_begin, q_count, k,
                                                             normals_out);
  return check_launch("knn_normals_grid");
}
