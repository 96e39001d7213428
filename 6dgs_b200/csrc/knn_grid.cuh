// Exact k-nearest-neighbour search on a uniform grid (host/device shared core, so the search logic can be
// unit-tested on the CPU against brute force -- tests/test_knn_grid_host.py compiles this header with g++).
//
// Points are bucketed into cubic cells of edge h over a box around the cloud (counting sort: cell_start /
// sorted_idx; points outside the box are clamped into the border cells, whose outward faces are treated as open).  A query visits the cells in growing Chebyshev shells around
// its own cell and stops as soon as its current k-th best distance is no larger than the distance to the
// surface of the cube of cells already visited -- every unvisited point is at least that far away, so the
// result is EXACTLY the brute-force answer.  Candidates are ordered by (distance, index), which is the order
// the brute-force kernel produces (ascending index scan with a strict < insertion), ties included.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef __CUDACC__
#define SIXDGS_HD
#else
#define SIXDGS_HD __host__ __device__ __forceinline__
#endif

namespace sixdgs {

struct KnnGrid {
  float lo[3];   // bounding-box minimum
  float h;       // cell edge
  int dim[3];    // cells per axis
};

SIXDGS_HD int knn_cell_coord(float p, float lo, float h, int dim) {
  int c = (int)floorf((p - lo) / h);
  return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

// insert (d, idx) into the ascending (d, idx)-ordered list of length k
SIXDGS_HD void knn_insert(float* bd, int* bi, int k, float d, int idx) {
  int pos = k - 1;
  while (pos > 0 && (bd[pos - 1] > d || (bd[pos - 1] == d && bi[pos - 1] > idx))) {
    bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos;
  }
  bd[pos] = d; bi[pos] = idx;
}

// Returns the number of shells visited (for statistics); bd/bi hold the k nearest (ascending) on return.
// max_shell < 0: unlimited.  If the search is cut by max_shell the function returns -1 (caller falls back).
SIXDGS_HD int knn_grid_query(const float* __restrict__ cloud, const KnnGrid& g, const int64_t* __restrict__ cell_start,
                             const int* __restrict__ sorted_idx, float px, float py, float pz, int k, float* bd, int* bi,
                             int max_shell) {
  for (int i = 0; i < k; ++i) { bd[i] = INFINITY; bi[i] = 0x7fffffff; }
  const int cx = knn_cell_coord(px, g.lo[0], g.h, g.dim[0]);
  const int cy = knn_cell_coord(py, g.lo[1], g.h, g.dim[1]);
  const int cz = knn_cell_coord(pz, g.lo[2], g.h, g.dim[2]);
  const int rmax_grid = max(max(max(cx, g.dim[0] - 1 - cx), max(cy, g.dim[1] - 1 - cy)), max(cz, g.dim[2] - 1 - cz));
  for (int r = 0;; ++r) {
    if (max_shell >= 0 && r > max_shell) return -1;
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.dim[2] - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.dim[1] - 1);
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.dim[0] - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const bool face_zy = (z == cz - r) || (z == cz + r) || (y == cy - r) || (y == cy + r);
        // on a z/y face of the shell every x belongs to it; otherwise only the two x faces do
        const int xstep = face_zy ? 1 : max(x1 - x0, 1);
        for (int x = x0; x <= x1; x += xstep) {
          if (!face_zy && x != cx - r && x != cx + r) continue;
          const int64_t cell = ((int64_t)z * g.dim[1] + y) * g.dim[0] + x;
          for (int64_t s = cell_start[cell]; s < cell_start[cell + 1]; ++s) {
            const int j = sorted_idx[s];
            const float dx = px - cloud[(int64_t)j * 3], dy = py - cloud[(int64_t)j * 3 + 1], dz = pz - cloud[(int64_t)j * 3 + 2];
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < bd[k - 1] || (d == bd[k - 1] && j < bi[k - 1])) knn_insert(bd, bi, k, d, j);
          }
        }
      }
    if (r >= rmax_grid) return r;  // the whole grid has been visited
    // distance from the query to the surface of the visited cube [c-r, c+r] (faces on the grid boundary do not count)
    float bound = INFINITY;
    if (cx - r > 0) bound = fminf(bound, px - (g.lo[0] + (float)(cx - r) * g.h));
    if (cx + r < g.dim[0] - 1) bound = fminf(bound, (g.lo[0] + (float)(cx + r + 1) * g.h) - px);
    if (cy - r > 0) bound = fminf(bound, py - (g.lo[1] + (float)(cy - r) * g.h));
    if (cy + r < g.dim[1] - 1) bound = fminf(bound, (g.lo[1] + (float)(cy + r + 1) * g.h) - py);
    if (cz - r > 0) bound = fminf(bound, pz - (g.lo[2] + (float)(cz - r) * g.h));
    if (cz + r < g.dim[2] - 1) bound = fminf(bound, (g.lo[2] + (float)(cz + r + 1) * g.h) - pz);
    // conservative: shave a few ulps off the bound so that rounding in the face coordinates can never
    // terminate the search early
    bound = bound - 4.0f * 1.1920929e-07f * (fabsf(px) + fabsf(py) + fabsf(pz) + g.h);
    if (bound > 0.f && bd[k - 1] <= bound * bound) return r;
  }
}

}  // namespace sixdgs
