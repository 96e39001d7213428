#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <math.h>
using std::max; using std::min;
#include "knn_grid.cuh"
using namespace sixdgs;
int main() {
  for (int trial = 0; trial < 8; ++trial) {
    const int m = (trial < 3 || trial >= 6) ? 3000 : 20000, k = 20;
    srand(trial + 1);
    std::vector<float> c(m * 3);
    auto rnd = []() { float s = 0; for (int i = 0; i < 6; ++i) s += rand() / (float)RAND_MAX; return (s - 3.f); };
    for (int i = 0; i < m; ++i) { c[i*3] = rnd() * (trial==1?10.f:1.f); c[i*3+1] = rnd(); c[i*3+2] = rnd() * (trial==2?0.01f:1.f); }
    if (trial == 4) for (int i = 0; i < 200; ++i) { c[i*3] = c[(i+200)*3]; c[i*3+1] = c[(i+200)*3+1]; c[i*3+2] = c[(i+200)*3+2]; }  // duplicates
    if (trial >= 6) for (int i = 0; i < 150; ++i) for (int a = 0; a < 3; ++a) c[i*3+a] *= (trial == 6 ? 50.f : 1000.f);  // far outliers
    KnnGrid g; float lo[3] = {1e30f,1e30f,1e30f}, hi[3] = {-1e30f,-1e30f,-1e30f};
    for (int i = 0; i < m; ++i) for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], c[i*3+a]); hi[a] = std::max(hi[a], c[i*3+a]); }
    if (trial >= 6) for (int a = 0; a < 3; ++a) { lo[a] = std::max(lo[a], -3.0f); hi[a] = std::min(hi[a], 3.0f); }  // robust box: outliers get clamped
    double vol = 1; for (int a = 0; a < 3; ++a) vol *= std::max(hi[a]-lo[a], 1e-6f);
    g.h = (float)cbrt(vol / (m / 8.0));
    for (int a = 0; a < 3; ++a) { g.lo[a] = lo[a]; g.dim[a] = std::max(1, (int)ceilf((hi[a]-lo[a]) / g.h + 1e-3f)); if (g.dim[a] > 256) g.dim[a] = 256; }
    for (int a = 0; a < 3; ++a) if ((hi[a]-lo[a]) / g.h > g.dim[a]) g.h = std::max(g.h, (hi[a]-lo[a]) / g.dim[a] * 1.0001f);
    const int64_t nc = (int64_t)g.dim[0]*g.dim[1]*g.dim[2];
    std::vector<int64_t> start(nc + 1, 0); std::vector<int> cell(m), sorted(m);
    for (int i = 0; i < m; ++i) { int x = knn_cell_coord(c[i*3], g.lo[0], g.h, g.dim[0]), y = knn_cell_coord(c[i*3+1], g.lo[1], g.h, g.dim[1]), z = knn_cell_coord(c[i*3+2], g.lo[2], g.h, g.dim[2]); cell[i] = (z*g.dim[1]+y)*g.dim[0]+x; start[cell[i]+1]++; }
    for (int64_t i = 0; i < nc; ++i) start[i+1] += start[i];
    std::vector<int64_t> cur(start.begin(), start.end()-1);
    for (int i = 0; i < m; ++i) sorted[cur[cell[i]]++] = i;   // ascending index inside a cell
    int bad = 0, maxr = 0;
    for (int q = 0; q < m; ++q) {
      float bd[32]; int bi[32];
      int r = knn_grid_query(c.data(), g, start.data(), sorted.data(), c[q*3], c[q*3+1], c[q*3+2], k, bd, bi, -1);
      maxr = std::max(maxr, r);
      // brute force with the same (d, idx) order
      float rd[32]; int ri[32]; for (int i = 0; i < k; ++i) { rd[i] = INFINITY; ri[i] = 0x7fffffff; }
      for (int j = 0; j < m; ++j) { float dx = c[q*3]-c[j*3], dy = c[q*3+1]-c[j*3+1], dz = c[q*3+2]-c[j*3+2]; float d = dx*dx+dy*dy+dz*dz;
        if (d < rd[k-1] || (d == rd[k-1] && j < ri[k-1])) knn_insert(rd, ri, k, d, j); }
      for (int i = 0; i < k; ++i) if (bi[i] != ri[i]) { ++bad; break; }
    }
    printf("trial %d: m=%d grid %dx%dx%d h=%.4f mismatching queries %d, max shell %d\n", trial, m, g.dim[0], g.dim[1], g.dim[2], g.h, bad, maxr);
  }
}
