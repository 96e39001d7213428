// Host build of normal_fit.cuh (+ eig3.cuh): exact k-NN by brute force, then the device routine that turns a
// neighbourhood into a normal.  Driven from tests/test_device_code_on_host.py against the reference fixture.
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <utility>
#include <vector>
#define __device__
#define __forceinline__ inline
#define __host__
#define __restrict__
#include "normal_fit.cuh"

extern "C" void host_knn_normals(const float* cloud, int64_t m, int64_t n_query, int k, float* normals) {
  std::vector<std::pair<float, int>> d((size_t)m);
  std::vector<int> bi((size_t)k);
  for (int64_t q = 0; q < n_query; ++q) {
    const float* a = cloud + q * 3;
    for (int64_t j = 0; j < m; ++j) {
      const float* b = cloud + j * 3;
      const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
      d[(size_t)j] = std::make_pair(dx * dx + dy * dy + dz * dz, (int)j);
    }
    std::partial_sort(d.begin(), d.begin() + k, d.end());  // (distance, index): ties to the lower index
    for (int i = 0; i < k; ++i) bi[(size_t)i] = d[(size_t)i].second;
    sixdgs::normal_from_neighbours(cloud, bi.data(), k, normals + q * 3);
  }
}
