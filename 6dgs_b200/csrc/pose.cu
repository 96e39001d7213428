// a13 / a14: least-squares line intersection and the pose tail.
// Reference: pose_estimation/line_intersection.py:75-154 (impl2), :29-34 (exclude_negatives),
//            :5-26 (make_rotation_mat); pose_estimation/test.py:157-198 (dedup, LS x2, watch dir,
//            rotation, singular -> identity).
// The 12-float normal-equation accumulator is reduced in fp64 (warp shuffles + one atomic per
// block), rounded to fp32, and solved in fp32 by LU with partial pivoting (what LAPACK's sgesv, used
// by torch.linalg.solve on CPU, does); det(R) comes from the same LU like torch.linalg.det.
#include "common.cuh"

namespace sixdgs {

// LU with partial pivoting of a 3x3 (row-major, in place).  Returns det; perm holds the row order.
__device__ inline float lu3(float* A, int* perm) {
  float det = 1.0f;
  perm[0] = 0; perm[1] = 1; perm[2] = 2;
  for (int c = 0; c < 3; ++c) {
    int p = c; float best = fabsf(A[perm[c] * 3 + c]);
    for (int r = c + 1; r < 3; ++r) {
      const float v = fabsf(A[perm[r] * 3 + c]);
      if (v > best) { best = v; p = r; }
    }
    if (p != c) { const int t = perm[c]; perm[c] = perm[p]; perm[p] = t; det = -det; }
    const float piv = A[perm[c] * 3 + c];
    det *= piv;
    for (int r = c + 1; r < 3; ++r) {
      const float f = A[perm[r] * 3 + c] / piv;
      A[perm[r] * 3 + c] = f;
      for (int cc = c + 1; cc < 3; ++cc) A[perm[r] * 3 + cc] -= f * A[perm[c] * 3 + cc];
    }
  }
  return det;
}

__device__ inline void lu3_solve(const float* LU, const int* perm, const float* b, float* x) {
  float y[3];
  for (int r = 0; r < 3; ++r) {
    float v = b[perm[r]];
    for (int c = 0; c < r; ++c) v -= LU[perm[r] * 3 + c] * y[c];
    y[r] = v;
  }
  for (int r = 2; r >= 0; --r) {
    float v = y[r];
    for (int c = r + 1; c < 3; ++c) v -= LU[perm[r] * 3 + c] * x[c];
    x[r] = v / LU[perm[r] * 3 + r];
  }
}

// centre = solve(R, q) with the reference's det < 1e-7 -> NaN guard.  Returns status bit.
__device__ inline int solve_centre(const double* acc /* R(9) q(3) */, float* centre) {
  float R[9], q[3];
  for (int i = 0; i < 9; ++i) R[i] = (float)acc[i];
  for (int i = 0; i < 3; ++i) q[i] = (float)acc[9 + i];
  int perm[3];
  const float det = lu3(R, perm);
  if (det < 1.0e-7f || det != det) {
    centre[0] = centre[1] = centre[2] = __int_as_float(0x7fc00000);
    return 1;
  }
  lu3_solve(R, perm, q, centre);
  return 0;
}

__device__ __forceinline__ void accumulate_ray(double* a, float ox, float oy, float oz, float dx, float dy, float dz,
                                               float w) {
  // P = I - d d^T ; R += w P ; q += w P o
  const float P[9] = {1.f - dx * dx, -dx * dy, -dx * dz, -dy * dx, 1.f - dy * dy, -dy * dz,
                      -dz * dx, -dz * dy, 1.f - dz * dz};
  const float pq[3] = {P[0] * ox + P[1] * oy + P[2] * oz, P[3] * ox + P[4] * oy + P[5] * oz,
                       P[6] * ox + P[7] * oy + P[8] * oz};
#pragma unroll
  for (int i = 0; i < 9; ++i) a[i] += (double)(P[i] * w);
#pragma unroll
  for (int i = 0; i < 3; ++i) a[9 + i] += (double)(pq[i] * w);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
ls_accumulate_kernel(const float* __restrict__ pts, const float* __restrict__ dirs, const float* __restrict__ w,
                     int64_t n, double* __restrict__ acc_out) {
  double a[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) a[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    accumulate_ray(a, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2],
                   w ? w[i] : 1.0f);
  __shared__ double sh[8][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const double v = warp_sum_d(a[i]);
    if (lane == 0) sh[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double v = 0.0;
    for (int wv = 0; wv < 8; ++wv) v += sh[wv][threadIdx.x];
    atomicAdd(&acc_out[threadIdx.x], v);
  }
}

__global__ void ls_solve_kernel(const double* __restrict__ acc, float* __restrict__ centre, int32_t* __restrict__ status) {
  if (threadIdx.x != 0) return;
  float c[3];
  const int st = solve_centre(acc, c);
  centre[0] = c[0]; centre[1] = c[1]; centre[2] = c[2];
  if (status) *status = st;
}

// make_rotation_mat(direction = -watch, up): rows [x; y; direction] (line_intersection.py:5-26); c2w = [inv(R) | c]
// with the reference's guards (test.py:194-198): det < 1e-7 -> identity rotation (status bit 1), any NaN in the result
// -> identity pose (status bit 2).  Returns the updated status.
__device__ inline int frame_to_c2w(const float* c, const float* watch, const float* up, int status, float* out) {
  const float dx = -watch[0], dy = -watch[1], dz = -watch[2];
  const float ux = up[0], uy = up[1], uz = up[2];
  float xx = uy * dz - uz * dy, xy = uz * dx - ux * dz, xz = ux * dy - uy * dx;
  const float xn = sqrtf(xx * xx + xy * xy + xz * xz);
  xx /= xn; xy /= xn; xz /= xn;
  float yx = dy * xz - dz * xy, yy = dz * xx - dx * xz, yz = dx * xy - dy * xx;
  const float yn = sqrtf(yx * yx + yy * yy + yz * yz);
  yx /= yn; yy /= yn; yz /= yn;
  float M[9] = {xx, xy, xz, yx, yy, yz, dx, dy, dz};
  float det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
              M[2] * (M[3] * M[7] - M[4] * M[6]);
  float inv[9];
  if (det < 1.0e-7f) {  // NaN compares false, like the reference: NaN falls through to the c2w check
    status |= 2;
    for (int i = 0; i < 9; ++i) inv[i] = (i % 4 == 0) ? 1.f : 0.f;
  } else {
    const float id = 1.0f / det;
    inv[0] = (M[4] * M[8] - M[5] * M[7]) * id; inv[1] = (M[2] * M[7] - M[1] * M[8]) * id; inv[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    inv[3] = (M[5] * M[6] - M[3] * M[8]) * id; inv[4] = (M[0] * M[8] - M[2] * M[6]) * id; inv[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    inv[6] = (M[3] * M[7] - M[4] * M[6]) * id; inv[7] = (M[1] * M[6] - M[0] * M[7]) * id; inv[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  }
  const float o[16] = {inv[0], inv[1], inv[2], c[0], inv[3], inv[4], inv[5], c[1],
                       inv[6], inv[7], inv[8], c[2], 0.f, 0.f, 0.f, 1.f};
  bool bad = false;
  for (int i = 0; i < 16; ++i) bad |= (o[i] != o[i]);
  if (bad) status |= 4;
  for (int i = 0; i < 16; ++i) out[i] = bad ? ((i % 5 == 0) ? 1.f : 0.f) : o[i];
  return status;
}

// All-ray weighted least squares (least_squared_loss.py:62-64, line_intersection.py:75-154) from the 13 sums the
// pass-2 epilogue of score_tc_mq.cu accumulates: sys = (R xx,xy,xz,yy,yz,zz | q | sum w d | sum w), every sum scaled
// by `scale` (1 / n_img: weights = score / n_img).  centre = solve(R, q) through the same fp32 LU + det < 1e-7 guard
// as the top-k path; watch = normalise(sum w d); c2w (nullable) from (centre, watch, up) like test.py:187-198.
__global__ void ls_system_solve_kernel(const double* __restrict__ sys, int n, double scale, const float* __restrict__ up,
                                       float* __restrict__ centre, float* __restrict__ watch, float* __restrict__ c2w,
                                       float* __restrict__ aux, int32_t* __restrict__ status_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const double* s = sys + (int64_t)b * 13;
  const double acc[12] = {s[0] * scale, s[1] * scale, s[2] * scale, s[1] * scale, s[3] * scale, s[4] * scale,
                          s[2] * scale, s[4] * scale, s[5] * scale, s[6] * scale, s[7] * scale, s[8] * scale};
  float c[3];
  int status = solve_centre(acc, c);
  const double n2 = sqrt(s[9] * s[9] + s[10] * s[10] + s[11] * s[11]);
  const float wv[3] = {(float)(s[9] / n2), (float)(s[10] / n2), (float)(s[11] / n2)};
  if (centre) { centre[b * 3] = c[0]; centre[b * 3 + 1] = c[1]; centre[b * 3 + 2] = c[2]; }
  if (watch) { watch[b * 3] = wv[0]; watch[b * 3 + 1] = wv[1]; watch[b * 3 + 2] = wv[2]; }
  if (c2w) {
    float out[16];
    status = frame_to_c2w(c, wv, up + b * 3, status, out);
    for (int i = 0; i < 16; ++i) c2w[b * 16 + i] = out[i];
  }
  if (aux) {
    float* a = aux + b * 8;
    a[0] = c[0]; a[1] = c[1]; a[2] = c[2]; a[3] = wv[0]; a[4] = wv[1]; a[5] = wv[2]; a[6] = (float)(s[12] * scale);
    a[7] = (float)status;
  }
  if (status_out) status_out[b] = status;
}

constexpr int kPoseMaxK = 1024;

__global__ void __launch_bounds__(128)
pose_tail_kernel(const float* __restrict__ rays_ori, const float* __restrict__ rays_dir, int64_t ray_stride,
                 const int64_t* __restrict__ idx, const float* __restrict__ vals, int k, const float* __restrict__ up,
                 float* __restrict__ c2w, float* __restrict__ aux) {
  __shared__ float so[kPoseMaxK][3], sd[kPoseMaxK][3], sw[kPoseMaxK];
  __shared__ unsigned char once[kPoseMaxK], keep[kPoseMaxK];
  __shared__ int s_n;
  __shared__ double s_acc[16];
  __shared__ float s_c[3];
  __shared__ int s_status;
  const int t = threadIdx.x;
  for (int i = t; i < k; i += blockDim.x) {
    const int64_t r = idx[i];
    so[i][0] = rays_ori[r * ray_stride]; so[i][1] = rays_ori[r * ray_stride + 1]; so[i][2] = rays_ori[r * ray_stride + 2];
    sd[i][0] = rays_dir[r * ray_stride]; sd[i][1] = rays_dir[r * ray_stride + 1]; sd[i][2] = rays_dir[r * ray_stride + 2];
    sw[i] = vals[i];
  }
  __syncthreads();
  // torch.unique(rows, return_counts) -> rows seen exactly once                     test.py:157
  for (int i = t; i < k; i += blockDim.x) {
    int cnt = 0;
    for (int j = 0; j < k; ++j)
      cnt += (so[j][0] == so[i][0] && so[j][1] == so[i][1] && so[j][2] == so[i][2]) ? 1 : 0;
    once[i] = (cnt == 1);
  }
  __syncthreads();
  // mask = torch.isin(rows, unique_once_rows, assume_unique=True).any(dim=1)             test.py:158-160
  // Reference quirk, reproduced: isin works ELEMENT-wise on the flattened coordinates, and with
  // assume_unique=True torch takes its sort-based path (stable sort of [elements ; test_elements],
  // flag = "next sorted entry is equal") whenever test.numel() >= 10 * elements.numel()^0.145
  // (ATen TensorCompare.cpp isin_Tensor_Tensor_out).  A coordinate is therefore flagged iff an equal
  // value occurs LATER in the concatenation: in a later flattened position of the rows themselves, or
  // anywhere in the once-only rows.  Net effect on duplicated origins: the first copy survives, later
  // copies are dropped.  With very few once-only rows torch uses plain membership instead.
  __shared__ int s_once;
  if (t == 0) {
    int c = 0;
    for (int i = 0; i < k; ++i) c += once[i];
    s_once = c;
  }
  __syncthreads();
  const bool sorting_path = (float)(3 * s_once) >= 10.0f * powf((float)(3 * k), 0.145f);
  for (int i = t; i < k; i += blockDim.x) {
    bool kp = false;
    for (int a = 0; a < 3 && !kp; ++a) {
      const float v = so[i][a];
      const int f = 3 * i + a;
      for (int j = 0; j < k && !kp; ++j)
        for (int b = 0; b < 3; ++b)
          if (so[j][b] == v && (once[j] || (sorting_path && (3 * j + b) > f))) { kp = true; break; }
    }
    keep[i] = kp;
  }
  __syncthreads();
  if (t == 0) {  // order-preserving compaction
    int n = 0;
    for (int i = 0; i < k; ++i)
      if (keep[i]) {
        if (n != i) {
          so[n][0] = so[i][0]; so[n][1] = so[i][1]; so[n][2] = so[i][2];
          sd[n][0] = sd[i][0]; sd[n][1] = sd[i][1]; sd[n][2] = sd[i][2];
          sw[n] = sw[i];
        }
        ++n;
      }
    s_n = n;
    s_status = 0;
  }
  __syncthreads();
  const int n = s_n;
  // unweighted LS over the kept rays (weights argument is commented out upstream, test.py:169-179)
  if (t < 32) {
    double a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = 0.0;
    double wsum = 0.0;
    for (int i = t; i < n; i += 32) {
      accumulate_ray(a, so[i][0], so[i][1], so[i][2], sd[i][0], sd[i][1], sd[i][2], 1.0f);
      wsum += (double)sw[i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = warp_sum_d(a[i]);
    wsum = warp_sum_d(wsum);
    if (t == 0) {
      for (int i = 0; i < 12; ++i) s_acc[i] = a[i];
      s_acc[12] = wsum;
      float c[3];
      s_status |= solve_centre(s_acc, c);
      s_c[0] = c[0]; s_c[1] = c[1]; s_c[2] = c[2];
    }
  }
  __syncthreads();
  // w <- (w / sum w) * [ (c - o) . d > 0 ] ; renormalise ; watch = normalise(sum w d)
  if (t < 32) {
    const float wsum = (float)s_acc[12];
    double s2 = 0.0;
    for (int i = t; i < n; i += 32) {
      const float vx = s_c[0] - so[i][0], vy = s_c[1] - so[i][1], vz = s_c[2] - so[i][2];
      const float pr = vx * sd[i][0] + vy * sd[i][1] + vz * sd[i][2];
      const float w1 = (sw[i] / wsum) * ((pr > 0.f) ? 1.0f : 0.0f);
      sw[i] = w1;
      s2 += (double)w1;
    }
    s2 = warp_sum_d(s2);
    const float w2sum = (float)s2;
    double wx = 0.0, wy = 0.0, wz = 0.0;
    for (int i = t; i < n; i += 32) {
      const float w2 = sw[i] / w2sum;
      wx += (double)(sd[i][0] * w2); wy += (double)(sd[i][1] * w2); wz += (double)(sd[i][2] * w2);
    }
    wx = warp_sum_d(wx); wy = warp_sum_d(wy); wz = warp_sum_d(wz);
    if (t == 0) {
      float fx = (float)wx, fy = (float)wy, fz = (float)wz;
      const float fn = sqrtf(fx * fx + fy * fy + fz * fz);
      fx /= fn; fy /= fn; fz /= fn;
      if (aux) { aux[0] = s_c[0]; aux[1] = s_c[1]; aux[2] = s_c[2]; aux[3] = fx; aux[4] = fy; aux[5] = fz; aux[6] = (float)n; }
      const float wv[3] = {fx, fy, fz};
      float out[16];
      const int status = frame_to_c2w(s_c, wv, up, s_status, out);
      for (int i = 0; i < 16; ++i) c2w[i] = out[i];
      if (aux) aux[7] = (float)status;
    }
  }
}

}  // namespace sixdgs

using namespace sixdgs;

extern "C" int sixdgs_line_intersect(const float* points, const float* dirs, const float* weights, int64_t n,
                                     float* centre, int32_t* status, void* workspace, void* stream) {
  SIXDGS_REQUIRE(points && dirs && centre && workspace, "null pointer (workspace: 12 doubles)");
  SIXDGS_REQUIRE(n >= 0, "negative size");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(workspace, 0, 12 * sizeof(double), s);
  if (e != cudaSuccess) { set_error("line_intersect memset: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  if (n > 0) {
    const int64_t want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < kNumSMs * 4 ? want : kNumSMs * 4);
    ls_accumulate_kernel<<<grid, 256, 0, s>>>(points, dirs, weights, n, (double*)workspace);
  }
  ls_solve_kernel<<<1, 32, 0, s>>>((const double*)workspace, centre, status);
  return check_launch("line_intersect");
}

extern "C" int sixdgs_ls_solve(const double* ls_sys, int n, double weight_scale, const float* up, float* centre, float* watch,
                               float* c2w, float* aux, int32_t* status, void* stream) {
  SIXDGS_REQUIRE(ls_sys && (centre || c2w), "null pointer");
  SIXDGS_REQUIRE(n >= 1, "bad size");
  SIXDGS_REQUIRE(!c2w || up, "c2w needs the camera-up vectors");
  ls_system_solve_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(ls_sys, n, weight_scale, up, centre, watch, c2w, aux,
                                                                          status);
  return check_launch("ls_solve");
}

extern "C" int sixdgs_pose_tail(const float* rays_ori, const float* rays_dir, int64_t ray_stride, const int64_t* idx,
                                const float* vals, int k, const float* up, float* c2w, float* aux, void* stream) {
  SIXDGS_REQUIRE(rays_ori && rays_dir && idx && vals && up && c2w, "null pointer");
  SIXDGS_REQUIRE(k >= 1 && k <= kPoseMaxK, "k must be in [1, 1024]");
  SIXDGS_REQUIRE(ray_stride >= 3, "ray_stride must be >= 3 floats");
  pose_tail_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(rays_ori, rays_dir, ray_stride, idx, vals, k, up, c2w, aux);
  return check_launch("pose_tail");
}

// candidate rows for the cross-rank top-k exchange: out[i] = (score, ori3, dir3) of the i-th local winner,
// rows i >= k_local are (-inf, 0...) so they can never be selected
__global__ void gather_candidates_kernel(const float* __restrict__ vals, const int64_t* __restrict__ idx, int k_local, int k,
                                         const float* __restrict__ ori, const float* __restrict__ dir,
                                         float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  float* o = out + (int64_t)i * 7;
  if (i < k_local) {
    const int64_t r = idx[i];
    o[0] = vals[i];
    o[1] = ori[r * 3]; o[2] = ori[r * 3 + 1]; o[3] = ori[r * 3 + 2];
    o[4] = dir[r * 3]; o[5] = dir[r * 3 + 1]; o[6] = dir[r * 3 + 2];
  } else {
    o[0] = -INFINITY;
    for (int j = 1; j < 7; ++j) o[j] = 0.f;
  }
}

extern "C" int sixdgs_gather_candidates(const float* vals, const int64_t* idx, int k_local, int k, const float* rays_ori,
                                        const float* rays_dir, float* out, void* stream) {
  SIXDGS_REQUIRE(out && (k_local == 0 || (vals && idx && rays_ori && rays_dir)), "null pointer");
  SIXDGS_REQUIRE(k >= 1 && k_local >= 0 && k_local <= k, "bad k");
  gather_candidates_kernel<<<(k + 127) / 128, 128, 0, (cudaStream_t)stream>>>(vals, idx, k_local, k, rays_ori, rays_dir, out);
  return check_launch("gather_candidates");
}
