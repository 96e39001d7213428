// a2 / a6 / a7 / a8: candidate-ray generation, one warp per ellipsoid.
//
// Reference: pose_estimation/quadricell.py:86-188 (perimeter, surface, degrade mask),
//            :100-160,191-319 (quadricell cell centres), :322-386 (rotate, mask, rays),
//            pose_estimation/sampling.py:116-124,176-251 (colour), utils/sh_utils.py:55-118,
//            utils/general_utils.py:103-126 (quaternion -> R), scene/gaussian_model.py:125-138.
//
// The reference materialises a [cells, 1000] arc-length table plus nonzero/argsort temporaries
// (232 MB per 1000 ellipsoids), which is why it caps the scene at 1000 ellipsoids.  Here each warp
// builds the 1000-entry table of ONE ring in shared memory (fp64-accumulated like torch's CPU
// cumsum), inverts it by binary search for that ring's cells, and streams the surviving rays out
// with warp-ballot compaction -- no [cells, 1000] temporaries, so the scene size is unbounded.
//
// Every ring's table is built ONCE (round 1 built it twice, in a count and in a fill pass): a cheap
// kernel counts the CELLS of each ellipsoid (ring layout only, no table), a scan turns them into slot
// offsets, the table pass writes each ellipsoid's surviving rays densely at its slot offset and records
// how many survived, and after a second scan a copy kernel moves the per-ellipsoid blocks to their final,
// gap-free positions (36 B/ray of extra traffic against ~8000 sin/cos/sqrt per ellipsoid saved).
//
// Every discrete decision (floor, trunc, strict <, > 0) is computed with the same fp32 operation
// order as the torch expressions; this file is compiled with -fmad=false so nvcc cannot contract
// a*b+c into an FMA and change a rounding.
#include "common.cuh"

namespace sixdgs {

constexpr int kWarpsPerBlock = 8;
constexpr int kTableMax = 1024;  // >= resolution

constexpr float kPiF = 3.14159265358979323846f;      // float(math.pi)
constexpr float kTwoPiF = 6.28318530717958647692f;   // float(2*math.pi)
constexpr float kFourPiF = 12.566370614359172f;      // float(4*math.pi)

__device__ __forceinline__ float perimeter(float b, float c) {
  // pi * ((b+c) + 3(b-c)^2 / (10(b+c) + sqrt(b^2 + 14bc + c^2)))        quadricell.py:86-97
  const float s = b + c;
  const float d = b - c;
  const float num = 3.0f * (d * d);
  const float den = 10.0f * s + sqrtf(b * b + (14.0f * b) * c + c * c);
  return kPiF * (s + num / den);
}

// correctly rounded fp32 pow via fp64 (torch CPU uses a <= 1 ulp vectorised powf)
__device__ __forceinline__ float pow_f(float x, float e) { return (float)pow((double)x, (double)e); }

__device__ __forceinline__ void ring_layout(float a, float b, float c, int target, float& side, long long& rings) {
  const float p = 1.6075f;
  const float ip = 0.62208398133748055987f;  // float(1/1.6075)
  const float acc = pow_f(a * b, p) + pow_f(a * c, p) + pow_f(b * c, p);
  const float surf = kFourPiF * pow_f(acc / 3.0f, ip);
  side = sqrtf(surf / (float)target);
  const float rb = floorf(perimeter(a, b) / (2.0f * side));
  const float rc = floorf(perimeter(a, c) / (2.0f * side));
  const float h = (rb + rc) * 0.5f;
  rings = (h == h && fabsf(h) < 9.0e18f) ? (long long)h : (long long)0x8000000000000000ull;
}

__device__ __forceinline__ float exp_f(float x) { return (float)exp((double)x); }

__global__ void degrade_mask_kernel(const float* __restrict__ scaling_raw, int64_t n, int target,
                                    uint8_t* __restrict__ valid, int32_t* __restrict__ rings_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = exp_f(scaling_raw[i * 3 + 0]);
  const float b = exp_f(scaling_raw[i * 3 + 1]);
  const float c = exp_f(scaling_raw[i * 3 + 2]);
  float side; long long rings;
  ring_layout(a, b, c, target, side, rings);
  valid[i] = rings < (long long)target ? 1 : 0;
  if (rings_out) rings_out[i] = (int32_t)max(min(rings, (long long)INT32_MAX), (long long)INT32_MIN);
}

__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// rgb = max(SH(sh, -dir) + 0.5, 0) with the reference's left-to-right accumulation (sh_utils.py:72-103)
__device__ __forceinline__ float sh_channel(const float* __restrict__ sh, int ch, int deg, float x, float y, float z) {
  const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
  float r = C0 * sh[0 * 3 + ch];
  if (deg > 0) {
    r = r - (C1 * y) * sh[1 * 3 + ch];
    r = r + (C1 * z) * sh[2 * 3 + ch];
    r = r - (C1 * x) * sh[3 * 3 + ch];
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      r = r + (1.0925484305920792f * xy) * sh[4 * 3 + ch];
      r = r + (-1.0925484305920792f * yz) * sh[5 * 3 + ch];
      r = r + (0.31539156525252005f * (2.0f * zz - xx - yy)) * sh[6 * 3 + ch];
      r = r + (-1.0925484305920792f * xz) * sh[7 * 3 + ch];
      r = r + (0.5462742152960396f * (xx - yy)) * sh[8 * 3 + ch];
      if (deg > 2) {
        r = r + ((-0.5900435899266435f * y) * (3.0f * xx - yy)) * sh[9 * 3 + ch];
        r = r + ((2.890611442640554f * xy) * z) * sh[10 * 3 + ch];
        r = r + ((-0.4570457994644658f * y) * (4.0f * zz - xx - yy)) * sh[11 * 3 + ch];
        r = r + ((0.3731763325901154f * z) * (2.0f * zz - 3.0f * xx - 3.0f * yy)) * sh[12 * 3 + ch];
        r = r + ((-0.4570457994644658f * x) * (4.0f * zz - xx - yy)) * sh[13 * 3 + ch];
        r = r + ((1.445305721320277f * z) * (xx - yy)) * sh[14 * 3 + ch];
        r = r + ((-0.5900435899266435f * x) * (xx - 3.0f * yy)) * sh[15 * 3 + ch];
      }
    }
  }
  return fmaxf(r + 0.5f, 0.0f);
}

// cells per ellipsoid = sum over its rings of floor(perimeter(b_r, c_r) / side): the same expressions, in the same
// order, as the ring loop of raygen_kernel (an upper bound of the ellipsoid's rays, exact for mode 1)
__global__ void __launch_bounds__(256)
cells_kernel(const float* __restrict__ scaling_raw, const int64_t* __restrict__ sel, int64_t m, int target,
             int32_t* __restrict__ cells_per_ell) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int64_t gid = sel ? sel[e] : e;
  const float a = exp_f(scaling_raw[gid * 3 + 0]);
  const float b = exp_f(scaling_raw[gid * 3 + 1]);
  const float c = exp_f(scaling_raw[gid * 3 + 2]);
  float side; long long rings_ll;
  ring_layout(a, b, c, target, side, rings_ll);
  const int T = (int)max((long long)0, min(rings_ll, (long long)1 << 20));
  long long cells = 0;
  for (int r = 0; r < T; ++r) {
    const float dr = (2.0f * a) / (float)T;
    const float xc = 0.5f * dr + dr * (float)r;
    const float xa = xc - a;
    const float shrink = 1.0f - (xa * xa) / (a * a);
    const float bs = sqrtf(shrink * (b * b));
    const float cs = sqrtf(shrink * (c * c));
    const float nf = floorf(perimeter(bs, cs) / side);
    if (!(nf >= 1.0f) || nf > 1.0e6f) continue;
    cells += (int)nf;
  }
  cells_per_ell[e] = (int32_t)min(cells, (long long)INT32_MAX);
}

// out[ray_off[e] + i] = tmp[slot_off[e] + i] for i < rays_per_ell[e]: one warp per ellipsoid, coalesced
__global__ void __launch_bounds__(256)
compact_kernel(const int64_t* __restrict__ slot_off, const int64_t* __restrict__ ray_off, const int32_t* __restrict__ rays_per_ell,
               int64_t m, const float* __restrict__ t_ori, const float* __restrict__ t_dir, const float* __restrict__ t_rgb,
               const int64_t* __restrict__ t_ell, float* __restrict__ ori, float* __restrict__ dir, float* __restrict__ rgb,
               int64_t* __restrict__ ell) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < m; e += warps_total) {
    const int n = rays_per_ell[e];
    const int64_t src = slot_off[e], dst = ray_off[e];
    for (int i = lane; i < n * 3; i += 32) {
      ori[dst * 3 + i] = t_ori[src * 3 + i];
      if (dir) dir[dst * 3 + i] = t_dir[src * 3 + i];
      if (rgb) rgb[dst * 3 + i] = t_rgb[src * 3 + i];
    }
    if (ell)
      for (int i = lane; i < n; i += 32) ell[dst + i] = t_ell[src + i];
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
raygen_kernel(const float* __restrict__ xyz, const float* __restrict__ scaling_raw,
              const float* __restrict__ rotation_raw, const float* __restrict__ features, int sh_degree, int sh_coeffs,
              const int64_t* __restrict__ sel, int64_t m, const float* __restrict__ normals, int target,
              int resolution, int mode, const int64_t* __restrict__ ray_offset,
              int32_t* __restrict__ rays_per_ell, int32_t* __restrict__ cells_per_ell,
              float* __restrict__ ori, float* __restrict__ dir, float* __restrict__ rgb,
              int64_t* __restrict__ ell_id) {
  __shared__ float s_table[kWarpsPerBlock][kTableMax];
  __shared__ float s_sh[kWarpsPerBlock][48];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* table = s_table[warp];
  float* shc = s_sh[warp];
  const int64_t warps_total = (int64_t)gridDim.x * kWarpsPerBlock;
  const int nsamp = resolution - 1;  // table samples per ring (999)

  for (int64_t e = (int64_t)blockIdx.x * kWarpsPerBlock + warp; e < m; e += warps_total) {
    const int64_t gid = sel ? sel[e] : e;
    const float a = exp_f(scaling_raw[gid * 3 + 0]);
    const float b = exp_f(scaling_raw[gid * 3 + 1]);
    const float c = exp_f(scaling_raw[gid * 3 + 2]);
    float side; long long rings_ll;
    ring_layout(a, b, c, target, side, rings_ll);
    const int T = (int)max((long long)0, min(rings_ll, (long long)1 << 20));

    // rotation: F.normalize then build_rotation's own normalisation (gaussian_model.py:133, general_utils.py:103-126)
    float R[9], mu[3] = {0.f, 0.f, 0.f}, nx = 0.f;
    if (mode == 0) {
      float q0 = rotation_raw[gid * 4 + 0], q1 = rotation_raw[gid * 4 + 1];
      float q2 = rotation_raw[gid * 4 + 2], q3 = rotation_raw[gid * 4 + 3];
      const float n1 = fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);
      q0 /= n1; q1 /= n1; q2 /= n1; q3 /= n1;
      const float n2 = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
      const float w = q0 / n2, x = q1 / n2, y = q2 / n2, z = q3 / n2;
      R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
      R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
      R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
      mu[0] = xyz[gid * 3 + 0]; mu[1] = xyz[gid * 3 + 1]; mu[2] = xyz[gid * 3 + 2];
      nx = normals[e * 3 + 0];
      {
        __syncwarp();
        // features is [N, sh_coeffs, 3] (get_features: 1 dc + rest coefficients, gaussian_model.py:136-140); only the
        // first (deg+1)^2 coefficients are read by eval_sh, whatever the stored count
        const int used = (sh_degree + 1) * (sh_degree + 1) * 3;
        for (int i = lane; i < used; i += 32) shc[i] = features[gid * (int64_t)(sh_coeffs * 3) + i];
        __syncwarp();
      }
    }
    const int64_t base = ray_offset[e];
    int kept_total = 0;
    long long cells_total = 0;

    for (int r = 0; r < T; ++r) {
      // slab geometry (quadricell.py:100-105,128-148)
      const float dr = (2.0f * a) / (float)T;
      const float xc = 0.5f * dr + dr * (float)r;
      const float xa = xc - a;
      const float shrink = 1.0f - (xa * xa) / (a * a);
      const float bs = sqrtf(shrink * (b * b));
      const float cs = sqrtf(shrink * (c * c));
      const float nf = floorf(perimeter(bs, cs) / side);
      if (!(nf >= 1.0f) || nf > 1.0e6f) continue;  // zero cells (or NaN): ring contributes nothing
      const int n = (int)nf;
      cells_total += n;
      const float dth = kTwoPiF / nf;
      const float pz = (0.5f * dr + dr * (float)r) - a;

      // arc-length table: I_0 = 0, I_{k+1} = I_k + sqrt(bs sin^2 + cs cos^2) * dth  (fp64 running sum,
      // each entry rounded to fp32 -- torch CPU cumsum semantics), then 2*pi * I / I_last.
      __syncwarp();
      double carry = 0.0;
      if (lane == 0) table[0] = 0.0f;
      for (int k0 = 0; k0 < nsamp; k0 += 32) {
        const int k = k0 + lane;
        double v = 0.0;
        if (k < nsamp) {
          const float th = (float)k * dth;
          float sn, cn;
          sn = sinf(th); cn = cosf(th);
          const float ds = sqrtf(bs * (sn * sn) + cs * (cn * cn));
          v = (double)(ds * dth);
        }
        const double inc = warp_scan_incl(v, lane);
        if (k < nsamp) table[k + 1] = (float)(carry + inc);
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      __syncwarp();
      const float last = table[nsamp];
      __syncwarp();
      for (int k = lane; k <= nsamp; k += 32) table[k] = kTwoPiF * (table[k] / last);
      __syncwarp();

      for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool act = j < n;
        bool keep = false;
        float px = 0.f, py = 0.f, rx = 0.f, ry = 0.f, rz = 0.f;
        if (act) {
          const float thj = (float)j * dth;
          // cnt = #{i in 1..nsamp : table[i] < thj}; theta' = table[max(cnt-1, 0)]  (quadricell.py:286-299)
          int lo = 0, hi = nsamp;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (table[1 + mid] < thj) lo = mid + 1; else hi = mid;
          }
          const float thp = table[lo > 0 ? lo - 1 : 0];
          px = bs * cosf(thp);
          py = cs * sinf(thp);
          if (mode == 0) {
            rx = R[0] * px + R[1] * py + R[2] * pz;
            ry = R[3] * px + R[4] * py + R[5] * pz;
            rz = R[6] * px + R[7] * py + R[8] * pz;
            keep = (nx * rx) > 0.0f;  // outer-product quirk: only the x components (quadricell.py:337-340)
          } else {
            keep = true;
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int64_t o = base + kept_total + __popc(bal & ((1u << lane) - 1u));
          if (mode == 0) {
            const float nn = fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), 1e-12f);
            const float dx = rx / nn, dy = ry / nn, dz = rz / nn;
            ori[o * 3 + 0] = rx + mu[0]; ori[o * 3 + 1] = ry + mu[1]; ori[o * 3 + 2] = rz + mu[2];
            dir[o * 3 + 0] = dx; dir[o * 3 + 1] = dy; dir[o * 3 + 2] = dz;
            rgb[o * 3 + 0] = sh_channel(shc, 0, sh_degree, -dx, -dy, -dz);
            rgb[o * 3 + 1] = sh_channel(shc, 1, sh_degree, -dx, -dy, -dz);
            rgb[o * 3 + 2] = sh_channel(shc, 2, sh_degree, -dx, -dy, -dz);
          } else {
            ori[o * 3 + 0] = px; ori[o * 3 + 1] = py; ori[o * 3 + 2] = pz;
          }
          if (ell_id) ell_id[o] = e;
        }
        kept_total += __popc(bal);
      }
    }
    if (lane == 0) {
      if (rays_per_ell) rays_per_ell[e] = kept_total;
      if (cells_per_ell) cells_per_ell[e] = (int32_t)min(cells_total, (long long)INT32_MAX);
    }
  }
}

// single-block exclusive scan int32 -> int64 (n up to ~1e7 is a few hundred microseconds; this runs
// once per scene preparation)
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    long long v = (i < n) ? (long long)in[i] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      long long w = s_warp[lane];
      long long winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - w;  // exclusive warp offsets
    }
    __syncthreads();
    const long long carry = s_carry;
    if (i < n) out[i] = carry + s_warp[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[warp] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = s_carry;
}

}  // namespace sixdgs

using namespace sixdgs;

extern "C" int sixdgs_degrade_mask(const float* scaling_raw, int64_t n, int target_points, uint8_t* valid,
                                   int32_t* rings_out, void* stream) {
  SIXDGS_REQUIRE(scaling_raw && valid, "null pointer");
  SIXDGS_REQUIRE(n >= 0 && target_points > 0, "bad size");
  if (n == 0) return SIXDGS_OK;
  degrade_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(scaling_raw, n, target_points,
                                                                                      valid, rings_out);
  return check_launch("degrade_mask");
}

static unsigned raygen_grid(int64_t m) {
  const int64_t want = (m + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)kNumSMs * 8;  // persistent: 8 CTAs of 8 warps per SM
  return (unsigned)(want < cap ? want : cap);
}

extern "C" int sixdgs_raygen_cells(const float* scaling_raw, const int64_t* sel, int64_t m, int target_points,
                                   int32_t* cells_per_ell, void* stream) {
  SIXDGS_REQUIRE(scaling_raw && cells_per_ell, "null pointer");
  SIXDGS_REQUIRE(m >= 0 && target_points > 0, "bad size");
  if (m == 0) return SIXDGS_OK;
  cells_kernel<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(scaling_raw, sel, m, target_points, cells_per_ell);
  return check_launch("raygen_cells");
}

extern "C" int sixdgs_raygen_fill(const float* xyz, const float* scaling_raw, const float* rotation_raw,
                                  const float* features, int sh_degree, int sh_coeffs, const int64_t* sel, int64_t m,
                                  const float* normals, int target_points, int resolution, int mode,
                                  const int64_t* slot_offset, float* ori, float* dir, float* rgb, int64_t* ell_id,
                                  int32_t* rays_per_ell, void* stream) {
  SIXDGS_REQUIRE(scaling_raw && slot_offset && ori, "null pointer");
  SIXDGS_REQUIRE(mode == 1 || (xyz && rotation_raw && normals && features && dir && rgb),
                 "mode 0 needs xyz, rotation, normals, features, dir and rgb");
  SIXDGS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
  SIXDGS_REQUIRE(mode == 1 || sh_coeffs >= (sh_degree + 1) * (sh_degree + 1), "features hold fewer than (sh_degree+1)^2 coefficients");
  SIXDGS_REQUIRE(resolution >= 2 && resolution <= kTableMax, "resolution must be in [2, 1024]");
  SIXDGS_REQUIRE(m >= 0 && target_points > 0, "bad size");
  if (m == 0) return SIXDGS_OK;
  raygen_kernel<<<raygen_grid(m), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      xyz, scaling_raw, rotation_raw, features, sh_degree, sh_coeffs, sel, m, normals, target_points, resolution, mode,
      slot_offset, rays_per_ell, nullptr, ori, dir, rgb, ell_id);
  return check_launch("raygen_fill");
}

extern "C" int sixdgs_raygen_compact(const int64_t* slot_offset, const int64_t* ray_offset, const int32_t* rays_per_ell,
                                     int64_t m, const float* tmp_ori, const float* tmp_dir, const float* tmp_rgb,
                                     const int64_t* tmp_ell, float* ori, float* dir, float* rgb, int64_t* ell_id,
                                     void* stream) {
  SIXDGS_REQUIRE(slot_offset && ray_offset && rays_per_ell && tmp_ori && ori, "null pointer");
  SIXDGS_REQUIRE((!dir || tmp_dir) && (!rgb || tmp_rgb) && (!ell_id || tmp_ell), "output without its source");
  SIXDGS_REQUIRE(m >= 0, "bad size");
  if (m == 0) return SIXDGS_OK;
  const int64_t want = (m + 7) / 8;
  const unsigned grid = (unsigned)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
  compact_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(slot_offset, ray_offset, rays_per_ell, m, tmp_ori, tmp_dir, tmp_rgb,
                                                         tmp_ell, ori, dir, rgb, ell_id);
  return check_launch("raygen_compact");
}

extern "C" int sixdgs_exclusive_scan(const int32_t* in, int64_t n, int64_t* out, void* stream) {
  SIXDGS_REQUIRE(in && out, "null pointer");
  SIXDGS_REQUIRE(n >= 0, "negative size");
  scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(in, n, out);
  return check_launch("exclusive_scan");
}
