// a10 (+ k_proj), EXACT tensor-core build of the f16x2 key cache: the five dense layers of the ray MLP as
// three-term split-fp16 tcgen05 GEMMs.  Reference: pose_estimation/ray_preprocessor.py:3-46,
// our_multihead_attention.py:75.
//
// Why.  The exact score mode (score_tc_mq.cu, f16x2) needs keys that agree with the reference's fp32 MLP to ~1e-6;
// the TF32 build (features_tc.cu) is only good to 1.5e-3 and the fp32 FMA build (features.cu) takes 2 s per
// 1M-Gaussian scene.  Here every activation and every weight is carried as an fp16 pair hi + lo (22 significant
// bits) and each layer computes  Ah.Wh + Ah.Wl + Al.Wh  in one fp32 TMEM accumulator (the dropped Al.Wl term is
// 2^-22 relative) -- fp32-grade results at tensor-core speed.
//
// Layout.  Every matrix is "x2": a row of width W is stored as [hi(W) | lo(W)] fp16 (the same layout as the
// SIXDGS_F16X2 key cache).  y[M, N] = act(x[M, K] w[N, K]^T + b[N]), K % 64 == 0, N % 128 == 0.  One persistent CTA
// per SM, tile = 128 rows x 128 columns; per 64-wide k-block the TMA brings four 128x64 fp16 boxes (Ah, Al, Wh, Wl:
// 64 KB, SWIZZLE_128B) through a 3-stage mbarrier ring and the MMA thread issues 3 x 4 tcgen05.mma kind::f16
// (M128 N128 K16); two TMEM accumulator buffers overlap the epilogue (bias, ReLU, hi/lo split, 16-byte row stores)
// with the next tile's MMAs.  Tiles are ordered n-fastest so the activation boxes are re-read from L2.
// Accumulation accuracy: the tensor core adds each MMA's result into the fp32 accumulator with truncation, a bias that
// grows with the number of MMAs chained on one accumulator (measured 6e-6 of max|k| after five layers with all three
// terms on one accumulator, against 7e-7 for the fp32 FMA build).  The two cross terms are 2^-11 of the main term, so
// they get their OWN accumulator (128 more TMEM columns per buffer: all 512 are used) and are added to the main one
// in fp32 registers by the epilogue: the main chain is K/16 MMAs instead of 3K/16.
// With both operands streaming a 128x128 tile moves 64 KB per 6.3 MFLOP (96 FLOP/B): the kernel is bound by the
// L2 -> SM path at roughly half the tensor peak, which still makes the build ~6x faster than the fp32 FMA path.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace sixdgs {

namespace {

constexpr int kX2Stages = 3;
constexpr int kX2BM = 128, kX2BN = 128, kX2BK = 64;   // 64 fp16 = one 128-byte swizzle row
constexpr int kX2Box = 128 * 128;                      // 16 KB per operand box
constexpr int kX2Threads = 256;
constexpr uint32_t kX2Idesc = umma_idesc(0 /*F16*/, kX2BM, kX2BN);

struct X2Smem {
  uint8_t ah[kX2Stages][kX2Box];
  uint8_t al[kX2Stages][kX2Box];
  uint8_t wh[kX2Stages][kX2Box];
  uint8_t wl[kX2Stages][kX2Box];
  uint64_t full[kX2Stages];
  uint64_t empty[kX2Stages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(X2Smem) + 1024 <= 232448, "x2 GEMM exceeds the shared-memory limit");

__device__ __forceinline__ void umma_f16_1sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kX2Idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// x2 row pointers: y_hi = y + row * ldy2 + col, y_lo = y_hi + y_lo_off  (halves)
__global__ void __launch_bounds__(kX2Threads, 1)
linear_x2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, int64_t m, int k,
                 int n, int x_lo_off, int w_lo_off, const float* __restrict__ bias, __half* __restrict__ y, int64_t ldy2,
                 int y_lo_off, int relu, float out_scale, unsigned int* __restrict__ absmax) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  X2Smem& sm = *reinterpret_cast<X2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles_n = n / kX2BN;
  const int64_t n_tiles = ((m + kX2BM - 1) / kX2BM) * n_tiles_n;
  const int nkb = k / kX2BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kX2Stages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full[a], 1); mbar_init(&sm.tmem_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = (int)(tile / n_tiles_n) * kX2BM;
        const int col0 = (int)(tile % n_tiles_n) * kX2BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&sm.full[stage], 4 * kX2Box);
          tma_load_2d(sm.ah[stage], &tmap_x, &sm.full[stage], kb * kX2BK, row0);
          tma_load_2d(sm.wh[stage], &tmap_w, &sm.full[stage], kb * kX2BK, col0);
          tma_load_2d(sm.wl[stage], &tmap_w, &sm.full[stage], w_lo_off + kb * kX2BK, col0);
          tma_load_2d(sm.al[stage], &tmap_x, &sm.full[stage], x_lo_off + kb * kX2BK, row0);
          if (++stage == kX2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int acc = (int)(it & 1);
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_main = tmem_base + (uint32_t)acc * (2 * kX2BN);  // hi . hi
        const uint32_t tmem_cross = tmem_main + kX2BN;                         // hi . lo + lo . hi
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t ah = smem_u32(sm.ah[stage]), al = smem_u32(sm.al[stage]);
          const uint32_t wh = smem_u32(sm.wh[stage]), wl = smem_u32(sm.wl[stage]);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_f16_1sm(tmem_main, umma_desc_sw128(ah + k4 * 32), umma_desc_sw128(wh + k4 * 32), (uint32_t)((kb | k4) != 0));
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_f16_1sm(tmem_cross, umma_desc_sw128(ah + k4 * 32), umma_desc_sw128(wl + k4 * 32), (uint32_t)((kb | k4) != 0));
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_f16_1sm(tmem_cross, umma_desc_sw128(al + k4 * 32), umma_desc_sw128(wh + k4 * 32), 1u);
          umma_commit_1sm(&sm.empty[stage]);
          if (++stage == kX2Stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_1sm(&sm.tmem_full[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    float amax = 0.f;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      const int64_t row = (tile / n_tiles_n) * kX2BM + row_in_tile;
      const int col0 = (int)(tile % n_tiles_n) * kX2BN;
      mbar_wait(&sm.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 2 * kX2BN);
      float cur[32], crs[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(taddr + c * 32, cur);
        tmem_ld32(taddr + kX2BN + c * 32, crs);
        tmem_ld_wait(cur);
        tmem_ld_wait(crs);
#pragma unroll
        for (int j = 0; j < 32; ++j) cur[j] += crs[j];  // fp32 round-to-nearest add of the two accumulators
        if (row < m) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + col0 + c * 32);
          __half* yh = y + row * ldy2 + col0 + c * 32;
          __half* yl = yh + y_lo_off;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 ba = __ldg(b4 + j / 4), bb = __ldg(b4 + j / 4 + 1);
            float o[8] = {cur[j] + ba.x, cur[j + 1] + ba.y, cur[j + 2] + ba.z, cur[j + 3] + ba.w,
                          cur[j + 4] + bb.x, cur[j + 5] + bb.y, cur[j + 6] + bb.z, cur[j + 7] + bb.w};
            __half h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = relu ? fmaxf(o[e], 0.f) : o[e];
              v *= out_scale;
              h[e] = __float2half_rn(v);
              l[e] = __float2half_rn(v - __half2float(h[e]));
              float av = fabsf(v);
              if (!(av == av)) av = INFINITY;
              amax = fmaxf(amax, av);
            }
            *reinterpret_cast<uint4*>(yh + j) = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
            *reinterpret_cast<uint4*>(yl + j) = make_uint4(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]), pack_h2(l[4], l[5]), pack_h2(l[6], l[7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
    }
    if (absmax) {
      amax = warp_max(amax);
      if (lane == 0 && amax > 0.f) atomicMax(absmax, __float_as_uint(amax));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

constexpr int kX2In = 141;      // MLP input width
constexpr int kX2InPad = 192;   // padded to k-blocks of 64
constexpr int kX2XW = 704;      // [h(512) | x(192)] concat width
constexpr int kX2Chunk = 1 << 17;

// MLP input (ray_preprocessor.py:3-9,36-44) as an x2 row slice: X[r, 512 + i] (hi) and X[r, 704 + 512 + i] (lo)
__global__ void pe_x2_kernel(const float* __restrict__ ori, const float* __restrict__ dir, const float* __restrict__ rgb,
                             int64_t n, __half* __restrict__ X) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  __half* xh = X + r * (2 * kX2XW) + 512;
  __half* xl = xh + kX2XW;
  float v[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) { v[i] = ori[r * 3 + i]; v[3 + i] = dir[r * 3 + i]; v[6 + i] = rgb[r * 3 + i]; }
  auto put = [&](int i, float f) {
    const __half h = __float2half_rn(f);
    xh[i] = h;
    xl[i] = __float2half_rn(f - __half2float(h));
  };
#pragma unroll
  for (int i = 0; i < 9; ++i) put(i, v[i]);
  int o = 9;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int nf = (g == 2) ? 6 : 8;
    for (int c = 0; c < 3; ++c)
      for (int f = 0; f < nf; ++f) {
        const float a = v[g * 3 + c] * (float)(1 << f);
        put(o + c * nf + f, sinf(a));
        put(o + 3 * nf + c * nf + f, cosf(a));
      }
    o += 6 * nf;
  }
  for (int i = kX2In; i < kX2InPad; ++i) { xh[i] = __float2half_rn(0.f); xl[i] = __float2half_rn(0.f); }
}

int launch_x2(const __half* x, int64_t m, int k, int64_t ldx2, int x_lo_off, int x_cols, const __half* w, int n, int w_lo_off,
              const float* b, __half* y, int64_t ldy2, int y_lo_off, int relu, float out_scale, unsigned int* absmax,
              cudaStream_t s) {
  CUtensorMap mx, mw;
  int rc;
  if ((rc = make_tmap_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x, (uint64_t)m, (uint64_t)x_cols, (uint64_t)ldx2 * 2, "linear_x2")))
    return rc;
  if ((rc = make_tmap_2d(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w, (uint64_t)n, (uint64_t)(2 * w_lo_off), (uint64_t)(2 * w_lo_off) * 2,
                         "linear_x2")))
    return rc;
  const size_t smem = sizeof(X2Smem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(linear_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("linear_x2 attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  const int64_t tiles = ((m + kX2BM - 1) / kX2BM) * (n / kX2BN);
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  linear_x2_kernel<<<grid, kX2Threads, smem, s>>>(mx, mw, m, k, n, x_lo_off, w_lo_off, b, y, ldy2, y_lo_off, relu, out_scale, absmax);
  return check_launch("linear_x2");
}

}  // namespace

}  // namespace sixdgs

using namespace sixdgs;

extern "C" size_t sixdgs_ray_features_x2_workspace(int64_t n) {
  const int64_t c = n < kX2Chunk ? (n > 0 ? n : 1) : kX2Chunk;
  return (size_t)c * 2 * (kX2XW + 512 + kFeat) * sizeof(__half) + 1024;
}

// Weights in x2 layout (built once by the host): w1 [512, 2*192] (mlp.0, 141 -> 192 zero padded), w2 [512, 2*512],
// w3 [512, 2*704] (mlp2.0: [h 512 | x 141 -> 192]), w4 [384, 2*512], wk [384, 2*384]; biases fp32.
extern "C" int sixdgs_ray_features_x2(const float* ori, const float* dir, const float* rgb, int64_t n, const void* w1,
                                      const float* b1, const void* w2, const float* b2, const void* w3, const float* b3,
                                      const void* w4, const float* b4, const void* wk, const float* bk, void* k_out,
                                      float* absmax, void* workspace, size_t workspace_bytes, void* stream) {
  SIXDGS_REQUIRE(ori && dir && rgb && w1 && b1 && w2 && b2 && w3 && b3 && w4 && b4 && wk && bk && k_out, "null pointer");
  SIXDGS_REQUIRE(n >= 0, "negative size");
  if (n == 0) return SIXDGS_OK;
  if (workspace == nullptr || workspace_bytes < sixdgs_ray_features_x2_workspace(n)) {
    set_error("ray_features_x2: workspace too small");
    return SIXDGS_EWORKSPACE;
  }
  SIXDGS_REQUIRE((reinterpret_cast<uintptr_t>(k_out) & 15) == 0, "k_out must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t cap = n < kX2Chunk ? n : kX2Chunk;
  __half* X = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  __half* H = X + cap * 2 * kX2XW;
  __half* F = H + cap * 2 * 512;
  const __half* W1 = (const __half*)w1; const __half* W2 = (const __half*)w2; const __half* W3 = (const __half*)w3;
  const __half* W4 = (const __half*)w4; const __half* WK = (const __half*)wk;
  for (int64_t r0 = 0; r0 < n; r0 += cap) {
    const int64_t c = (n - r0) < cap ? (n - r0) : cap;
    pe_x2_kernel<<<(unsigned)((c + 127) / 128), 128, 0, s>>>(ori + r0 * 3, dir + r0 * 3, rgb + r0 * 3, c, X);
    int rc = check_launch("pe_x2");
    if (rc) return rc;
    // mlp.0: x(192) -> H(512), relu.  The input slice starts at column 512 of X; its lo half sits 704 further on.
    if ((rc = launch_x2(X + 512, c, kX2InPad, 2 * kX2XW, kX2XW, 2 * kX2XW - 512, W1, 512, kX2InPad, b1, H, 2 * 512, 512, 1, 1.f, nullptr, s))) return rc;
    // mlp.2: H -> X[:, :512], relu
    if ((rc = launch_x2(H, c, 512, 2 * 512, 512, 2 * 512, W2, 512, 512, b2, X, 2 * kX2XW, kX2XW, 1, 1.f, nullptr, s))) return rc;
    // mlp2.0: [h, x](704) -> H(512), relu
    if ((rc = launch_x2(X, c, kX2XW, 2 * kX2XW, kX2XW, 2 * kX2XW, W3, 512, kX2XW, b3, H, 2 * 512, 512, 1, 1.f, nullptr, s))) return rc;
    // mlp2.2: H -> F(384)
    if ((rc = launch_x2(H, c, 512, 2 * 512, 512, 2 * 512, W4, kFeat, 512, b4, F, 2 * kFeat, kFeat, 0, 1.f, nullptr, s))) return rc;
    // k_proj: F -> keys (x16, the SIXDGS_F16X2 format)
    if ((rc = launch_x2(F, c, kFeat, 2 * kFeat, kFeat, 2 * kFeat, WK, kFeat, kFeat, bk, (__half*)k_out + r0 * 2 * kFeat, 2 * kFeat, kFeat, 0,
                        16.f, (unsigned int*)absmax, s)))
      return rc;
  }
  return SIXDGS_OK;
}
