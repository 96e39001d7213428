// EXPERIMENTAL (opt-in, impl = 2 of sixdgs_ray_features; not on any default path, validated in round 2):
// TF32 GEMM for the ray-feature MLP on CTA pairs with full-width output tiles.
//
// Why: the 1-CTA kernel in features_tc.cu moves 32 KB of operands per 524 k MACs (16 MAC/B); at the TF32 tensor
// peak that is 128 B/clk/SM of L2->SM traffic against the ~42 the chip's L2 delivers, and ncu shows the tensor pipe
// only 38-47 % active (profiles/prepare_kernels_r1.md).  Here a pair of CTAs owns a 256-row x N tile (N = 512 or
// 384 = the whole layer width): every k-block each CTA loads its own 128 x-rows (16 KB) once and HALF of the weight
// rows (N/2 x 128 B), and the pair issues two cta_group::2 UMMAs (M256 x N/2 x K8) that share the x tile ->
// 43.7 MAC/B for N = 512 (47 B/clk/SM at peak).  The accumulator is the full N columns of TMEM (single buffered:
// 2 x 512 does not fit), so the epilogue of a tile is not overlapped with the next tile's MMAs.
#include "tc_common.cuh"

namespace sixdgs {

constexpr int kL2Stages = 4;
constexpr int kL2ATile = 128 * 128;  // 128 rows x 32 fp32
constexpr int kL2BTile = 128 * 128;  // up to 128 rows x 32 fp32 per N-half per CTA
constexpr int kL2Threads = 256;

struct __align__(1024) L2Smem {
  uint8_t a[kL2Stages][kL2ATile];
  uint8_t b[kL2Stages][2][kL2BTile];
  uint64_t full[kL2Stages];
  uint64_t empty[kL2Stages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
};

template <int NH>
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(umma_idesc(2, 256, NH)), "r"(accumulate)
      : "memory");
}

template <typename TO>
__device__ __forceinline__ void store8_2(TO* p, const float* v);
template <>
__device__ __forceinline__ void store8_2<float>(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8_2<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  uint4 u;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[4], v[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[6], v[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(p) = u;
}

// N_TOTAL = 512 or 384; NH = N_TOTAL / 2 is the UMMA N of each of the two MMAs per k-step.
template <typename TO, int N_TOTAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kL2Threads, 1)
linear_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, int64_t m, int k,
                  const float* __restrict__ bias, TO* __restrict__ y, int64_t ldc, int relu, int round_tf32) {
  constexpr int NH = N_TOTAL / 2;
  constexpr int BROWS = NH / 2;  // weight rows each CTA loads per N-half (128 or 96)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  L2Smem& sm = *reinterpret_cast<L2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int64_t n_tiles = (m + 255) / 256;
  const int nkb = k / 32;
  constexpr uint32_t kStageBytes = 2u * (kL2ATile + 2u * BROWS * 128u);  // both CTAs

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kL2Stages; ++s) { mbar_init(&sm.full[s], 2); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.tmem_full, 1);
    mbar_init(&sm.tmem_empty, 8);  // 4 epilogue warps x 2 CTAs, on the leader
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        const int row0 = (int)(tile * 256 + rank * 128);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&sm.full[stage], kStageBytes);
          else mbar_arrive_cluster(&sm.full[stage], 0);
          tma_load_2sm(sm.a[stage], &tmap_x, &sm.full[stage], kb * 32, row0);
          // weight rows of N-half h that this CTA contributes: [h*NH + rank*BROWS, +BROWS)
          tma_load_2sm(sm.b[stage][0], &tmap_w, &sm.full[stage], kb * 32, (int)(0 * NH + rank * BROWS));
          tma_load_2sm(sm.b[stage][1], &tmap_w, &sm.full[stage], kb * 32, (int)(1 * NH + rank * BROWS));
          if (++stage == kL2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
        mbar_wait(&sm.tmem_empty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t aa = smem_u32(sm.a[stage]);
          const uint32_t b0 = smem_u32(sm.b[stage][0]), b1 = smem_u32(sm.b[stage][1]);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t da = umma_desc_sw128(aa + k4 * 32);
            umma_tf32_2sm<NH>(tmem_base, da, umma_desc_sw128(b0 + k4 * 32), (uint32_t)((kb | k4) != 0));
            umma_tf32_2sm<NH>(tmem_base + NH, da, umma_desc_sw128(b1 + k4 * 32), (uint32_t)((kb | k4) != 0));
          }
          umma_commit_2sm(&sm.empty[stage]);
          if (++stage == kL2Stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&sm.tmem_full);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int quad = warp & 3;
    int64_t it = 0;
    for (int64_t tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
      const int64_t row = tile * 256 + rank * 128 + quad * 32 + lane;
      mbar_wait(&sm.tmem_full, (uint32_t)(it & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
      float va[32], vb[32];
      tmem_ld32(taddr, va);
      tmem_ld_wait(va);
      constexpr int NCHUNK = N_TOTAL / 32;
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        float(&cur)[32] = (c & 1) ? vb : va;
        float(&nxt)[32] = (c & 1) ? va : vb;
        if (c + 1 < NCHUNK) tmem_ld32(taddr + (c + 1) * 32, nxt);
        if (row < m) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + c * 32);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 ba = __ldg(b4 + j / 4), bb = __ldg(b4 + j / 4 + 1);
            float o[8] = {cur[j] + ba.x, cur[j + 1] + ba.y, cur[j + 2] + ba.z, cur[j + 3] + ba.w,
                          cur[j + 4] + bb.x, cur[j + 5] + bb.y, cur[j + 6] + bb.z, cur[j + 7] + bb.w};
            if (relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
            }
            if (round_tf32) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(o[e]));
                o[e] = __uint_as_float(t);
              }
            }
            store8_2<TO>(y + row * ldc + c * 32 + j, o);
          }
        }
        if (c + 1 < NCHUNK) tmem_ld_wait(nxt);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&sm.tmem_empty, 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// tensor map with an explicit box height (weight tiles of 96 rows for the 384-wide layers)
static int make_tmap_rows(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                          uint32_t box_rows) {
  auto enc = tmap_encoder();
  if (!enc) { set_error("linear_tc2: cuTensorMapEncodeTiled unavailable"); return SIXDGS_EUNSUPPORTED; }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
  const cuuint32_t box[2] = {32, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("linear_tc2: cuTensorMapEncodeTiled failed (%d)", (int)r); return SIXDGS_ECUDA; }
  return SIXDGS_OK;
}

template <typename TO, int N_TOTAL>
static int launch_tc2_n(const float* x, int64_t m, int k, int64_t lda, const float* w, const float* b, TO* y, int64_t ldc,
                        int relu, int round_tf32, cudaStream_t s) {
  CUtensorMap mx, mw;
  int rc;
  if ((rc = make_tmap_rows(&mx, x, (uint64_t)m, (uint64_t)k, (uint64_t)lda * 4, 128))) return rc;
  if ((rc = make_tmap_rows(&mw, w, (uint64_t)N_TOTAL, (uint64_t)k, (uint64_t)k * 4, N_TOTAL / 4))) return rc;
  const size_t smem = sizeof(L2Smem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(linear_tc2_kernel<TO, N_TOTAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("linear_tc2 attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  linear_tc2_kernel<TO, N_TOTAL><<<(kNumSMs / 2) * 2, kL2Threads, smem, s>>>(mx, mw, m, k, b, y, ldc, relu, round_tf32);
  return check_launch("linear_tc2");
}

template <typename TO>
int launch_linear_tc2(const float* x, int64_t m, int k, int64_t lda, const float* w, const float* b, int n, TO* y,
                      int64_t ldc, int relu, int round_tf32, cudaStream_t s) {
  if (k % 32 != 0 || (n != 512 && n != 384) || (lda % 4) != 0 || (ldc % 8) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) {
    set_error("linear_tc2: k %% 32, n in {512, 384}, lda %% 4, ldc %% 8 and 16-byte alignment are required");
    return SIXDGS_EINVAL;
  }
  return n == 512 ? launch_tc2_n<TO, 512>(x, m, k, lda, w, b, y, ldc, relu, round_tf32, s)
                  : launch_tc2_n<TO, 384>(x, m, k, lda, w, b, y, ldc, relu, round_tf32, s);
}

template int launch_linear_tc2<float>(const float*, int64_t, int, int64_t, const float*, const float*, int, float*, int64_t,
                                      int, int, cudaStream_t);
template int launch_linear_tc2<__nv_bfloat16>(const float*, int64_t, int, int64_t, const float*, const float*, int,
                                              __nv_bfloat16*, int64_t, int, int, cudaStream_t);

}  // namespace sixdgs
