// a10 (+ k_proj), tensor-core build of the key cache: the five dense layers of the ray MLP as TF32
// tcgen05 GEMMs (fp32 operands read straight from shared memory, fp32 accumulation in TMEM).
// Reference: pose_estimation/ray_preprocessor.py:36-46, our_multihead_attention.py:75.
//
// y[M, N] = act(x[M, K] w[N, K]^T + b[N]) with x, w fp32 row-major (K-major for the MMA), K % 32 == 0,
// N % 128 == 0.  One persistent CTA per SM; tile = 128 rows x 128 columns:
//   warp 0      TMA producer: x tile [128 x 32 fp32] and w tile [128 x 32 fp32] per k-block (16 KB each,
//               SWIZZLE_128B) through a 6-stage mbarrier ring;
//   warp 1      MMA issuer: per k-block four tcgen05.mma.cta_group::1.kind::tf32 (M128 N128 K8);
//   warp 2      TMEM allocator (2 accumulator buffers x 128 columns);
//   warps 4..7  epilogue: tcgen05.ld -> +bias -> ReLU -> fp32 / bf16 stores (one output row per thread,
//               full 128-byte lines), overlapped with the next tile's MMAs.
// Tiles are ordered n-fastest so the x tile is re-read from L2, not HBM, by the N/128 CTAs that share it.
// Used for the bf16 (throughput) key cache; the exact (fp32-key) mode keeps the fp32 SIMT GEMMs.
// kind::tf32 reads fp32 words and TRUNCATES them to 10 mantissa bits, which biases every product towards
// zero (measured: 3e-3 relative on the keys after five layers).  Operands are therefore pre-rounded to
// TF32 with round-to-nearest where they are produced: the weights once at packing time, the activations
// in the epilogue that writes them (`round_tf32`), the MLP input in the PE kernel.  The truncation is
// then exact and the remaining error is unbiased (2^-12 per operand).
#include "tc_common.cuh"

namespace sixdgs {

constexpr int kLtStages = 6;
constexpr int kLtBM = 128, kLtBN = 128, kLtBK = 32;  // 32 fp32 = 128 B swizzle row
constexpr int kLtTileBytes = 128 * 128;               // 16 KB per operand per stage
constexpr int kLtThreads = 256;
constexpr uint32_t kLtIdesc = umma_idesc(2 /*TF32*/, kLtBM, kLtBN);

struct __align__(1024) LtSmem {
  uint8_t a[kLtStages][kLtTileBytes];
  uint8_t b[kLtStages][kLtTileBytes];
  uint64_t full[kLtStages];
  uint64_t empty[kLtStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void umma_tf32_1sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kLtIdesc), "r"(accumulate)
      : "memory");
}

template <typename TO>
__device__ __forceinline__ void store8(TO* p, const float* v);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  uint4 u;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[4], v[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[6], v[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename TO>
__global__ void __launch_bounds__(kLtThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, int64_t m, int k,
                 int n, const float* __restrict__ bias, TO* __restrict__ y, int64_t ldc, int relu, int round_tf32) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  LtSmem& sm = *reinterpret_cast<LtSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles_n = n / kLtBN;
  const int64_t n_tiles = ((m + kLtBM - 1) / kLtBM) * n_tiles_n;
  const int nkb = k / kLtBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kLtStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full[a], 1); mbar_init(&sm.tmem_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(256)
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
        const int row0 = (int)(tile / n_tiles_n) * kLtBM;
        const int col0 = (int)(tile % n_tiles_n) * kLtBN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&sm.full[stage], 2 * kLtTileBytes);
          tma_load_2d(sm.a[stage], &tmap_x, &sm.full[stage], kb * kLtBK, row0);
          tma_load_2d(sm.b[stage], &tmap_w, &sm.full[stage], kb * kLtBK, col0);
          if (++stage == kLtStages) { stage = 0; phase ^= 1; }
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
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * kLtBN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t aa = smem_u32(sm.a[stage]), ba = smem_u32(sm.b[stage]);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)  // 4 x K8 (32 bytes) per 128-byte swizzle row
            umma_tf32_1sm(tmem_d, umma_desc_sw128(aa + k4 * 32), umma_desc_sw128(ba + k4 * 32), (uint32_t)((kb | k4) != 0));
          umma_commit_1sm(&sm.empty[stage]);
          if (++stage == kLtStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_1sm(&sm.tmem_full[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      const int64_t row = (tile / n_tiles_n) * kLtBM + row_in_tile;
      const int col0 = (int)(tile % n_tiles_n) * kLtBN;
      mbar_wait(&sm.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kLtBN);
      float va[32], vb[32];
      tmem_ld32(taddr, va);
      tmem_ld_wait(va);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float(&cur)[32] = (c & 1) ? vb : va;
        float(&nxt)[32] = (c & 1) ? va : vb;
        if (c + 1 < 4) tmem_ld32(taddr + (c + 1) * 32, nxt);
        if (row < m) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + col0 + c * 32);
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
            store8<TO>(y + row * ldc + col0 + c * 32 + j, o);
          }
        }
        if (c + 1 < 4) tmem_ld_wait(nxt);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

constexpr int kLtStagedStages = 5;
struct __align__(1024) LtStagedSmem {
  uint8_t a[kLtStagedStages][kLtTileBytes];
  uint8_t b[kLtStagedStages][kLtTileBytes];
  uint8_t stage_out[2][kLtTileBytes];  // two 128-row x 128-byte boxes for the TMA tensor stores
  uint64_t full[kLtStagedStages];
  uint64_t empty[kLtStagedStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// impl = 1 (default since round 2; impl = 3 is the direct-store kernel above, kept as the bit-identity reference):
// same GEMM with a shared-memory staged, TMA-store epilogue.
template <typename TO>
__global__ void __launch_bounds__(kLtThreads, 1)
linear_tc_staged_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_y, int64_t m, int k,
                 int n, const float* __restrict__ bias, TO* __restrict__ y, int64_t ldc, int relu, int round_tf32) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  LtStagedSmem& sm = *reinterpret_cast<LtStagedSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles_n = n / kLtBN;
  const int64_t n_tiles = ((m + kLtBM - 1) / kLtBM) * n_tiles_n;
  const int nkb = k / kLtBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kLtStagedStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full[a], 1); mbar_init(&sm.tmem_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(256)
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
        const int row0 = (int)(tile / n_tiles_n) * kLtBM;
        const int col0 = (int)(tile % n_tiles_n) * kLtBN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&sm.full[stage], 2 * kLtTileBytes);
          tma_load_2d(sm.a[stage], &tmap_x, &sm.full[stage], kb * kLtBK, row0);
          tma_load_2d(sm.b[stage], &tmap_w, &sm.full[stage], kb * kLtBK, col0);
          if (++stage == kLtStagedStages) { stage = 0; phase ^= 1; }
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
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * kLtBN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t aa = smem_u32(sm.a[stage]), ba = smem_u32(sm.b[stage]);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)  // 4 x K8 (32 bytes) per 128-byte swizzle row
            umma_tf32_1sm(tmem_d, umma_desc_sw128(aa + k4 * 32), umma_desc_sw128(ba + k4 * 32), (uint32_t)((kb | k4) != 0));
          umma_commit_1sm(&sm.empty[stage]);
          if (++stage == kLtStagedStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_1sm(&sm.tmem_full[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // Staged epilogue: TMEM -> registers (+bias, ReLU, TF32 rounding) -> shared memory in the TMA 128-byte swizzle
    // -> one TMA tensor store per 128-row x 128-byte box.  (The direct per-thread row stores of linear_tc_kernel
    // issue 32 separate cache lines per warp instruction.)  Two staging boxes alternate; thread `et == 0` owns
    // the bulk-group bookkeeping.
    const int quad = warp & 3;
    const int et = tid - 128;                 // 0..127 within the epilogue
    const int r = quad * 32 + lane;           // row inside the tile == TMEM lane
    constexpr int kColsPerBox = 128 / (int)sizeof(TO);  // 32 fp32 or 64 bf16 columns per 128-byte box row
    constexpr int kChunksPerBox = kColsPerBox / 32;
    uint32_t box_counter = 0;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      const int row0 = (int)(tile / n_tiles_n) * kLtBM;
      const int col0 = (int)(tile % n_tiles_n) * kLtBN;
      mbar_wait(&sm.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kLtBN);
      float va[32], vb[32];
      tmem_ld32(taddr, va);
      tmem_ld_wait(va);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float(&cur)[32] = (c & 1) ? vb : va;
        float(&nxt)[32] = (c & 1) ? va : vb;
        if (c + 1 < 4) tmem_ld32(taddr + (c + 1) * 32, nxt);
        const int sub = c % kChunksPerBox;    // position of this 32-column chunk inside its box
        uint8_t* box = sm.stage_out[box_counter & 1];
        if (sub == 0) {
          // the box buffer is reused every second box: its previous TMA store must have finished reading it
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        const float4* b4 = reinterpret_cast<const float4*>(bias + col0 + c * 32);
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
          // 16-byte unit index inside the 128-byte row, XOR-swizzled with (row & 7) like CU_TENSOR_MAP_SWIZZLE_128B
          if (sizeof(TO) == 4) {
            const int u0 = (j / 4), u1 = (j / 4 + 1);
            *reinterpret_cast<float4*>(box + r * 128 + ((u0 ^ (r & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(box + r * 128 + ((u1 ^ (r & 7)) << 4)) = make_float4(o[4], o[5], o[6], o[7]);
          } else {
            const int u = sub * 4 + j / 8;
            uint4 pk;
            __nv_bfloat162 t2;
            t2 = __floats2bfloat162_rn(o[0], o[1]); pk.x = *reinterpret_cast<uint32_t*>(&t2);
            t2 = __floats2bfloat162_rn(o[2], o[3]); pk.y = *reinterpret_cast<uint32_t*>(&t2);
            t2 = __floats2bfloat162_rn(o[4], o[5]); pk.z = *reinterpret_cast<uint32_t*>(&t2);
            t2 = __floats2bfloat162_rn(o[6], o[7]); pk.w = *reinterpret_cast<uint32_t*>(&t2);
            *reinterpret_cast<uint4*>(box + r * 128 + ((u ^ (r & 7)) << 4)) = pk;
          }
        }
        if (sub == kChunksPerBox - 1) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (et == 0) {
            const int box_col = col0 + (c / kChunksPerBox) * kColsPerBox;
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmap_y)),
                         "r"(smem_u32(box)), "r"(box_col), "r"(row0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++box_counter;
        }
        if (c + 1 < 4) tmem_ld_wait(nxt);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// y = act(x w^T + b) on the tensor cores.  x [m, k] fp32 with row stride lda (elements), w [n, k] fp32 dense.
template <typename TO>
int launch_linear_tc(const float* x, int64_t m, int k, int64_t lda, const float* w, const float* b, int n, TO* y,
                     int64_t ldc, int relu, int round_tf32, cudaStream_t s) {
  if (k % kLtBK != 0 || n % kLtBN != 0 || (lda % 4) != 0 || (ldc % 8) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) {
    set_error("linear_tc: k %% 32, n %% 128, lda %% 4, ldc %% 8 and 16-byte alignment are required");
    return SIXDGS_EINVAL;
  }
  CUtensorMap mx, mw;
  int rc;
  if ((rc = make_tmap_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, (uint64_t)m, (uint64_t)k, (uint64_t)lda * 4, "linear_tc")))
    return rc;
  if ((rc = make_tmap_2d(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, w, (uint64_t)n, (uint64_t)k, (uint64_t)k * 4, "linear_tc")))
    return rc;
  const size_t smem = sizeof(LtSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("linear_tc attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  const int64_t tiles = ((m + kLtBM - 1) / kLtBM) * (n / kLtBN);
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  linear_tc_kernel<TO><<<grid, kLtThreads, smem, s>>>(mx, mw, m, k, n, b, y, ldc, relu, round_tf32);
  return check_launch("linear_tc");
}

template <typename TO>
int launch_linear_tc_staged(const float* x, int64_t m, int k, int64_t lda, const float* w, const float* b, int n, TO* y,
                            int64_t ldc, int relu, int round_tf32, cudaStream_t s) {
  if (k % kLtBK != 0 || n % kLtBN != 0 || (lda % 4) != 0 || ((ldc * sizeof(TO)) % 16) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) {
    set_error("linear_tc_staged: k %% 32, n %% 128 and 16-byte aligned rows are required");
    return SIXDGS_EINVAL;
  }
  CUtensorMap mx, mw, my;
  int rc;
  if ((rc = make_tmap_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, (uint64_t)m, (uint64_t)k, (uint64_t)lda * 4, "linear_tc_staged")))
    return rc;
  if ((rc = make_tmap_2d(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, w, (uint64_t)n, (uint64_t)k, (uint64_t)k * 4, "linear_tc_staged")))
    return rc;
  if ((rc = make_tmap_2d(&my, sizeof(TO) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (int)sizeof(TO), y,
                         (uint64_t)m, (uint64_t)n, (uint64_t)ldc * sizeof(TO), "linear_tc_staged")))
    return rc;
  const size_t smem = sizeof(LtStagedSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(linear_tc_staged_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("linear_tc_staged attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  const int64_t tiles = ((m + kLtBM - 1) / kLtBM) * (n / kLtBN);
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  linear_tc_staged_kernel<TO><<<grid, kLtThreads, smem, s>>>(mx, mw, my, m, k, n, b, y, ldc, relu, round_tf32);
  return check_launch("linear_tc_staged");
}

template int launch_linear_tc_staged<float>(const float*, int64_t, int, int64_t, const float*, const float*, int, float*,
                                            int64_t, int, int, cudaStream_t);
template int launch_linear_tc_staged<__nv_bfloat16>(const float*, int64_t, int, int64_t, const float*, const float*, int,
                                                    __nv_bfloat16*, int64_t, int, int, cudaStream_t);

template int launch_linear_tc<float>(const float*, int64_t, int, int64_t, const float*, const float*, int, float*, int64_t,
                                     int, int, cudaStream_t);
template int launch_linear_tc<__nv_bfloat16>(const float*, int64_t, int, int64_t, const float*, const float*, int,
                                             __nv_bfloat16*, int64_t, int, int, cudaStream_t);

}  // namespace sixdgs
