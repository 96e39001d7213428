// a11, several queries per key sweep (SURVEY §8d "ray-score, batch" row), in two key formats:
//   bf16   (FMT 0): one MMA term, 768 B/ray -- the throughput mode (3e-2 score tolerance);
//   f16f8  (FMT 2): the main term Qh.Kh in fp16 and the two 2^-11 cross terms in e4m3 at twice the tensor rate (two
//          term-units instead of three); the cross terms are then good to ~5 %, i.e. the logits to ~0.05 x the fp16
//          rounding error: inside the 1e-3 bar for logit standard deviations up to ~12, without the margin of f16x2;
//   f16x2  (FMT 1): every key and every query element is carried as an fp16 pair hi + lo (22 significant
//          bits) and the logit is the three-term product  Qh.Kh + Qh.Kl + Ql.Kh  accumulated in one fp32 TMEM
//          accumulator (the dropped Ql.Kl term is 2^-22 relative) -- the EXACT tensor-core mode: scores agree with the
//          fp32 reference to ~1e-5 even for a peaked (trained) softmax where bf16 logits are off by 1e-1.
// Validated on B200 in round 2 (bit-identical to score_tc.cu in bf16 mode; profiles/validate_experimental_r2.log).
// Reference: pose_estimation/our_multihead_attention.py:4-12,70-79; identification_module.py:80-82;
// fused all-ray weighted least squares: least_squared_loss.py:47-64, line_intersection.py:75-154.
//
// Why.  score_tc.cu streams the whole bf16 key cache from HBM twice per query.  At 256 FLOP/B that kernel sits on
// the B200 ridge and, run back to back, is limited by board power -- and roughly a quarter of that
// energy is DRAM access.  Here one sweep over the keys serves up to kMqMaxQ queries: every key byte crosses
// HBM -> SM once per pass per BATCH; the tensor work, the MUFU work and the SM ingest are unchanged (the other
// operand, the image tokens of the batch, 192 KB per query, streams from L2 instead).
//
// Roles are those of score_tc.cu with the operands' lifetimes swapped (persistent, CTA pairs, cta_group::2):
//   * the CTA's 128 rays of the current pair-tile stay resident in shared memory (six 128x64 bf16 k-blocks, 96 KB)
//     while the batch's queries are run against them; the slices are refilled one by one as the LAST query of the
//     tile retires them, so the next tile's keys arrive under the tail of this one;
//   * each CTA's half (128 tokens) of every query streams through a 6-stage ring of the same 16 KB k-blocks
//     (from L2; its own producer warp, so it runs ahead of the key refills);
//   * accumulators: TMEM, 128 lanes x 256 fp32 columns per (tile, query), double buffered;
//   * pass 1: D = Q K^T (lanes = tokens), per-(query, token) running (max, sum-exp2) kept in shared memory slots
//     private to each epilogue thread; pass 2: D = K Q^T (lanes = rays), c_token per query in shared memory.
//   * warp roles: 0 = key producer, 3 = token producer, 1 = MMA issuer (leader CTA), 2 = TMEM allocator,
//     4..11 = epilogue.
// f16x2 layout: a cache row is [hi(384) | lo(384)] fp16 (1536 B) holding 16*k; the query workspace rows are laid out
// the same way and hold 64*log2(e)/sqrt(384)*q, so the accumulator is 2^10 * logit*log2(e) and the epilogue multiplies
// by 2^-10 inside the FMA that subtracts the softmax offset (power-of-two scales: exact; they keep the lo halves out
// of the fp16 subnormal range).  The resident slices are the hi halves (96 KB); the ring (7 stages) carries, per
// k-block of a query, the query's hi block, the tile's lo key block (HBM once, then L2 for the other queries of the
// batch) and the query's lo block -- 288 KB per (tile, query) against three times the MMA work of the bf16 mode, i.e.
// the same L2 -> SM rate.
// Optional pass-2 epilogue: the all-ray weighted least-squares system of least_squared_loss.py:62-64
// (weights = score / n_img) -- per-CTA fp64 partial sums of R (6), q (3), sum w d (3), sum w (1) from the scores the
// epilogue already holds (+24 B/ray for origin and direction), reduced by sixdgs_ls_reduce / one 13-double allreduce.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "tc_common.cuh"

namespace sixdgs {

namespace {

constexpr int kMqMaxQ = 8;
constexpr int kMqLs = 13;  // R (xx,xy,xz,yy,yz,zz), q (3), sum w d (3), sum w
constexpr int kMqTileRays = 256;
constexpr int kMqKBlocks = kFeat / 64;  // 6
constexpr int kMqKBBytes = 128 * 128;   // 128 rows x 128 B
constexpr int kMqThreads = 384;
constexpr int kMqPairs = kNumSMs / 2;
constexpr float kMqLog2e = 1.4426950408889634f;
constexpr float kMqLn2 = 0.6931471805599453f;
constexpr float kMqQScale = 1.4426950408889634f / 19.595917942265423f;  // log2(e) / sqrt(384)

constexpr int kFmtBf16 = 0, kFmtF16x2 = 1, kFmtF16F8 = 2;
template <int FMT>
struct MqCfg {
  static constexpr int kStages = FMT == kFmtF16x2 ? 7 : (FMT == kFmtF16F8 ? 4 : 6);
  static constexpr int kResident = FMT == kFmtF16F8 ? 9 : 6;   // resident 16 KB key boxes per CTA (f16f8: 6 x Kh + 3 x Kh8)
  static constexpr int kRowBytes = FMT == kFmtBf16 ? 768 : 1536;   // bytes per cache / query row
  // bf16 (1) or fp16 (0) -> fp32, M = N = 256; kind::f8f6f4 uses the same field with E4M3 = 0
  static constexpr uint32_t kIdesc = umma_idesc(FMT == kFmtBf16 ? 1u : 0u, 256, 256);
};
constexpr float kMqKeyScale = 16.0f;                 // f16x2: stored key = 16 k
constexpr float kMqQryScale = 64.0f;                 // f16x2: stored query = 64 log2(e)/sqrt(384) q
constexpr float kMqSplitOut = 1.0f / 1024.0f;        // accumulator -> log2-domain logit
// f16f8 row: [hi fp16 (768 B) | hi8 e4m3 (384 B) | lo8 e4m3 (384 B)].  The e4m3 copies carry power-of-two scales chosen so that
// the cross products land on the scale of the main term (one shared accumulator) and inside e4m3's normal range:
//   keys:    hi8 = e4m3(hi / 64)  (= k / 4),            lo8 = e4m3(64 lo)
//   queries: hi8 = e4m3(hi / 64)  (= log2e/sqrt(384) q), lo8 = e4m3(64 lo)        -> Qh8.Kl8 = Qh.Kl,  Ql8.Kh8 = Ql.Kh
constexpr float kMqF8Down = 1.0f / 64.0f, kMqF8Up = 64.0f;

// (no struct-level alignment attribute: the kernel aligns the base to 1024 B by hand and sizeof must not be padded --
// the f16x2 variant uses all but 768 B of the 227 KB a CTA can have)
template <int STAGES, int RES>
struct MqSmem {
  uint8_t kt[RES][kMqKBBytes];         //  96 / 144 KB: this CTA's 128 rays of the current tile (hi halves; f16f8: + e4m3 copy)
  uint8_t qs[STAGES][kMqKBBytes];      //  96 / 112 / 64 KB: token ring (f16x2: query hi, key lo, query lo blocks)
  uint64_t k_full[RES];
  uint64_t k_empty[RES];
  uint64_t q_full[STAGES];
  uint64_t q_empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  union alignas(16) {
    struct {
      float cst[kMqMaxQ][kMaxTokens];  // pass 2: c_token = m*log2e + log2 z per query (+inf: token unused)
      double ls[kMqMaxQ][4][kMqLs];    // pass 2, fused LS: per (query, half-0 epilogue warp) partial sums
    };
    float2 stat[kMqMaxQ][2][128];      // pass 1: (running max, running sum) per query / column half / token row
  };
  float xch[2][128];                   // column-half exchange (pass 1's final merge borrows ring stage 0 for a second one)
};
static_assert(sizeof(MqSmem<7, 6>) + 1024 <= 232448, "f16x2 variant exceeds the 227 KB shared-memory limit");
static_assert(sizeof(MqSmem<4, 9>) + 1024 <= 232448, "f16f8 variant exceeds the 227 KB shared-memory limit");

__device__ __forceinline__ void mq_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mq_umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float mq_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void mq_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// fold 32 logits (one TMEM chunk) of one token into its running (max, sum-exp2); nv = real rays in the chunk
// (sc = accumulator -> log2-logit scale; 1 in bf16 mode, where fmaf(x, 1, -m) == x - m bit for bit)
__device__ __forceinline__ void mq_fold(const float (&cur)[32], int nv, float& run_m, float& run_z, const float sc) {
  if (nv >= 32) {
    float m0 = fmaxf(cur[0], cur[1]), m1 = fmaxf(cur[2], cur[3]), m2 = fmaxf(cur[4], cur[5]), m3 = fmaxf(cur[6], cur[7]);
#pragma unroll
    for (int j = 8; j < 32; j += 8) {
      m0 = fmaxf(m0, fmaxf(cur[j + 0], cur[j + 1]));
      m1 = fmaxf(m1, fmaxf(cur[j + 2], cur[j + 3]));
      m2 = fmaxf(m2, fmaxf(cur[j + 4], cur[j + 5]));
      m3 = fmaxf(m3, fmaxf(cur[j + 6], cur[j + 7]));
    }
    const float mn = fmaxf(run_m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sc);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      s0 += mq_ex2(fmaf(cur[j + 0], sc, -mn));
      s1 += mq_ex2(fmaf(cur[j + 1], sc, -mn));
      s2 += mq_ex2(fmaf(cur[j + 2], sc, -mn));
      s3 += mq_ex2(fmaf(cur[j + 3], sc, -mn));
    }
    run_z = run_z * mq_ex2(run_m - mn) + ((s0 + s1) + (s2 + s3));
    run_m = mn;
  } else if (nv > 0) {
    float cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) cm = fmaxf(cm, (j < nv) ? cur[j] : -INFINITY);
    const float mn = fmaxf(run_m, cm * sc);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += (j < nv) ? mq_ex2(fmaf(cur[j], sc, -mn)) : 0.f;
    run_z = run_z * mq_ex2(run_m - mn) + s;
    run_m = mn;
  }
}

// Qb[b, t, :] = bf16(q[b, t, :] * log2(e)/sqrt(384)) for t < n_img, 0 otherwise
__global__ void mq_qprep_kernel(const float* __restrict__ q, int n_img, int nq, __nv_bfloat16* __restrict__ qb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nq * kMaxTokens * kFeat) return;
  const int t = (int)((i / kFeat) % kMaxTokens);
  qb[i] = __float2bfloat16_rn(t < n_img ? q[i] * kMqQScale : 0.0f);
}
// f16x2: row [hi(384) | lo(384)] of 64 * log2(e)/sqrt(384) * q   (hi = fp16(x), lo = fp16(x - hi))
__global__ void mq_qprep_split_kernel(const float* __restrict__ q, int n_img, int nq, __half* __restrict__ qb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nq * kMaxTokens * kFeat) return;
  const int64_t row = i / kFeat;
  const int col = (int)(i % kFeat);
  const int t = (int)(row % kMaxTokens);
  const float x = t < n_img ? q[i] * (kMqQScale * kMqQryScale) : 0.0f;
  const __half hi = __float2half_rn(x);
  qb[row * (2 * kFeat) + col] = hi;
  qb[row * (2 * kFeat) + kFeat + col] = __float2half_rn(x - __half2float(hi));
}
// fp32 keys [n,384] -> f16x2 rows [n, 768] = [hi | lo] of 16 k; absmax (nullable) receives max |16 k| (as float bits)
__global__ void mq_split_keys_kernel(const float* __restrict__ k, int64_t n, __half* __restrict__ out,
                                     unsigned int* __restrict__ absmax) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float a = 0.f;
  if (i < n * kFeat) {
    const int64_t row = i / kFeat;
    const int col = (int)(i % kFeat);
    const float x = k[i] * kMqKeyScale;
    const __half hi = __float2half_rn(x);
    out[row * (2 * kFeat) + col] = hi;
    out[row * (2 * kFeat) + kFeat + col] = __float2half_rn(x - __half2float(hi));
    a = fabsf(x);
    if (!(a == a)) a = INFINITY;  // NaN keys must trip the range check too
  }
  if (absmax) {
    a = warp_max(a);
    if ((threadIdx.x & 31) == 0 && a > 0.f) atomicMax(absmax, __float_as_uint(a));  // non-negative floats order as uints
  }
}

__device__ __forceinline__ uint8_t mq_e4m3(float x) {
  return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3);
}
// f16f8 query rows: [hi fp16 | e4m3(hi / 64) | e4m3(64 lo)] of 64 * log2(e)/sqrt(384) * q
__global__ void mq_qprep_f8_kernel(const float* __restrict__ q, int n_img, int nq, uint8_t* __restrict__ qb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nq * kMaxTokens * kFeat) return;
  const int64_t row = i / kFeat;
  const int col = (int)(i % kFeat);
  const int t = (int)(row % kMaxTokens);
  const float x = t < n_img ? q[i] * (kMqQScale * kMqQryScale) : 0.0f;
  const __half hi = __float2half_rn(x);
  const float hf = __half2float(hi);
  uint8_t* r = qb + row * 1536;
  reinterpret_cast<__half*>(r)[col] = hi;
  r[768 + col] = mq_e4m3(hf * kMqF8Down);
  r[1152 + col] = mq_e4m3((x - hf) * kMqF8Up);
}
// f16x2 key rows [hi | lo] (fp16) -> f16f8 rows [hi | e4m3(hi / 64) | e4m3(64 lo)], IN PLACE (same 1536 B): one warp
// per row reads the whole row into registers before it overwrites the lo half
__global__ void __launch_bounds__(256) mq_keys_to_f8_kernel(uint8_t* __restrict__ keys, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  uint8_t* r = keys + row * 1536;
  const __half* h = reinterpret_cast<const __half*>(r);
  float hi[12], lo[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) { hi[j] = __half2float(h[lane + 32 * j]); lo[j] = __half2float(h[kFeat + lane + 32 * j]); }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    r[768 + lane + 32 * j] = mq_e4m3(hi[j] * kMqF8Down);
    r[1152 + lane + 32 * j] = mq_e4m3(lo[j] * kMqF8Up);
  }
}

template <int PASS, int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMqThreads, 1)
score_tc_mq_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_q,
                   int64_t n_rays, int n_img, int nq,
                   float* __restrict__ part_m, float* __restrict__ part_z,      // pass 1 out [nq, pairs, 256]
                   const float* __restrict__ gm, const float* __restrict__ gz,  // pass 2 in  [nq, 256]
                   float* __restrict__ scores, int64_t score_stride,            // pass 2 out [nq, score_stride]
                   const float* __restrict__ rays_ori, const float* __restrict__ rays_dir,  // pass 2, fused LS (nullable)
                   double* __restrict__ ls_part) {                              // pass 2 out [nq, 2*pairs, 13]
  using Cfg = MqCfg<FMT>;
  constexpr bool SPLIT = FMT == kFmtF16x2;
  constexpr bool F8 = FMT == kFmtF16F8;
  constexpr int kStages = Cfg::kStages;
  constexpr int kRes = Cfg::kResident;
  constexpr float kSc = FMT == kFmtBf16 ? 1.0f : kMqSplitOut;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using Smem = MqSmem<kStages, kRes>;
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int64_t n_tiles = (n_rays + kMqTileRays - 1) / kMqTileRays;
  const bool fuse_ls = PASS == 2 && ls_part != nullptr;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRes; ++s) {
      mbar_init(&sm.k_full[s], 2);   // leader's expect_tx arrive + peer's remote arrive
      mbar_init(&sm.k_empty[s], 1);  // one multicast tcgen05.commit (after the tile's last query)
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.q_full[s], 2);
      mbar_init(&sm.q_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sm.tmem_full[a], 1);
      mbar_init(&sm.tmem_empty[a], 16);  // 8 epilogue warps x 2 CTAs, on the leader's copy
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (PASS == 2) {
    for (int i = tid; i < nq * kMaxTokens; i += kMqThreads) {
      const int t = i % kMaxTokens;
      (&sm.cst[0][0])[i] = (t < n_img) ? (gm[i] * kMqLog2e + log2f(gz[i])) : INFINITY;
    }
    for (int i = tid; i < kMqMaxQ * 4 * kMqLs; i += kMqThreads) (&sm.ls[0][0][0])[i] = 0.0;
  } else {
    for (int i = tid; i < nq * 2 * 128; i += kMqThreads) (&sm.stat[0][0][0])[i] = make_float2(-INFINITY, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ================================ key producer (both CTAs): the resident (hi) slices ================================
    if (lane == 0) {
      uint32_t phase = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        const int row0 = (int)(tile * kMqTileRays + rank * 128);
        for (int kb = 0; kb < kRes; ++kb) {
          mbar_wait(&sm.k_empty[kb], phase ^ 1);  // the previous tile's last query is done with this slice
          if (leader) mbar_arrive_expect_tx(&sm.k_full[kb], 2 * kMqKBBytes);
          else mbar_arrive_cluster(&sm.k_full[kb], 0);
          // f16f8 maps are byte-typed: hi blocks at 128 kb, the e4m3 copy of hi at 768 + 128 j
          const int col = F8 ? (kb < 6 ? kb * 128 : 768 + (kb - 6) * 128) : kb * 64;
          tma_load_2sm(sm.kt[kb], &tmap_k, &sm.k_full[kb], col, row0);
        }
        phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ================================ ring producer (both CTAs) ================================
    // bf16: the query's k-block.  f16x2: query hi block, key lo block of the tile, query lo block.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto push = [&](const CUtensorMap* map, int c0, int c1) {
        mbar_wait(&sm.q_empty[stage], phase ^ 1);
        if (leader) mbar_arrive_expect_tx(&sm.q_full[stage], 2 * kMqKBBytes);
        else mbar_arrive_cluster(&sm.q_full[stage], 0);
        tma_load_2sm(sm.qs[stage], map, &sm.q_full[stage], c0, c1);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      };
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        const int krow0 = (int)(tile * kMqTileRays + rank * 128);
        for (int b = 0; b < nq; ++b) {
          const int row0 = b * kMaxTokens + (int)rank * 128;
          if (F8) {
            for (int j = 0; j < 3; ++j) {
              push(&tmap_q, (2 * j) * 128, row0);        // Qh block 2j     (fp16, 64 elements)
              push(&tmap_q, (2 * j + 1) * 128, row0);    // Qh block 2j + 1
              push(&tmap_q, 768 + j * 128, row0);        // Qh8 block j     (e4m3, 128 elements)
              push(&tmap_k, 1152 + j * 128, krow0);      // Kl8 block j
              push(&tmap_q, 1152 + j * 128, row0);       // Ql8 block j
            }
          } else {
            for (int kb = 0; kb < kMqKBlocks; ++kb) {
              push(&tmap_q, kb * 64, row0);
              if (SPLIT) {
                push(&tmap_k, kFeat + kb * 64, krow0);
                push(&tmap_q, kFeat + kb * 64, row0);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA, one thread) ================================
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t qphase = 0, kphase = 0;
      int64_t it = 0;  // (tile, query) counter
      auto next = [&]() { if (++stage == kStages) { stage = 0; qphase ^= 1; } };
      // four K=16 MMAs over one 64-wide k-block; `tok` / `key` are the shared-memory blocks of the two operands
      auto group = [&](uint32_t tmem_d, uint32_t tok, uint32_t key, bool first) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint64_t dq = umma_desc_sw128(tok + k4 * 32);
          const uint64_t dk = umma_desc_sw128(key + k4 * 32);
          const uint32_t accum = (uint32_t)(!(first && k4 == 0));
          if (PASS == 1) mq_umma(tmem_d, dq, dk, Cfg::kIdesc, accum);  // D[token, ray]
          else mq_umma(tmem_d, dk, dq, Cfg::kIdesc, accum);            // D[ray, token]
        }
      };
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        for (int b = 0; b < nq; ++b, ++it) {
          const int acc = (int)(it & 1);
          const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
          mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
          if (F8) {
            // main term in fp16 (K = 16 per MMA, 64-element blocks), cross terms in e4m3 (K = 32 per MMA, 128-element
            // blocks): per j two Qh.Kh groups, then Qh8.Kl8 and Ql8.Kh8
            auto group8 = [&](uint32_t tok, uint32_t key) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t dq = umma_desc_sw128(tok + k4 * 32);
                const uint64_t dk = umma_desc_sw128(key + k4 * 32);
                if (PASS == 1) mq_umma_f8(tmem_d, dq, dk, Cfg::kIdesc, 1u);
                else mq_umma_f8(tmem_d, dk, dq, Cfg::kIdesc, 1u);
              }
            };
            for (int j = 0; j < 3; ++j) {
              for (int h = 0; h < 2; ++h) {
                const int kb = 2 * j + h;
                if (b == 0) mbar_wait(&sm.k_full[kb], kphase);
                mbar_wait(&sm.q_full[stage], qphase);
                tc_fence_after();
                group(tmem_d, smem_u32(sm.qs[stage]), smem_u32(sm.kt[kb]), kb == 0);   // Qh . Kh
                umma_commit_2sm(&sm.q_empty[stage]);
                if (b == nq - 1) umma_commit_2sm(&sm.k_empty[kb]);
                next();
              }
              mbar_wait(&sm.q_full[stage], qphase);                                     // Qh8
              const int s_q8 = stage;
              next();
              mbar_wait(&sm.q_full[stage], qphase);                                     // Kl8
              tc_fence_after();
              group8(smem_u32(sm.qs[s_q8]), smem_u32(sm.qs[stage]));                    // Qh8 . Kl8
              umma_commit_2sm(&sm.q_empty[s_q8]);
              umma_commit_2sm(&sm.q_empty[stage]);
              next();
              if (b == 0) mbar_wait(&sm.k_full[6 + j], kphase);
              mbar_wait(&sm.q_full[stage], qphase);                                     // Ql8
              tc_fence_after();
              group8(smem_u32(sm.qs[stage]), smem_u32(sm.kt[6 + j]));                   // Ql8 . Kh8
              umma_commit_2sm(&sm.q_empty[stage]);
              if (b == nq - 1) umma_commit_2sm(&sm.k_empty[6 + j]);
              next();
            }
          } else {
            for (int kb = 0; kb < kMqKBlocks; ++kb) {
              if (b == 0) mbar_wait(&sm.k_full[kb], kphase);
              const uint32_t kh = smem_u32(sm.kt[kb]);
              mbar_wait(&sm.q_full[stage], qphase);
              tc_fence_after();
              const int s_qh = stage;
              group(tmem_d, smem_u32(sm.qs[s_qh]), kh, kb == 0);                     // Qh . Kh
              next();
              if (SPLIT) {
                mbar_wait(&sm.q_full[stage], qphase);
                tc_fence_after();
                group(tmem_d, smem_u32(sm.qs[s_qh]), smem_u32(sm.qs[stage]), false);  // Qh . Kl
                umma_commit_2sm(&sm.q_empty[s_qh]);
                umma_commit_2sm(&sm.q_empty[stage]);
                next();
                mbar_wait(&sm.q_full[stage], qphase);
                tc_fence_after();
                group(tmem_d, smem_u32(sm.qs[stage]), kh, false);                     // Ql . Kh
                umma_commit_2sm(&sm.q_empty[stage]);
                next();
              } else {
                umma_commit_2sm(&sm.q_empty[s_qh]);                  // token stage free in both CTAs
              }
              if (b == nq - 1) umma_commit_2sm(&sm.k_empty[kb]);     // key slice free once the tile's last query used it
            }
          }
          umma_commit_2sm(&sm.tmem_full[acc]);
        }
        kphase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue (both CTAs, 8 warps) ================================
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    int64_t it = 0;
    for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
      // fused LS: this thread's ray (pass 2: lanes = rays) -- loaded once per tile, used by every query of the batch
      float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
      const int64_t ray = tile * kMqTileRays + rank * 128 + row;
      if (fuse_ls && half == 0 && ray < n_rays) {
        ox = __ldg(rays_ori + ray * 3 + 0); oy = __ldg(rays_ori + ray * 3 + 1); oz = __ldg(rays_ori + ray * 3 + 2);
        dx = __ldg(rays_dir + ray * 3 + 0); dy = __ldg(rays_dir + ray * 3 + 1); dz = __ldg(rays_dir + ray * 3 + 2);
      }
      for (int b = 0; b < nq; ++b, ++it) {
        const int acc = (int)(it & 1);
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&sm.tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256 + half * 128);
        float va[32], vb[32];
        if (PASS == 1) {
          const int64_t col0 = tile * kMqTileRays + half * 128;
          const int valid = (int)min((int64_t)128, max((int64_t)0, n_rays - col0));
          float2 st = sm.stat[b][half][row];
          tmem_ld32(taddr, va);
          tmem_ld_wait(va);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float(&cur)[32] = (c & 1) ? vb : va;
            float(&nxt)[32] = (c & 1) ? va : vb;
            if (c + 1 < 4) tmem_ld32(taddr + (c + 1) * 32, nxt);
            mq_fold(cur, valid - c * 32, st.x, st.y, kSc);
            if (c + 1 < 4) tmem_ld_wait(nxt);
          }
          sm.stat[b][half][row] = st;
        } else {
          float s = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          tmem_ld32(taddr, va);
          tmem_ld_wait(va);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float(&cur)[32] = (c & 1) ? vb : va;
            float(&nxt)[32] = (c & 1) ? va : vb;
            if (c + 1 < 4) tmem_ld32(taddr + (c + 1) * 32, nxt);
            const float4* cc = reinterpret_cast<const float4*>(&sm.cst[b][half * 128 + c * 32]);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 c4 = cc[j4];
              s += mq_ex2(fmaf(cur[j4 * 4 + 0], kSc, -c4.x));
              s1 += mq_ex2(fmaf(cur[j4 * 4 + 1], kSc, -c4.y));
              s2 += mq_ex2(fmaf(cur[j4 * 4 + 2], kSc, -c4.z));
              s3 += mq_ex2(fmaf(cur[j4 * 4 + 3], kSc, -c4.w));
            }
            if (c + 1 < 4) tmem_ld_wait(nxt);
          }
          s = (s + s1) + (s2 + s3);
          if (half == 1) sm.xch[acc][row] = s;
          mq_epi_sync();
          if (half == 0) {
            const float sc = s + sm.xch[acc][row];
            if (ray < n_rays) scores[(int64_t)b * score_stride + ray] = sc;
            if (fuse_ls) {
              // w (I - d d^T), w (I - d d^T) o, w d, w   with w = this ray's score (0 for padding rows)
              const float w = (ray < n_rays) ? sc : 0.f;
              const float dd = dx * ox + dy * oy + dz * oz;
              float v[kMqLs] = {w * (1.f - dx * dx), w * (-dx * dy), w * (-dx * dz), w * (1.f - dy * dy), w * (-dy * dz),
                                w * (1.f - dz * dz), w * (ox - dx * dd), w * (oy - dy * dd), w * (oz - dz * dd),
                                w * dx, w * dy, w * dz, w};
#pragma unroll
              for (int j = 0; j < kMqLs; ++j) v[j] = warp_sum(v[j]);
              if (lane == 0) {
#pragma unroll
                for (int j = 0; j < kMqLs; ++j) sm.ls[b][quad][j] += (double)v[j];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&sm.tmem_empty[acc], 0);
      }
    }
    if (PASS == 1) {
      // merge the two column halves of every (query, token) and write this pair's partial row
      for (int b = 0; b < nq; ++b) {
        const float2 st = sm.stat[b][half][row];
        // every MMA has retired (the last accumulator was drained above), so the ring is idle: stage 0 is scratch
        float(*xch2)[128] = reinterpret_cast<float(*)[128]>(sm.qs[0]);
        if (half == 1) { sm.xch[b & 1][row] = st.x; xch2[b & 1][row] = st.y; }
        mq_epi_sync();
        if (half == 0) {
          const float om = sm.xch[b & 1][row], oz = xch2[b & 1][row];
          const float mn = fmaxf(st.x, om);
          float z = 0.f;
          if (st.x != -INFINITY) z += st.y * mq_ex2(st.x - mn);
          if (om != -INFINITY) z += oz * mq_ex2(om - mn);
          const int tok = (int)rank * 128 + row;
          const int64_t o = ((int64_t)b * n_pairs + pair) * kMaxTokens + tok;
          part_m[o] = (mn == -INFINITY) ? -INFINITY : mn * kMqLn2;  // natural-log units
          part_z[o] = z;
        }
      }
    } else if (fuse_ls) {
      mq_epi_sync();  // every half-0 warp has finished its last accumulation
      if (half == 0 && quad == 0 && lane < kMqLs) {
        for (int b = 0; b < nq; ++b)
          ls_part[((int64_t)b * gridDim.x + blockIdx.x) * kMqLs + lane] =
              ((sm.ls[b][0][lane] + sm.ls[b][1][lane]) + sm.ls[b][2][lane]) + sm.ls[b][3][lane];
      }
    }
  }

  // ================================ teardown ================================
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// sum the per-CTA partial systems in a fixed order: part [nq, n_cta, 13] -> out [nq, 13] (deterministic)
__global__ void mq_ls_reduce_kernel(const double* __restrict__ part, int n_cta, double* __restrict__ out) {
  const int b = blockIdx.x, j = threadIdx.x;
  if (j >= kMqLs) return;
  double a = 0.0;
  for (int c = 0; c < n_cta; ++c) a += part[((int64_t)b * n_cta + c) * kMqLs + j];
  out[b * kMqLs + j] = a;
}

template <int FMT>
int mq_make_map(CUtensorMap* map, const void* base, uint64_t rows) {
  constexpr int row_bytes = MqCfg<FMT>::kRowBytes;
  if (FMT == kFmtF16F8)  // mixed fp16 / e4m3 rows: a byte-typed map (128-byte wide boxes either way)
    return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, base, rows, row_bytes, (uint64_t)row_bytes, "score_tc_mq");
  return make_tmap_2d(map, FMT == kFmtF16x2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows,
                      row_bytes / 2, (uint64_t)row_bytes, "score_tc_mq");
}

inline size_t mq_workspace(int nq) { return (size_t)(nq < 1 ? 1 : nq) * kMaxTokens * 2 * kFeat * 2 + 1024; }

template <int PASS, int FMT>
int mq_launch(const void* kc, int64_t n_rays, const float* q, int nq, int n_img, float* pm, float* pz, const float* m,
              const float* z, float* scores, int64_t score_stride, const float* ori, const float* dir, double* ls_part,
              void* ws, size_t ws_bytes, cudaStream_t s) {
  if (nq < 1 || nq > kMqMaxQ) { set_error("score_tc_mq: n_queries must be in [1, %d]", kMqMaxQ); return SIXDGS_EINVAL; }
  if (ws == nullptr || ws_bytes < mq_workspace(nq)) { set_error("score_tc_mq: workspace too small"); return SIXDGS_EWORKSPACE; }
  if ((reinterpret_cast<uintptr_t>(kc) & 15) != 0) { set_error("score_tc_mq: key cache must be 16-byte aligned"); return SIXDGS_EINVAL; }
  if (n_rays > (int64_t)INT32_MAX - 1024) { set_error("score_tc_mq: n_rays exceeds the TMA coordinate range"); return SIXDGS_EINVAL; }
  void* qb = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  const int64_t nel = (int64_t)nq * kMaxTokens * kFeat;
  if (FMT == kFmtF16x2) mq_qprep_split_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, s>>>(q, n_img, nq, (__half*)qb);
  else if (FMT == kFmtF16F8) mq_qprep_f8_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, s>>>(q, n_img, nq, (uint8_t*)qb);
  else mq_qprep_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, s>>>(q, n_img, nq, (__nv_bfloat16*)qb);
  CUtensorMap mk, mq;
  int rc;
  if ((rc = mq_make_map<FMT>(&mk, kc, (uint64_t)n_rays))) return rc;
  if ((rc = mq_make_map<FMT>(&mq, qb, (uint64_t)nq * kMaxTokens))) return rc;
  const size_t smem = sizeof(MqSmem<MqCfg<FMT>::kStages, MqCfg<FMT>::kResident>) + 1024;
  cudaError_t e = cudaFuncSetAttribute(score_tc_mq_kernel<PASS, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("score_tc_mq attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  score_tc_mq_kernel<PASS, FMT><<<kMqPairs * 2, kMqThreads, smem, s>>>(mk, mq, n_rays, n_img, nq, pm, pz, m, z, scores,
                                                                          score_stride, ori, dir, ls_part);
  return check_launch("score_tc_mq");
}

}  // namespace

// single-query entry points of the fp16-pair formats (api.cu dispatches impl 1 + SIXDGS_F16X2 / SIXDGS_F16F8 here)
int score_tc_split_pass1(const void* kc, int k_dtype, int64_t n_rays, const float* q, int n_img, float* pm, float* pz, void* ws,
                         size_t ws_bytes, cudaStream_t s) {
  if (k_dtype == SIXDGS_F16F8)
    return mq_launch<1, kFmtF16F8>(kc, n_rays, q, 1, n_img, pm, pz, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, ws,
                                   ws_bytes, s);
  return mq_launch<1, kFmtF16x2>(kc, n_rays, q, 1, n_img, pm, pz, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, ws,
                                 ws_bytes, s);
}
int score_tc_split_pass2(const void* kc, int k_dtype, int64_t n_rays, const float* q, int n_img, const float* m, const float* z,
                         float* scores, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (k_dtype == SIXDGS_F16F8)
    return mq_launch<2, kFmtF16F8>(kc, n_rays, q, 1, n_img, nullptr, nullptr, m, z, scores, n_rays, nullptr, nullptr, nullptr,
                                   ws, ws_bytes, s);
  return mq_launch<2, kFmtF16x2>(kc, n_rays, q, 1, n_img, nullptr, nullptr, m, z, scores, n_rays, nullptr, nullptr, nullptr, ws,
                                 ws_bytes, s);
}
size_t score_tc_split_workspace() { return mq_workspace(1); }

}  // namespace sixdgs

using namespace sixdgs;

extern "C" int sixdgs_score_batch_max(void) { return kMqMaxQ; }
extern "C" int sixdgs_score_batch_parts(void) { return kMqPairs; }
extern "C" size_t sixdgs_score_batch_workspace(int n_queries) { return mq_workspace(n_queries); }
extern "C" int sixdgs_ls_partial_rows(void) { return kMqPairs * 2; }

static int mq_check(const void* k_cache, int k_dtype, int64_t n_rays, int n_img) {
  SIXDGS_REQUIRE(k_cache, "null pointer");
  SIXDGS_REQUIRE(k_dtype == SIXDGS_BF16 || k_dtype == SIXDGS_F16X2 || k_dtype == SIXDGS_F16F8,
                 "the batched path needs a bf16, f16x2 or f16f8 key cache");
  SIXDGS_REQUIRE(n_rays >= 1 && n_img >= 1 && n_img <= kMaxTokens, "bad sizes");
  return SIXDGS_OK;
}

template <int PASS>
static int mq_dispatch(int k_dtype, const void* k_cache, int64_t n_rays, const float* q, int nq, int n_img, float* pm, float* pz,
                       const float* m, const float* z, float* scores, int64_t score_stride, const float* ori, const float* dir,
                       double* ls_part, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (k_dtype == SIXDGS_F16X2)
    return mq_launch<PASS, kFmtF16x2>(k_cache, n_rays, q, nq, n_img, pm, pz, m, z, scores, score_stride, ori, dir, ls_part, ws, ws_bytes, s);
  if (k_dtype == SIXDGS_F16F8)
    return mq_launch<PASS, kFmtF16F8>(k_cache, n_rays, q, nq, n_img, pm, pz, m, z, scores, score_stride, ori, dir, ls_part, ws, ws_bytes, s);
  return mq_launch<PASS, kFmtBf16>(k_cache, n_rays, q, nq, n_img, pm, pz, m, z, scores, score_stride, ori, dir, ls_part, ws, ws_bytes, s);
}

extern "C" int sixdgs_score_pass1_batch(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries,
                                        int n_img, float* part_m, float* part_z, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  SIXDGS_REQUIRE(q && part_m && part_z, "null pointer");
  int rc = mq_check(k_cache, k_dtype, n_rays, n_img);
  if (rc) return rc;
  return mq_dispatch<1>(k_dtype, k_cache, n_rays, q, n_queries, n_img, part_m, part_z, nullptr, nullptr, nullptr, 0, nullptr,
                        nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

static int mq_pass2(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries, int n_img,
                    const float* m, const float* z, float* scores, int64_t score_stride, const float* ori, const float* dir,
                    double* ls_part, void* workspace, size_t workspace_bytes, void* stream) {
  SIXDGS_REQUIRE(q && m && z && scores, "null pointer");
  int rc = mq_check(k_cache, k_dtype, n_rays, n_img);
  if (rc) return rc;
  SIXDGS_REQUIRE(score_stride >= n_rays, "score_stride < n_rays");
  return mq_dispatch<2>(k_dtype, k_cache, n_rays, q, n_queries, n_img, nullptr, nullptr, m, z, scores, score_stride, ori, dir,
                        ls_part, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int sixdgs_score_pass2_batch(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries,
                                        int n_img, const float* m, const float* z, float* scores, int64_t score_stride,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  return mq_pass2(k_cache, k_dtype, n_rays, q, n_queries, n_img, m, z, scores, score_stride, nullptr, nullptr, nullptr,
                  workspace, workspace_bytes, stream);
}

extern "C" int sixdgs_score_pass2_batch_ls(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries,
                                           int n_img, const float* m, const float* z, float* scores, int64_t score_stride,
                                           const float* rays_ori, const float* rays_dir, double* ls_part, double* ls_sys,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  SIXDGS_REQUIRE(rays_ori && rays_dir && ls_part && ls_sys, "null pointer");
  int rc = mq_pass2(k_cache, k_dtype, n_rays, q, n_queries, n_img, m, z, scores, score_stride, rays_ori, rays_dir, ls_part,
                    workspace, workspace_bytes, stream);
  if (rc) return rc;
  mq_ls_reduce_kernel<<<n_queries, 32, 0, (cudaStream_t)stream>>>(ls_part, kMqPairs * 2, ls_sys);
  return check_launch("ls_reduce");
}

extern "C" int sixdgs_split_keys(const float* k_f32, int64_t n, void* k_out, float* absmax, void* stream) {
  SIXDGS_REQUIRE(k_f32 && k_out, "null pointer");
  SIXDGS_REQUIRE(n >= 0, "negative size");
  if (n == 0) return SIXDGS_OK;
  const int64_t nel = n * kFeat;
  mq_split_keys_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, (cudaStream_t)stream>>>(k_f32, n, (__half*)k_out,
                                                                                         (unsigned int*)absmax);
  return check_launch("split_keys");
}

extern "C" int sixdgs_keys_f16x2_to_f16f8(void* keys, int64_t n, void* stream) {
  SIXDGS_REQUIRE(keys, "null pointer");
  SIXDGS_REQUIRE(n >= 0, "negative size");
  if (n == 0) return SIXDGS_OK;
  mq_keys_to_f8_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>((uint8_t*)keys, n);
  return check_launch("keys_f16x2_to_f16f8");
}
