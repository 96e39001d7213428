// a11, throughput path: softmax-over-rays attention scores on the 5th-gen tensor cores.
// Reference: pose_estimation/our_multihead_attention.py:4-12,70-79; identification_module.py:80-82.
//
// Work shape.  One query = 256 image tokens x n_rays keys of width 384.  With the key cache in bf16
// the arithmetic intensity is 2*256*384 FLOP per 768-byte key = 256 FLOP/B, i.e. right at the B200
// ridge (~253 FLOP/B from the measured 1.66 PFLOP/s and 6.57 TB/s): the kernel has to keep HBM, the
// tensor pipe and the MUFU (one exp2 per logit) busy at the same time.
//
// Design (persistent, one CTA per SM, CTA pairs = clusters of 2, tcgen05 cta_group::2):
//   * Q (256 x 384 bf16 = 192 KB) does not fit next to a K pipeline in one SM, so each CTA of a pair
//     keeps HALF of Q (128 tokens, 96 KB) resident in shared memory for the whole kernel and the pair
//     issues M=256 x N=256 x K=16 UMMAs that read both halves;
//   * every pair-tile is 256 rays: each CTA TMA-loads its own 128 rays as six 128x64 bf16 k-blocks
//     (16 KB, SWIZZLE_128B) through a 7-stage mbarrier ring -> every key byte crosses HBM->SM once;
//   * accumulators live in TMEM (128 lanes x 256 fp32 columns per CTA), double buffered (512 cols),
//     so the epilogue of tile t overlaps the MMAs of tile t+1;
//   * pass 1 issues D = Q K^T  (lanes = tokens, columns = rays): each epilogue thread owns one token
//     and folds its columns into a running (max, sum-exp2) -- no cross-lane traffic;
//     pass 2 issues D = K Q^T  (lanes = rays, columns = tokens): each thread owns one ray and sums
//     exp2(s - c_token) over the tokens, c = max + log2(sum) broadcast from shared memory.
//     The operands are the same shared-memory tiles in both passes (both are K-major), only the A/B
//     roles swap, which is what lets both reductions stay thread-local.
//   * warp roles: 0 = TMA producer, 1 = MMA issuer (leader CTA only), 2 = TMEM allocator,
//     4..11 = epilogue (two warps per TMEM lane quadrant, 128 columns each).
// log2(e)/sqrt(384) is folded into Q when it is converted to bf16, so the epilogue is FADD + MUFU.EX2
// + FADD per logit.
#include "tc_common.cuh"

namespace sixdgs {

constexpr int kTcStages = 7;
constexpr int kTcTileRays = 256;               // rays per CTA-pair tile
constexpr int kTcKBlocks = kFeat / 64;          // 6 k-blocks of 64 bf16 (128 B rows)
constexpr int kTcKBBytes = 128 * 128;           // 128 rows x 128 B
constexpr int kTcThreads = 384;
constexpr int kTcPairs = kNumSMs / 2;           // 74 clusters
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kQScale = 1.4426950408889634f / 19.595917942265423f;  // log2(e) / sqrt(384)

struct __align__(1024) TcSmem {
  uint8_t q[kTcKBlocks][kTcKBBytes];       //  96 KB: this CTA's 128 tokens
  uint8_t k[kTcStages][kTcKBBytes];        // 112 KB: K pipeline
  uint64_t full[kTcStages];
  uint64_t empty[kTcStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t q_full;
  uint32_t tmem_base;
  uint32_t pad;
  float cst[kMaxTokens];                   // pass 2: c_token = m*log2e + log2 z  (+inf for padded tokens)
  float xch[2][128];                       // column-half exchange
  float xch2[2][128];
};

// ------------------------------------------------------------------------------------ PTX helpers
// kind::f16, bf16 x bf16 -> fp32, A and B K-major, M = 256 (cta_group::2), N = 256
constexpr uint32_t kIdesc = umma_idesc(1, 256, 256);

__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ------------------------------------------------------------------------------------ Q preparation
// Qb[t, :] = bf16(q[t, :] * log2(e)/sqrt(384)) for t < n_img, 0 otherwise  (256 x 384)
__global__ void tc_qprep_kernel(const float* __restrict__ q, int n_img, __nv_bfloat16* __restrict__ qb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kMaxTokens * kFeat) return;
  const int t = i / kFeat;
  qb[i] = __float2bfloat16_rn(t < n_img ? q[i] * kQScale : 0.0f);
}

// ------------------------------------------------------------------------------------ main kernel
template <int PASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_q,
                int64_t n_rays, int n_img,
                float* __restrict__ part_m, float* __restrict__ part_z,     // pass 1 out [pairs, 256]
                const float* __restrict__ gm, const float* __restrict__ gz, // pass 2 in  [256]
                float* __restrict__ scores) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int64_t n_tiles = (n_rays + kTcTileRays - 1) / kTcTileRays;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&sm.full[s], 2);   // leader's expect_tx arrive + peer's remote arrive
      mbar_init(&sm.empty[s], 1);  // one multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sm.tmem_full[a], 1);    // one multicast tcgen05.commit
      mbar_init(&sm.tmem_empty[a], 16);  // 8 epilogue warps x 2 CTAs (lane 0 of each), leader's copy is used
    }
    mbar_init(&sm.q_full, 2);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (PASS == 2) {
    for (int t = tid; t < kMaxTokens; t += kTcThreads)
      sm.cst[t] = (t < n_img) ? (gm[t] * kLog2e + log2f(gz[t])) : INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      // resident half of Q: rows [128*rank, 128*rank+128), six k-blocks, accounted on the leader's q_full
      if (leader) mbar_arrive_expect_tx(&sm.q_full, 2 * kTcKBlocks * kTcKBBytes);
      else mbar_arrive_cluster(&sm.q_full, 0);
      for (int kb = 0; kb < kTcKBlocks; ++kb) tma_load_2sm(sm.q[kb], &tmap_q, &sm.q_full, kb * 64, (int)rank * 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        const int row0 = (int)(tile * kTcTileRays + rank * 128);
        for (int kb = 0; kb < kTcKBlocks; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&sm.full[stage], 2 * kTcKBBytes);
          else mbar_arrive_cluster(&sm.full[stage], 0);
          tma_load_2sm(sm.k[stage], &tmap_k, &sm.full[stage], kb * 64, row0);
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA, one thread) ================================
    if (leader && lane == 0) {
      mbar_wait(&sm.q_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
        const int acc = (int)(it & 1);
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < kTcKBlocks; ++kb) {
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t qa = smem_u32(sm.q[kb]);
          const uint32_t ka = smem_u32(sm.k[stage]);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t dq = umma_desc_sw128(qa + k4 * 32);
            const uint64_t dk = umma_desc_sw128(ka + k4 * 32);
            if (PASS == 1) umma_2sm(tmem_d, dq, dk, (uint32_t)((kb | k4) != 0));  // D[token, ray]
            else umma_2sm(tmem_d, dk, dq, (uint32_t)((kb | k4) != 0));            // D[ray, token]
          }
          umma_commit_2sm(&sm.empty[stage]);  // frees this K stage in both CTAs when the MMAs retire
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&sm.tmem_full[acc]);  // accumulator ready for both CTAs' epilogues
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue (both CTAs, 8 warps) ================================
    const int quad = warp & 3;          // TMEM lane quadrant this warp may touch
    const int half = (warp - 4) >> 2;   // which 128 columns
    const int row = quad * 32 + lane;   // TMEM lane == token (pass 1) / ray within the CTA's 128 (pass 2)
    float run_m = -INFINITY, run_z = 0.f;
    int64_t it = 0;
    for (int64_t tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&sm.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256 + half * 128);
      float va[32], vb[32];
      if (PASS == 1) {
        const int64_t col0 = tile * kTcTileRays + half * 128;  // first ray of this warp's columns
        const int valid = (int)min((int64_t)128, max((int64_t)0, n_rays - col0));
        tmem_ld32(taddr, va);
        tmem_ld_wait(va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float(&cur)[32] = (c & 1) ? vb : va;
          float(&nxt)[32] = (c & 1) ? va : vb;
          if (c + 1 < 4) tmem_ld32(taddr + (c + 1) * 32, nxt);
          const int nv = valid - c * 32;  // columns of this chunk that are real rays
          if (nv >= 32) {
            // full chunk (every tile but the last): 4 independent max / sum chains
            float m0 = fmaxf(cur[0], cur[1]), m1 = fmaxf(cur[2], cur[3]), m2 = fmaxf(cur[4], cur[5]),
                  m3 = fmaxf(cur[6], cur[7]);
#pragma unroll
            for (int j = 8; j < 32; j += 8) {
              m0 = fmaxf(m0, fmaxf(cur[j + 0], cur[j + 1]));
              m1 = fmaxf(m1, fmaxf(cur[j + 2], cur[j + 3]));
              m2 = fmaxf(m2, fmaxf(cur[j + 4], cur[j + 5]));
              m3 = fmaxf(m3, fmaxf(cur[j + 6], cur[j + 7]));
            }
            const float mn = fmaxf(fmaxf(run_m, fmaxf(m0, m1)), fmaxf(m2, m3));
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              s0 += ex2(cur[j + 0] - mn);
              s1 += ex2(cur[j + 1] - mn);
              s2 += ex2(cur[j + 2] - mn);
              s3 += ex2(cur[j + 3] - mn);
            }
            run_z = run_z * ex2(run_m - mn) + ((s0 + s1) + (s2 + s3));
            run_m = mn;
          } else if (nv > 0) {
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = fmaxf(cm, (j < nv) ? cur[j] : -INFINITY);
            const float mn = fmaxf(run_m, cm);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += (j < nv) ? ex2(cur[j] - mn) : 0.f;
            run_z = run_z * ex2(run_m - mn) + s;
            run_m = mn;
          }
          if (c + 1 < 4) tmem_ld_wait(nxt);
        }
      } else {
        float s = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        tmem_ld32(taddr, va);
        tmem_ld_wait(va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float(&cur)[32] = (c & 1) ? vb : va;
          float(&nxt)[32] = (c & 1) ? va : vb;
          if (c + 1 < 4) tmem_ld32(taddr + (c + 1) * 32, nxt);
          const float4* cc = reinterpret_cast<const float4*>(&sm.cst[half * 128 + c * 32]);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 c4 = cc[j4];
            s += ex2(cur[j4 * 4 + 0] - c4.x);
            s1 += ex2(cur[j4 * 4 + 1] - c4.y);
            s2 += ex2(cur[j4 * 4 + 2] - c4.z);
            s3 += ex2(cur[j4 * 4 + 3] - c4.w);
          }
          if (c + 1 < 4) tmem_ld_wait(nxt);
        }
        s = (s + s1) + (s2 + s3);
        // combine the two column halves of each ray, then one coalesced 128-float store per CTA
        if (half == 1) sm.xch[acc][row] = s;
        epi_bar_sync();
        if (half == 0) {
          const int64_t ray = tile * kTcTileRays + rank * 128 + row;
          if (ray < n_rays) scores[ray] = s + sm.xch[acc][row];
        }
      }
      // release this accumulator stage to the MMA issuer (leader CTA's barrier)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&sm.tmem_empty[acc], 0);
    }
    if (PASS == 1) {
      if (half == 1) { sm.xch[0][row] = run_m; sm.xch2[0][row] = run_z; }
      epi_bar_sync();
      if (half == 0) {
        const float om = sm.xch[0][row], oz = sm.xch2[0][row];
        const float mn = fmaxf(run_m, om);
        float z = 0.f;
        if (run_m != -INFINITY) z += run_z * ex2(run_m - mn);
        if (om != -INFINITY) z += oz * ex2(om - mn);
        const int tok = (int)rank * 128 + row;
        part_m[(int64_t)pair * kMaxTokens + tok] = (mn == -INFINITY) ? -INFINITY : mn * kLn2;  // natural-log units
        part_z[(int64_t)pair * kMaxTokens + tok] = z;
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

// ------------------------------------------------------------------------------------ host side
static int make_map(CUtensorMap* map, const void* base, uint64_t rows) {
  return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, kFeat, (uint64_t)kFeat * 2, "score_tc");
}

size_t score_tc_workspace() { return (size_t)kMaxTokens * kFeat * sizeof(__nv_bfloat16) + 1024; }
int score_tc_parts() { return kTcPairs; }

template <int PASS>
static int launch_tc(const void* kc, int64_t n_rays, const float* q, int n_img, float* pm, float* pz, const float* m,
                     const float* z, float* scores, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws == nullptr || ws_bytes < score_tc_workspace()) { set_error("score_tc: workspace too small"); return SIXDGS_EWORKSPACE; }
  if ((reinterpret_cast<uintptr_t>(kc) & 15) != 0) { set_error("score_tc: key cache must be 16-byte aligned"); return SIXDGS_EINVAL; }
  if (n_rays > (int64_t)INT32_MAX - 1024) { set_error("score_tc: n_rays exceeds the TMA coordinate range"); return SIXDGS_EINVAL; }
  __nv_bfloat16* qb = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  tc_qprep_kernel<<<(kMaxTokens * kFeat + 255) / 256, 256, 0, s>>>(q, n_img, qb);
  CUtensorMap mk, mq;
  int rc;
  if ((rc = make_map(&mk, kc, (uint64_t)n_rays))) return rc;
  if ((rc = make_map(&mq, qb, kMaxTokens))) return rc;
  const size_t smem = sizeof(TcSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(score_tc_kernel<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("score_tc attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  score_tc_kernel<PASS><<<kTcPairs * 2, kTcThreads, smem, s>>>(mk, mq, n_rays, n_img, pm, pz, m, z, scores);
  return check_launch("score_tc");
}

int score_tc_pass1(const void* kc, int64_t n_rays, const float* q, int n_img, float* pm, float* pz, void* ws,
                   size_t ws_bytes, cudaStream_t s) {
  return launch_tc<1>(kc, n_rays, q, n_img, pm, pz, nullptr, nullptr, nullptr, ws, ws_bytes, s);
}
int score_tc_pass2(const void* kc, int64_t n_rays, const float* q, int n_img, const float* m, const float* z,
                   float* scores, void* ws, size_t ws_bytes, cudaStream_t s) {
  return launch_tc<2>(kc, n_rays, q, n_img, nullptr, nullptr, m, z, scores, ws, ws_bytes, s);
}

}  // namespace sixdgs
