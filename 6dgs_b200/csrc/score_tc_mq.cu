// a11, several queries per key sweep (SURVEY §8d "ray-score, batch" row).  EXPERIMENTAL: built and exported, not on
// any default path until it has been validated and timed on a B200 (tests/test_experimental.py).
// Reference: pose_estimation/our_multihead_attention.py:4-12,70-79; identification_module.py:80-82.
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
#include "tc_common.cuh"

namespace sixdgs {

namespace {

constexpr int kMqMaxQ = 8;
constexpr int kMqStages = 6;
constexpr int kMqTileRays = 256;
constexpr int kMqKBlocks = kFeat / 64;  // 6
constexpr int kMqKBBytes = 128 * 128;   // 128 rows x 128 B
constexpr int kMqThreads = 384;
constexpr int kMqPairs = kNumSMs / 2;
constexpr float kMqLog2e = 1.4426950408889634f;
constexpr float kMqLn2 = 0.6931471805599453f;
constexpr float kMqQScale = 1.4426950408889634f / 19.595917942265423f;  // log2(e) / sqrt(384)

struct __align__(1024) MqSmem {
  uint8_t kt[kMqKBlocks][kMqKBBytes];  //  96 KB: this CTA's 128 rays of the current tile
  uint8_t qs[kMqStages][kMqKBBytes];   //  96 KB: token ring
  uint64_t k_full[kMqKBlocks];
  uint64_t k_empty[kMqKBlocks];
  uint64_t q_full[kMqStages];
  uint64_t q_empty[kMqStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  union __align__(16) {
    float cst[kMqMaxQ][kMaxTokens];   // pass 2: c_token = m*log2e + log2 z per query (+inf: token unused)
    float2 stat[kMqMaxQ][2][128];     // pass 1: (running max, running sum) per query / column half / token row
  };
  float xch[2][128];
  float xch2[2][128];
};

constexpr uint32_t kMqIdesc = umma_idesc(1, 256, 256);  // bf16 x bf16 -> fp32, M = 256 over the pair, N = 256

__device__ __forceinline__ void mq_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kMqIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float mq_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void mq_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// fold 32 logits (one TMEM chunk) of one token into its running (max, sum-exp2); nv = real rays in the chunk
__device__ __forceinline__ void mq_fold(const float (&cur)[32], int nv, float& run_m, float& run_z) {
  if (nv >= 32) {
    float m0 = fmaxf(cur[0], cur[1]), m1 = fmaxf(cur[2], cur[3]), m2 = fmaxf(cur[4], cur[5]), m3 = fmaxf(cur[6], cur[7]);
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
      s0 += mq_ex2(cur[j + 0] - mn);
      s1 += mq_ex2(cur[j + 1] - mn);
      s2 += mq_ex2(cur[j + 2] - mn);
      s3 += mq_ex2(cur[j + 3] - mn);
    }
    run_z = run_z * mq_ex2(run_m - mn) + ((s0 + s1) + (s2 + s3));
    run_m = mn;
  } else if (nv > 0) {
    float cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) cm = fmaxf(cm, (j < nv) ? cur[j] : -INFINITY);
    const float mn = fmaxf(run_m, cm);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += (j < nv) ? mq_ex2(cur[j] - mn) : 0.f;
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

template <int PASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMqThreads, 1)
score_tc_mq_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_q,
                   int64_t n_rays, int n_img, int nq,
                   float* __restrict__ part_m, float* __restrict__ part_z,      // pass 1 out [nq, pairs, 256]
                   const float* __restrict__ gm, const float* __restrict__ gz,  // pass 2 in  [nq, 256]
                   float* __restrict__ scores, int64_t score_stride) {          // pass 2 out [nq, score_stride]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MqSmem& sm = *reinterpret_cast<MqSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int64_t n_tiles = (n_rays + kMqTileRays - 1) / kMqTileRays;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMqKBlocks; ++s) {
      mbar_init(&sm.k_full[s], 2);   // leader's expect_tx arrive + peer's remote arrive
      mbar_init(&sm.k_empty[s], 1);  // one multicast tcgen05.commit (after the tile's last query)
    }
    for (int s = 0; s < kMqStages; ++s) {
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
  } else {
    for (int i = tid; i < nq * 2 * 128; i += kMqThreads) (&sm.stat[0][0][0])[i] = make_float2(-INFINITY, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ================================ key producer (both CTAs) ================================
    if (lane == 0) {
      uint32_t phase = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        const int row0 = (int)(tile * kMqTileRays + rank * 128);
        for (int kb = 0; kb < kMqKBlocks; ++kb) {
          mbar_wait(&sm.k_empty[kb], phase ^ 1);  // the previous tile's last query is done with this slice
          if (leader) mbar_arrive_expect_tx(&sm.k_full[kb], 2 * kMqKBBytes);
          else mbar_arrive_cluster(&sm.k_full[kb], 0);
          tma_load_2sm(sm.kt[kb], &tmap_k, &sm.k_full[kb], kb * 64, row0);
        }
        phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ================================ token producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        for (int b = 0; b < nq; ++b) {
          const int row0 = b * kMaxTokens + (int)rank * 128;
          for (int kb = 0; kb < kMqKBlocks; ++kb) {
            mbar_wait(&sm.q_empty[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&sm.q_full[stage], 2 * kMqKBBytes);
            else mbar_arrive_cluster(&sm.q_full[stage], 0);
            tma_load_2sm(sm.qs[stage], &tmap_q, &sm.q_full[stage], kb * 64, row0);
            if (++stage == kMqStages) { stage = 0; phase ^= 1; }
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
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        for (int b = 0; b < nq; ++b, ++it) {
          const int acc = (int)(it & 1);
          const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
          mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
          for (int kb = 0; kb < kMqKBlocks; ++kb) {
            if (b == 0) mbar_wait(&sm.k_full[kb], kphase);
            mbar_wait(&sm.q_full[stage], qphase);
            tc_fence_after();
            const uint32_t qa = smem_u32(sm.qs[stage]);
            const uint32_t ka = smem_u32(sm.kt[kb]);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t dq = umma_desc_sw128(qa + k4 * 32);
              const uint64_t dk = umma_desc_sw128(ka + k4 * 32);
              if (PASS == 1) mq_umma(tmem_d, dq, dk, (uint32_t)((kb | k4) != 0));  // D[token, ray]
              else mq_umma(tmem_d, dk, dq, (uint32_t)((kb | k4) != 0));            // D[ray, token]
            }
            umma_commit_2sm(&sm.q_empty[stage]);                   // token stage free in both CTAs
            if (b == nq - 1) umma_commit_2sm(&sm.k_empty[kb]);     // key slice free once the tile's last query used it
            if (++stage == kMqStages) { stage = 0; qphase ^= 1; }
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
            mq_fold(cur, valid - c * 32, st.x, st.y);
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
              s += mq_ex2(cur[j4 * 4 + 0] - c4.x);
              s1 += mq_ex2(cur[j4 * 4 + 1] - c4.y);
              s2 += mq_ex2(cur[j4 * 4 + 2] - c4.z);
              s3 += mq_ex2(cur[j4 * 4 + 3] - c4.w);
            }
            if (c + 1 < 4) tmem_ld_wait(nxt);
          }
          s = (s + s1) + (s2 + s3);
          if (half == 1) sm.xch[acc][row] = s;
          mq_epi_sync();
          if (half == 0) {
            const int64_t ray = tile * kMqTileRays + rank * 128 + row;
            if (ray < n_rays) scores[(int64_t)b * score_stride + ray] = s + sm.xch[acc][row];
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
        if (half == 1) { sm.xch[b & 1][row] = st.x; sm.xch2[b & 1][row] = st.y; }
        mq_epi_sync();
        if (half == 0) {
          const float om = sm.xch[b & 1][row], oz = sm.xch2[b & 1][row];
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

int mq_make_map(CUtensorMap* map, const void* base, uint64_t rows) {
  return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, kFeat, (uint64_t)kFeat * 2, "score_tc_mq");
}

template <int PASS>
int mq_launch(const void* kc, int64_t n_rays, const float* q, int nq, int n_img, float* pm, float* pz, const float* m,
              const float* z, float* scores, int64_t score_stride, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (nq < 1 || nq > kMqMaxQ) { set_error("score_tc_mq: n_queries must be in [1, %d]", kMqMaxQ); return SIXDGS_EINVAL; }
  if (ws == nullptr || ws_bytes < (size_t)nq * kMaxTokens * kFeat * sizeof(__nv_bfloat16) + 1024) {
    set_error("score_tc_mq: workspace too small");
    return SIXDGS_EWORKSPACE;
  }
  if ((reinterpret_cast<uintptr_t>(kc) & 15) != 0) { set_error("score_tc_mq: key cache must be 16-byte aligned"); return SIXDGS_EINVAL; }
  if (n_rays > (int64_t)INT32_MAX - 1024) { set_error("score_tc_mq: n_rays exceeds the TMA coordinate range"); return SIXDGS_EINVAL; }
  __nv_bfloat16* qb = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  const int64_t nel = (int64_t)nq * kMaxTokens * kFeat;
  mq_qprep_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, s>>>(q, n_img, nq, qb);
  CUtensorMap mk, mq;
  int rc;
  if ((rc = mq_make_map(&mk, kc, (uint64_t)n_rays))) return rc;
  if ((rc = mq_make_map(&mq, qb, (uint64_t)nq * kMaxTokens))) return rc;
  const size_t smem = sizeof(MqSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(score_tc_mq_kernel<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("score_tc_mq attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  score_tc_mq_kernel<PASS><<<kMqPairs * 2, kMqThreads, smem, s>>>(mk, mq, n_rays, n_img, nq, pm, pz, m, z, scores, score_stride);
  return check_launch("score_tc_mq");
}

}  // namespace

}  // namespace sixdgs

using namespace sixdgs;

extern "C" int sixdgs_score_batch_max(void) { return kMqMaxQ; }
extern "C" int sixdgs_score_batch_parts(void) { return kMqPairs; }
extern "C" size_t sixdgs_score_batch_workspace(int n_queries) {
  return (size_t)(n_queries < 1 ? 1 : n_queries) * kMaxTokens * kFeat * sizeof(__nv_bfloat16) + 1024;
}

extern "C" int sixdgs_score_pass1_batch(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries,
                                        int n_img, float* part_m, float* part_z, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  SIXDGS_REQUIRE(k_cache && q && part_m && part_z, "null pointer");
  SIXDGS_REQUIRE(k_dtype == SIXDGS_BF16, "the batched path needs a bf16 key cache");
  SIXDGS_REQUIRE(n_rays >= 1 && n_img >= 1 && n_img <= kMaxTokens, "bad sizes");
  return mq_launch<1>(k_cache, n_rays, q, n_queries, n_img, part_m, part_z, nullptr, nullptr, nullptr, 0, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}

extern "C" int sixdgs_score_pass2_batch(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries,
                                        int n_img, const float* m, const float* z, float* scores, int64_t score_stride,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  SIXDGS_REQUIRE(k_cache && q && m && z && scores, "null pointer");
  SIXDGS_REQUIRE(k_dtype == SIXDGS_BF16, "the batched path needs a bf16 key cache");
  SIXDGS_REQUIRE(n_rays >= 1 && n_img >= 1 && n_img <= kMaxTokens && score_stride >= n_rays, "bad sizes");
  return mq_launch<2>(k_cache, n_rays, q, n_queries, n_img, nullptr, nullptr, m, z, scores, score_stride, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}
