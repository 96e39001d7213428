// a11, exact-precision path: softmax-over-rays attention scores on fp32 CUDA cores.
// Reference: pose_estimation/our_multihead_attention.py:4-12 (logits / sqrt(d), softmax over rays),
//            :70-79 (q/k projections happen before), identification_module.py:80-82 (sum over tokens).
//
// The reference materialises the [n_img, n_rays] map; a 1M-Gaussian scene has ~29M rays (30 GB).
// Two streaming passes over the key cache instead:
//   pass 1: running per-token (max, sum-exp) over the rays   -> partial rows, merged afterwards
//   pass 2: score[r] = sum_i exp(L_ir - m_i) / z_i
// This file is the parity path (fp32 FMA, true division by sqrt(384), expf); the throughput path is
// the tcgen05 kernel in score_tc.cu.  Persistent grid: 2 CTAs per SM, each looping over 64-ray tiles.
#include "common.cuh"
#include "gemm_simt.cuh"

namespace sixdgs {

constexpr int kSBM = 64;   // rays per tile
constexpr int kSBN = 256;  // tokens
constexpr int kSimtParts = kNumSMs * 2;
constexpr float kSqrtD = 19.595917942265423f;  // sqrt(384)

struct ScoreSmem {
  GemmSmem<kSBM, kSBN> g;
  float ex_m[8][kSBN];
  float ex_z[8][kSBN];
};

template <typename TK, int PASS>
__global__ void __launch_bounds__(256)
score_simt_kernel(const TK* __restrict__ kc, int64_t n_rays, const float* __restrict__ q, int n_img,
                  float* __restrict__ part_m, float* __restrict__ part_z,      // pass 1 out
                  const float* __restrict__ gm, const float* __restrict__ gz,  // pass 2 in
                  float* __restrict__ scores, float* __restrict__ attn_map) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScoreSmem& sm = *reinterpret_cast<ScoreSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid % (kSBN / 8), ty = tid / (kSBN / 8);  // tx == lane, ty == warp
  const int64_t n_tiles = (n_rays + kSBM - 1) / kSBM;

  float run_m = -INFINITY, run_z = 0.f;  // pass 1: thread t owns token t
  float tok_m[8], tok_iz[8];             // pass 2: stats of this thread's 8 tokens
  if (PASS == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = tile_col<kSBN>(tx, j);
      tok_m[j] = (t < n_img) ? gm[t] : 0.f;
      tok_iz[j] = (t < n_img) ? gz[t] : 1.f;
    }
  }

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kSBM;
    const int rows = (int)min((int64_t)kSBM, n_rays - r0);
    float acc[8][8];
    gemm_nt_mainloop<kSBM, kSBN, TK>(kc + r0 * kFeat, kFeat, rows, q, kFeat, n_img, kFeat, sm.g, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = acc[i][j] / kSqrtD;

    if (PASS == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (tile_row<kSBM>(ty, i) < rows) mx = fmaxf(mx, acc[i][j]);
        float z = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (tile_row<kSBM>(ty, i) < rows) z += expf(acc[i][j] - mx);
        const int t = tile_col<kSBN>(tx, j);
        sm.ex_m[warp][t] = mx;
        sm.ex_z[warp][t] = z;
      }
      __syncthreads();
      {
        float mx = run_m;
#pragma unroll
        for (int w = 0; w < 8; ++w) mx = fmaxf(mx, sm.ex_m[w][tid]);
        float z = (run_m == -INFINITY) ? 0.f : run_z * expf(run_m - mx);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float pm = sm.ex_m[w][tid];
          if (pm != -INFINITY) z += sm.ex_z[w][tid] * expf(pm - mx);
        }
        run_m = mx; run_z = z;
      }
      __syncthreads();
    } else {
      float s[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = tile_col<kSBN>(tx, j);
          const float p = (t < n_img) ? expf(acc[i][j] - tok_m[j]) / tok_iz[j] : 0.f;
          if (attn_map != nullptr && t < n_img) {
            const int row = tile_row<kSBM>(ty, i);
            if (row < rows) attn_map[(int64_t)t * n_rays + r0 + row] = p;
          }
          v += p;
        }
        s[i] = warp_sum(v);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = tile_row<kSBM>(ty, i);
          if (row < rows) scores[r0 + row] = s[i];
        }
      }
    }
  }
  if (PASS == 1) {
    part_m[(int64_t)blockIdx.x * kMaxTokens + tid] = run_m;
    part_z[(int64_t)blockIdx.x * kMaxTokens + tid] = run_z;
  }
}

// log-sum-exp merge of partial (max, sum-exp) rows; also merges the rows gathered from other ranks
// Rows are read in `n_groups` groups of `n_parts` consecutive rows, group g starting at row g * group_stride: lets one
// query pick its rows out of an all-gathered [rank][query][part] table without a copy.
__global__ void score_merge_kernel(const float* __restrict__ part_m, const float* __restrict__ part_z, int n_parts,
                                   int n_groups, int64_t group_stride, int n_img, const uint8_t* __restrict__ token_valid,
                                   float* __restrict__ m, float* __restrict__ z) {
  const int t = threadIdx.x;
  if (t >= kMaxTokens) return;
  float mx = -INFINITY;
  for (int g = 0; g < n_groups; ++g)
    for (int p = 0; p < n_parts; ++p) mx = fmaxf(mx, part_m[(g * group_stride + p) * kMaxTokens + t]);
  double acc = 0.0;
  for (int g = 0; g < n_groups; ++g)
    for (int p = 0; p < n_parts; ++p) {
      const int64_t row = g * group_stride + p;
      const float pm = part_m[row * kMaxTokens + t];
      if (pm != -INFINITY) acc += (double)part_z[row * kMaxTokens + t] * (double)expf(pm - mx);
    }
  // masked-out tokens get (m, z) = (+inf, +inf): exp(L - inf) / inf == 0 in pass 2 whatever L is
  const bool live = (t < n_img) && (token_valid == nullptr || token_valid[t] != 0);
  m[t] = live ? mx : INFINITY;
  z[t] = live ? (float)acc : INFINITY;
}

template <typename TK>
static int launch_simt(int pass, const TK* kc, int64_t n_rays, const float* q, int n_img, float* pm, float* pz,
                       const float* m, const float* z, float* scores, float* attn, cudaStream_t s) {
  const size_t smem = sizeof(ScoreSmem);
  cudaError_t e;
  if (pass == 1) {
    e = cudaFuncSetAttribute(score_simt_kernel<TK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("score_simt attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
    score_simt_kernel<TK, 1><<<kSimtParts, 256, smem, s>>>(kc, n_rays, q, n_img, pm, pz, nullptr, nullptr, nullptr, nullptr);
  } else {
    e = cudaFuncSetAttribute(score_simt_kernel<TK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("score_simt attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
    score_simt_kernel<TK, 2><<<kSimtParts, 256, smem, s>>>(kc, n_rays, q, n_img, nullptr, nullptr, m, z, scores, attn);
  }
  return check_launch("score_simt");
}

int score_simt_pass1(const void* kc, int k_dtype, int64_t n_rays, const float* q, int n_img, float* pm, float* pz,
                     cudaStream_t s) {
  if (k_dtype == SIXDGS_F32)
    return launch_simt<float>(1, (const float*)kc, n_rays, q, n_img, pm, pz, nullptr, nullptr, nullptr, nullptr, s);
  return launch_simt<__nv_bfloat16>(1, (const __nv_bfloat16*)kc, n_rays, q, n_img, pm, pz, nullptr, nullptr, nullptr,
                                    nullptr, s);
}
int score_simt_pass2(const void* kc, int k_dtype, int64_t n_rays, const float* q, int n_img, const float* m,
                     const float* z, float* scores, float* attn, cudaStream_t s) {
  if (k_dtype == SIXDGS_F32)
    return launch_simt<float>(2, (const float*)kc, n_rays, q, n_img, nullptr, nullptr, m, z, scores, attn, s);
  return launch_simt<__nv_bfloat16>(2, (const __nv_bfloat16*)kc, n_rays, q, n_img, nullptr, nullptr, m, z, scores,
                                    attn, s);
}
int score_simt_parts() { return kSimtParts; }

}  // namespace sixdgs

using namespace sixdgs;

extern "C" int sixdgs_score_merge(const float* part_m, const float* part_z, int n_parts, int n_groups,
                                  int64_t group_stride, int n_img, const uint8_t* token_valid, float* m, float* z,
                                  void* stream) {
  SIXDGS_REQUIRE(part_m && part_z && m && z, "null pointer");
  SIXDGS_REQUIRE(n_parts > 0 && n_groups > 0 && group_stride >= 0 && n_img > 0 && n_img <= kMaxTokens, "bad size");
  score_merge_kernel<<<1, kMaxTokens, 0, (cudaStream_t)stream>>>(part_m, part_z, n_parts, n_groups, group_stride, n_img,
                                                                 token_valid, m, z);
  return check_launch("score_merge");
}
