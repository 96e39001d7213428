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


// ------------------------------------------------------------------------------------ backward (training)
// d score / d logits of  score_r = sum_i A_ir,  A = softmax over rays of L = q k^T / sqrt(384)
// (reference: autograd through our_multihead_attention.py:4-12 + identification_module.py:80-82, driven by
// train.py:146-176).  With g_r = dLoss/dscore_r:   dL_ir = A_ir (g_r - gbar_i),  gbar_i = sum_r A_ir g_r.
// Two streaming passes over the fp32 keys, same tiles and the same logits as the forward kernels:
//   MODE 1: per-CTA partial gbar rows (merged in a fixed order by score_bwd_reduce_kernel);
//   MODE 2: dlogits = A (g - gbar) / sqrt(384) written row-major [n_rays, 256] and transposed [256, ldt]
//           (the two operand layouts of dk = dlogits q and dq = dlogits^T k, which run as sixdgs_linear GEMMs).
template <int MODE>
__global__ void __launch_bounds__(256)
score_bwd_kernel(const float* __restrict__ kc, int64_t n_rays, const float* __restrict__ q, int n_img,
                 const float* __restrict__ gm, const float* __restrict__ gz, const float* __restrict__ g,
                 const float* __restrict__ gbar, float* __restrict__ part_gbar, float* __restrict__ dl,
                 float* __restrict__ dlt, int64_t ldt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScoreSmem& sm = *reinterpret_cast<ScoreSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tx = tid % (kSBN / 8), ty = tid / (kSBN / 8);
  const int64_t n_tiles = (n_rays + kSBM - 1) / kSBM;
  float tok_m[8], tok_z[8], tok_gb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = tile_col<kSBN>(tx, j);
    tok_m[j] = (t < n_img) ? gm[t] : INFINITY;
    tok_z[j] = (t < n_img) ? gz[t] : INFINITY;
    tok_gb[j] = (MODE == 2 && t < n_img) ? gbar[t] : 0.f;
  }
  float run = 0.f;  // MODE 1: thread t owns token t
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kSBM;
    const int rows = (int)min((int64_t)kSBM, n_rays - r0);
    float acc[8][8];
    gemm_nt_mainloop<kSBM, kSBN, float>(kc + r0 * kFeat, kFeat, rows, q, kFeat, n_img, kFeat, sm.g, acc);
    float gr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = tile_row<kSBM>(ty, i);
      gr[i] = (row < rows) ? g[r0 + row] : 0.f;
    }
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float p = expf(acc[i][j] / kSqrtD - tok_m[j]) / tok_z[j];  // masked tokens: exp(-inf) / inf = 0
          v += (tile_row<kSBM>(ty, i) < rows) ? p * gr[i] : 0.f;
        }
        sm.ex_m[warp][tile_col<kSBN>(tx, j)] = v;
      }
      __syncthreads();
#pragma unroll
      for (int w = 0; w < 8; ++w) run += sm.ex_m[w][tid];
      __syncthreads();
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = tile_row<kSBM>(ty, i);
        if (row >= rows) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = tile_col<kSBN>(tx, j);
          const float p = expf(acc[i][j] / kSqrtD - tok_m[j]) / tok_z[j];
          const float d = (t < n_img) ? p * (gr[i] - tok_gb[j]) / kSqrtD : 0.f;
          dl[(r0 + row) * kSBN + t] = d;
          if (dlt != nullptr) dlt[(int64_t)t * ldt + r0 + row] = d;
        }
      }
    }
  }
  if (MODE == 1) part_gbar[(int64_t)blockIdx.x * kMaxTokens + tid] = run;
}

__global__ void score_bwd_reduce_kernel(const float* __restrict__ part, int n_parts, float* __restrict__ gbar) {
  const int t = threadIdx.x;
  double a = 0.0;
  for (int p = 0; p < n_parts; ++p) a += (double)part[(int64_t)p * kMaxTokens + t];
  gbar[t] = (float)a;
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

// ---- training: backward of the ray score w.r.t. the logits (fp32 keys; see score_bwd_kernel) ----------------------
extern "C" int sixdgs_score_backward_parts(void) { return kSimtParts; }

extern "C" int sixdgs_score_backward_gbar(const float* k_f32, int64_t n_rays, const float* q, int n_img, const float* m,
                                          const float* z, const float* grad_scores, float* part_gbar, float* gbar,
                                          void* stream) {
  SIXDGS_REQUIRE(k_f32 && q && m && z && grad_scores && part_gbar && gbar, "null pointer");
  SIXDGS_REQUIRE(n_rays > 0 && n_img > 0 && n_img <= kMaxTokens, "bad size");
  const size_t smem = sizeof(ScoreSmem);
  cudaError_t e = cudaFuncSetAttribute(score_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("score_bwd attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  score_bwd_kernel<1><<<kSimtParts, 256, smem, (cudaStream_t)stream>>>(k_f32, n_rays, q, n_img, m, z, grad_scores, nullptr,
                                                                       part_gbar, nullptr, nullptr, 0);
  score_bwd_reduce_kernel<<<1, kMaxTokens, 0, (cudaStream_t)stream>>>(part_gbar, kSimtParts, gbar);
  return check_launch("score_backward_gbar");
}

extern "C" int sixdgs_score_backward_dlogits(const float* k_f32, int64_t n_rays, const float* q, int n_img, const float* m,
                                             const float* z, const float* grad_scores, const float* gbar, float* dlogits,
                                             float* dlogits_t, int64_t ldt, void* stream) {
  SIXDGS_REQUIRE(k_f32 && q && m && z && grad_scores && gbar && dlogits, "null pointer");
  SIXDGS_REQUIRE(n_rays > 0 && n_img > 0 && n_img <= kMaxTokens, "bad size");
  SIXDGS_REQUIRE(dlogits_t == nullptr || ldt >= n_rays, "ldt < n_rays");
  const size_t smem = sizeof(ScoreSmem);
  cudaError_t e = cudaFuncSetAttribute(score_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("score_bwd attr: %s", cudaGetErrorString(e)); return SIXDGS_ECUDA; }
  score_bwd_kernel<2><<<kSimtParts, 256, smem, (cudaStream_t)stream>>>(k_f32, n_rays, q, n_img, m, z, grad_scores, gbar,
                                                                       nullptr, dlogits, dlogits_t, ldt);
  return check_launch("score_backward_dlogits");
}
