// a12: top-k of the ray scores (reference identification_module.py:131, torch.topk sorted descending).
// MSB-first 8-bit radix select on order-preserving float keys: 4 histogram sweeps fix the exact
// k-th key, one gather sweep collects everything above it plus the ties, one CTA sorts the k winners
// (ties by lowest index first, deterministically -- also when there are more ties than the tie table holds).
// All control state lives in the caller's workspace, so the whole thing is stream-ordered with no
// host round trip (the reference's torch.topk is also a single device op).
#include "common.cuh"

namespace sixdgs {

constexpr int kTieCap = 4096;
constexpr int kTopkMaxK = 1024;

struct TopkState {
  uint32_t prefix, mask, k_rem, n_gt, n_eq, pad[3];
  uint32_t hist[4][256];
};

__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t b = __float_as_uint(f);
  if (b == 0x80000000u) b = 0u;  // -0.0 == +0.0 (torch.topk compares values, not bit patterns)
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void topk_init_kernel(TopkState* st, int k) {
  const int t = threadIdx.x;
  for (int i = t; i < 4 * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0;
  if (t == 0) { st->prefix = 0; st->mask = 0; st->k_rem = (uint32_t)k; st->n_gt = 0; st->n_eq = 0; }
}

__global__ void __launch_bounds__(256) topk_hist_kernel(const float* __restrict__ x, int64_t n, TopkState* st, int pass) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t prefix = st->prefix, mask = st->mask;
  const int shift = 24 - 8 * pass;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t u = f2key(x[i]);
    if ((u & mask) == prefix) atomicAdd(&h[(u >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&st->hist[pass][threadIdx.x], h[threadIdx.x]);
}

// pick the digit of the k-th key from this pass's histogram: the largest d with  #(digit >= d) >= k_rem  (0 if none).
// One 256-thread CTA, suffix sums in shared memory (the round-1 single-thread walk cost ~10 us per pass).
__global__ void __launch_bounds__(256) topk_select_kernel(TopkState* st, int pass) {
  __shared__ uint32_t suf[257];
  const int t = threadIdx.x;
  const int shift = 24 - 8 * pass;
  const uint32_t k_rem = st->k_rem;
  suf[t] = st->hist[pass][t];
  if (t == 0) suf[256] = 0;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    const uint32_t v = (t + off < 256) ? suf[t + off] : 0u;
    __syncthreads();
    suf[t] += v;
    __syncthreads();
  }
  if ((t == 0 || suf[t] >= k_rem) && suf[t + 1] < k_rem) {  // exactly one thread (suf is non-increasing)
    st->prefix |= ((uint32_t)t) << shift;
    st->mask |= 0xffu << shift;
    st->k_rem = k_rem - suf[t + 1];
  }
}

__global__ void __launch_bounds__(256)
topk_gather_kernel(const float* __restrict__ x, int64_t n, TopkState* st, uint32_t* __restrict__ ckey,
                   int64_t* __restrict__ cidx, int64_t* __restrict__ tie_idx, int k) {
  const uint32_t T = st->prefix;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t u = f2key(x[i]);
    if (u > T) {
      const uint32_t s = atomicAdd(&st->n_gt, 1u);
      if (s < (uint32_t)k) { ckey[s] = u; cidx[s] = i; }
    } else if (u == T) {
      const uint32_t s = atomicAdd(&st->n_eq, 1u);
      if (s < (uint32_t)kTieCap) tie_idx[s] = i;
    }
  }
}

// one CTA: take the k_rem lowest-index ties, then bitonic-sort the k winners by (key desc, idx asc)
__global__ void __launch_bounds__(1024)
topk_final_kernel(TopkState* st, const uint32_t* __restrict__ ckey, const int64_t* __restrict__ cidx,
                  int64_t* __restrict__ tie_idx, int k, float* __restrict__ vals, int64_t* __restrict__ idx) {
  __shared__ uint32_t skey[kTopkMaxK];
  __shared__ long long sidx[kTopkMaxK];
  __shared__ long long stie[kTieCap];
  const int t = threadIdx.x;
  const uint32_t T = st->prefix;
  const int n_gt = (int)min(st->n_gt, (uint32_t)k);
  const int n_tie = (int)min(st->n_eq, (uint32_t)kTieCap);
  const int need = k - n_gt;
  // sort ties by index (bitonic over kTieCap, padded with +inf)
  for (int i = t; i < kTieCap; i += blockDim.x) stie[i] = (i < n_tie) ? tie_idx[i] : 0x7fffffffffffffffLL;
  __syncthreads();
  int tp = 1;
  while (tp < n_tie) tp <<= 1;
  for (int sz = 2; sz <= tp; sz <<= 1)
    for (int st2 = sz >> 1; st2 > 0; st2 >>= 1) {
      for (int i = t; i < tp; i += blockDim.x) {
        const int j = i ^ st2;
        if (j > i) {
          const bool up = ((i & sz) == 0);
          const long long a = stie[i], b = stie[j];
          if ((a > b) == up) { stie[i] = b; stie[j] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = t; i < kTopkMaxK; i += blockDim.x) {
    if (i < n_gt) { skey[i] = ckey[i]; sidx[i] = cidx[i]; }
    else if (i < k && (i - n_gt) < need) { skey[i] = T; sidx[i] = stie[i - n_gt]; }
    else { skey[i] = 0; sidx[i] = 0x7fffffffffffffffLL; }
  }
  __syncthreads();
  for (int sz = 2; sz <= kTopkMaxK; sz <<= 1)
    for (int st2 = sz >> 1; st2 > 0; st2 >>= 1) {
      const int i = t, j = i ^ st2;
      if (j > i) {
        const bool desc = ((i & sz) == 0);
        const uint32_t ka = skey[i], kb = skey[j];
        const long long ia = sidx[i], ib = sidx[j];
        const bool a_first = (ka > kb) || (ka == kb && ia < ib);  // a should precede b in the final order
        if (a_first != desc) { skey[i] = kb; skey[j] = ka; sidx[i] = ib; sidx[j] = ia; }
      }
      __syncthreads();
    }
  if (t < k) { vals[t] = key2f(skey[t]); idx[t] = sidx[t]; }
}

// Ties at the k-th key beyond kTieCap: the gather above records ties in atomic arrival order, which is only harmless
// while ALL of them fit the table (they are sorted by index afterwards).  With more than kTieCap equal keys (many
// zero / underflowed scores, an all-masked query) the survivors would depend on scheduling.  This single-CTA kernel
// exits at once unless that happened; otherwise it rescans the scores in index order and rewrites the table with the
// lowest-index ties -- deterministic, and identical on every rank.  It stops as soon as `need` ties are found.
__global__ void __launch_bounds__(1024)
topk_ties_kernel(const float* __restrict__ x, int64_t n, TopkState* st, int64_t* __restrict__ tie_idx, int k) {
  if (st->n_eq <= (uint32_t)kTieCap) return;
  __shared__ uint32_t wsum[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const uint32_t T = st->prefix;
  const uint32_t need = (uint32_t)k - min(st->n_gt, (uint32_t)k);
  uint32_t found = 0;
  for (int64_t base = 0; base < n && found < need; base += 1024) {
    const int64_t i = base + t;
    const bool tie = i < n && f2key(x[i]) == T;
    const uint32_t bal = __ballot_sync(0xffffffffu, tie);
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int j = 0; j < 32; ++j) { const uint32_t c = wsum[j]; if (j < w) before += c; total += c; }
    const uint32_t pos = found + before + __popc(bal & ((1u << lane) - 1u));
    if (tie && pos < (uint32_t)kTieCap) tie_idx[pos] = i;
    found += total;
    __syncthreads();
  }
  if (t == 0) st->n_eq = min(found, (uint32_t)kTieCap);  // `found` is the same in every thread
}

// n <= 4096 (e.g. the cross-rank candidate table, world * k rows): one CTA, one launch -- load everything,
// bitonic sort by (key desc, index asc), emit the first k.  Same ordering as the radix path.
constexpr int kSmallN = 4096;

__global__ void __launch_bounds__(1024)
topk_small_kernel(const float* __restrict__ x, int n, int k, float* __restrict__ vals, int64_t* __restrict__ idx) {
  __shared__ uint32_t skey[kSmallN];
  __shared__ int sidx[kSmallN];
  const int t = threadIdx.x;
  int np = 1;
  while (np < n) np <<= 1;
  for (int i = t; i < np; i += blockDim.x) {
    skey[i] = (i < n) ? f2key(x[i]) : 0u;
    sidx[i] = (i < n) ? i : 0x7fffffff;
  }
  __syncthreads();
  for (int sz = 2; sz <= np; sz <<= 1)
    for (int st2 = sz >> 1; st2 > 0; st2 >>= 1) {
      for (int i = t; i < np; i += blockDim.x) {
        const int j = i ^ st2;
        if (j > i) {
          const bool in_order = ((i & sz) == 0);
          const uint32_t ka = skey[i], kb = skey[j];
          const int ia = sidx[i], ib = sidx[j];
          const bool a_first = (ka > kb) || (ka == kb && ia < ib);
          if (a_first != in_order) { skey[i] = kb; skey[j] = ka; sidx[i] = ib; sidx[j] = ia; }
        }
      }
      __syncthreads();
    }
  for (int i = t; i < k; i += blockDim.x) { vals[i] = key2f(skey[i]); idx[i] = sidx[i]; }
}

}  // namespace sixdgs

using namespace sixdgs;

extern "C" size_t sixdgs_topk_workspace(int64_t n, int k) {
  (void)n;
  return sizeof(TopkState) + (size_t)k * (sizeof(uint32_t) + sizeof(int64_t)) + kTieCap * sizeof(int64_t) + 64;
}

extern "C" int sixdgs_topk(const float* scores, int64_t n, int k, float* vals, int64_t* idx, void* workspace,
                           size_t workspace_bytes, void* stream) {
  SIXDGS_REQUIRE(scores && vals && idx && workspace, "null pointer");
  SIXDGS_REQUIRE(k >= 1 && k <= kTopkMaxK, "k must be in [1, 1024]");
  SIXDGS_REQUIRE(n >= k, "selected index k out of range");  // torch.topk raises the same condition
  if (workspace_bytes < sixdgs_topk_workspace(n, k)) {
    set_error("topk: workspace too small");
    return SIXDGS_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= kSmallN) {
    topk_small_kernel<<<1, 1024, 0, s>>>(scores, (int)n, k, vals, idx);
    return check_launch("topk_small");
  }
  unsigned char* w = (unsigned char*)workspace;
  TopkState* st = (TopkState*)w;
  w += sizeof(TopkState);
  int64_t* cidx = (int64_t*)w; w += (size_t)k * sizeof(int64_t);
  int64_t* tie = (int64_t*)w; w += kTieCap * sizeof(int64_t);
  uint32_t* ckey = (uint32_t*)w;
  const int64_t want = (n + 256 * 8 - 1) / (256 * 8);
  const unsigned grid = (unsigned)(want < kNumSMs * 8 ? (want > 0 ? want : 1) : kNumSMs * 8);
  topk_init_kernel<<<1, 256, 0, s>>>(st, k);
  for (int p = 0; p < 4; ++p) {
    topk_hist_kernel<<<grid, 256, 0, s>>>(scores, n, st, p);
    topk_select_kernel<<<1, 256, 0, s>>>(st, p);
  }
  topk_gather_kernel<<<grid, 256, 0, s>>>(scores, n, st, ckey, cidx, tie, k);
  topk_ties_kernel<<<1, 1024, 0, s>>>(scores, n, st, tie, k);  // no-op unless more than kTieCap ties at the k-th key
  topk_final_kernel<<<1, 1024, 0, s>>>(st, ckey, cidx, tie, k, vals, idx);
  return check_launch("topk");
}
