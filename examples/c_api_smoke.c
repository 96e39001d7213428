/* Plain-C use of libsixdgs.so (no Python, no torch): least-squares intersection of a few rays, a top-k, and the exact
 * tensor-core ray score (fp32 keys -> f16x2 cache -> two passes) whose scores must sum to the number of tokens.
 * Build (from the repo root):
 *   gcc -std=c99 -Iinclude -I/usr/local/cuda/include examples/c_api_smoke.c -o c_api_smoke \
 *       -L6dgs_b200/csrc -lsixdgs -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/6dgs_b200/csrc
 * It is compiled (not run -- it needs a GPU) by tests/test_abi_and_host.py, which also proves that
 * include/sixdgs.h is valid C and that every declared symbol links. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>

#include "sixdgs.h"

#define CK(call)                                                            \
  do {                                                                      \
    int rc_ = (call);                                                       \
    if (rc_ != SIXDGS_OK) {                                                 \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, sixdgs_last_error());   \
      return 1;                                                             \
    }                                                                       \
  } while (0)

int main(void) {
  printf("libsixdgs ABI v%d, device supported: %d\n", sixdgs_version(), sixdgs_device_supported());
  /* three rays through (1, 2, 3) */
  const float h_o[9] = {0, 0, 0, 5, 0, 0, 0, 7, 1};
  float h_d[9];
  for (int i = 0; i < 3; ++i) {
    float v[3] = {1 - h_o[i * 3], 2 - h_o[i * 3 + 1], 3 - h_o[i * 3 + 2]};
    float n = 0;
    for (int a = 0; a < 3; ++a) n += v[a] * v[a];
    n = 1.0f / sqrtf(n);
    for (int a = 0; a < 3; ++a) h_d[i * 3 + a] = v[a] * n;
  }
  float *d_o, *d_d, *d_c;
  int32_t* d_status;
  double* d_ws;
  cudaMalloc((void**)&d_o, sizeof h_o);
  cudaMalloc((void**)&d_d, sizeof h_d);
  cudaMalloc((void**)&d_c, 3 * sizeof(float));
  cudaMalloc((void**)&d_status, sizeof(int32_t));
  cudaMalloc((void**)&d_ws, 12 * sizeof(double));
  cudaMemcpy(d_o, h_o, sizeof h_o, cudaMemcpyHostToDevice);
  cudaMemcpy(d_d, h_d, sizeof h_d, cudaMemcpyHostToDevice);
  CK(sixdgs_line_intersect(d_o, d_d, NULL, 3, d_c, d_status, d_ws, NULL));
  float c[3];
  cudaMemcpy(c, d_c, sizeof c, cudaMemcpyDeviceToHost);
  printf("centre = (%.4f, %.4f, %.4f)   expected (1, 2, 3)\n", c[0], c[1], c[2]);

  /* top-2 of five scores */
  const float h_s[5] = {0.1f, 0.9f, 0.3f, 0.7f, 0.2f};
  float *d_s, *d_v;
  int64_t* d_i;
  void* d_tw;
  size_t tws = sixdgs_topk_workspace(5, 2);
  cudaMalloc((void**)&d_s, sizeof h_s);
  cudaMalloc((void**)&d_v, 2 * sizeof(float));
  cudaMalloc((void**)&d_i, 2 * sizeof(int64_t));
  cudaMalloc(&d_tw, tws);
  cudaMemcpy(d_s, h_s, sizeof h_s, cudaMemcpyHostToDevice);
  CK(sixdgs_topk(d_s, 5, 2, d_v, d_i, d_tw, tws, NULL));
  int64_t idx[2];
  cudaMemcpy(idx, d_i, sizeof idx, cudaMemcpyDeviceToHost);
  printf("top-2 indices = %lld %lld   expected 1 3\n", (long long)idx[0], (long long)idx[1]);

  /* exact tensor-core score: 1000 pseudo-random keys, one query of 256 tokens */
  enum { NR = 1000, NT = SIXDGS_MAX_TOKENS, D = SIXDGS_FEAT };
  float* h_k = (float*)malloc(sizeof(float) * NR * D);
  float* h_q = (float*)malloc(sizeof(float) * NT * D);
  unsigned s = 12345u;
  for (int i = 0; i < NR * D; ++i) { s = s * 1664525u + 1013904223u; h_k[i] = ((s >> 8) / 8388608.0f - 1.0f) * 0.7f; }
  for (int i = 0; i < NT * D; ++i) { s = s * 1664525u + 1013904223u; h_q[i] = ((s >> 8) / 8388608.0f - 1.0f) * 3.0f; }
  float *d_k, *d_q, *d_pm, *d_pz, *d_m, *d_z, *d_sc, *d_absmax;
  void *d_keys, *d_sws;
  const int parts = sixdgs_score_parts(1);
  const size_t sws = sixdgs_score_workspace(1);
  cudaMalloc((void**)&d_k, sizeof(float) * NR * D);
  cudaMalloc((void**)&d_q, sizeof(float) * NT * D);
  cudaMalloc(&d_keys, (size_t)NR * 2 * D * 2);               /* SIXDGS_F16X2: [hi(384) | lo(384)] fp16 per ray */
  cudaMalloc((void**)&d_pm, sizeof(float) * parts * NT);
  cudaMalloc((void**)&d_pz, sizeof(float) * parts * NT);
  cudaMalloc((void**)&d_m, sizeof(float) * NT);
  cudaMalloc((void**)&d_z, sizeof(float) * NT);
  cudaMalloc((void**)&d_sc, sizeof(float) * NR);
  cudaMalloc((void**)&d_absmax, sizeof(float));
  cudaMalloc(&d_sws, sws);
  cudaMemset(d_absmax, 0, sizeof(float));
  cudaMemcpy(d_k, h_k, sizeof(float) * NR * D, cudaMemcpyHostToDevice);
  cudaMemcpy(d_q, h_q, sizeof(float) * NT * D, cudaMemcpyHostToDevice);
  CK(sixdgs_split_keys(d_k, NR, d_keys, d_absmax, NULL));
  CK(sixdgs_score_pass1(d_keys, SIXDGS_F16X2, NR, d_q, NT, d_pm, d_pz, 1, d_sws, sws, NULL));
  CK(sixdgs_score_merge(d_pm, d_pz, parts, 1, 0, NT, NULL, d_m, d_z, NULL));
  CK(sixdgs_score_pass2(d_keys, SIXDGS_F16X2, NR, d_q, NT, d_m, d_z, d_sc, NULL, 1, d_sws, sws, NULL));
  float* h_sc = (float*)malloc(sizeof(float) * NR);
  cudaMemcpy(h_sc, d_sc, sizeof(float) * NR, cudaMemcpyDeviceToHost);
  double sum = 0.0;
  for (int i = 0; i < NR; ++i) sum += h_sc[i];
  printf("sum of the ray scores = %.4f   expected %d (every token's softmax sums to one)\n", sum, NT);
  free(h_k); free(h_q); free(h_sc);
  return 0;
}
