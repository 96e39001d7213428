/*
 * sixdgs.h -- C ABI of libsixdgs.so: the B200 (sm_100a) kernels behind the 6DGS single-query
 * pose-estimation hot path.
 *
 * The reference (mbortolon97/6dgs) has NO native boundary on this path -- it is pure torch ops in
 * the pose_estimation package -- so the entry points below are what a ctypes/pybind binding inside the
 * reference's Python functions would call.  The calling convention follows the
 * reference's only native precedent, free functions over raw device pointers
 * (submodules/simple-knn/simple_knn.h:18, spatial.cu:15-26).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory
 *     (inputs, outputs, workspaces); the library never allocates or frees device memory and keeps
 *     no global state except a thread-local last-error string;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit device sync;
 *   - return 0 on success, negative SIXDGS_E* otherwise; never throws across the ABI;
 *   - numerical singularities are not errors: they follow the reference conventions (NaN centre
 *     when det(R) < 1e-7, identity rotation when the frame is singular) and raise a status flag;
 *   - row-major fp32 everywhere unless a dtype argument says otherwise; ray/ellipsoid indices int64
 *     (the reference uses torch.long).
 */
#ifndef SIXDGS_H
#define SIXDGS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIXDGS_OK 0
#define SIXDGS_EINVAL (-1)   /* bad argument (null pointer, negative size, unsupported dtype)   */
#define SIXDGS_ECUDA (-2)    /* a CUDA runtime call or launch failed; see sixdgs_last_error()   */
#define SIXDGS_EWORKSPACE (-3) /* workspace too small                                            */
#define SIXDGS_EUNSUPPORTED (-4) /* device is not sm_100 / feature not available                 */

#define SIXDGS_F32 0
#define SIXDGS_BF16 1
/* exact tensor-core key format: a row is [hi(384) | lo(384)] fp16 (1536 B) holding 16*k as hi + lo, hi = fp16(16 k),
 * lo = fp16(16 k - hi): 22 significant bits.  |k| must stay below 4094 (sixdgs_split_keys reports the maximum). */
#define SIXDGS_F16X2 2
/* fast variant of it: a row is [hi fp16 (768 B) | e4m3(hi / 64) (384 B) | e4m3(64 lo) (384 B)]; the two cross terms of the
 * logit run as e4m3 MMAs at twice the rate (scores to ~3e-4 at a logit standard deviation of 6.5, growing linearly
 * with it: inside the 1e-3 bar up to ~12).  Built from an F16X2 cache in place by sixdgs_keys_f16x2_to_f16f8. */
#define SIXDGS_F16F8 3

#define SIXDGS_FEAT 384      /* ray / image embedding width (DINOv2 ViT-S/14), backbone.py:17  */
#define SIXDGS_MAX_TOKENS 256 /* 16x16 backbone grid, backbone.py:16                            */

int sixdgs_version(void);
const char* sixdgs_last_error(void);
/* 1 if the current device can run the tcgen05/TMA kernels (compute capability 10.x). */
int sixdgs_device_supported(void);

/* ---- a2: degrade mask + ring layout -------- reference pose_estimation/quadricell.py:171-188 ----
 * scaling_raw[n,3] is the LOG scale (GaussianModel._scaling; get_scaling = exp, gaussian_model.py:125).
 * valid[i] = rings(i) < target.  rings_out (nullable) receives the slab count. */
int sixdgs_degrade_mask(const float* scaling_raw, int64_t n, int target_points,
                        uint8_t* valid, int32_t* rings_out, void* stream);

/* ---- a4: exact k-nearest-neighbour normals -------- pose_estimation/sampling.py:37-113 ----------
 * cloud[m,3]; for query rows [q_begin, q_begin+q_count) writes normals_out[q_count,3]: the
 * smallest-eigenvalue eigenvector (a5) of the centred k-NN scatter matrix, majority-sign
 * disambiguated, normalised.  k <= 32.  Brute force, exact (the reference uses cdist + topk). */
int sixdgs_knn_normals(const float* cloud, int64_t m, int64_t q_begin, int64_t q_count, int k,
                       float* normals_out, void* stream);

/* Same result on a uniform grid, O(m) instead of O(m^2) (exhaustive shell search, exact).  The caller supplies the
 * grid: lower corner grid_lo_host[3], cubic cell edge `cell`, cells per axis dims_host[3] (HOST values; every point
 * must lie inside, points outside are clamped to the border cells which keeps the search exact only if none is).
 * workspace >= sixdgs_knn_grid_workspace(m, dims[0]*dims[1]*dims[2]). */
size_t sixdgs_knn_grid_workspace(int64_t m, int64_t n_cells);
int sixdgs_knn_normals_grid(const float* cloud, int64_t m, int64_t q_begin, int64_t q_count, int k,
                            const float* grid_lo_host, float cell, const int* dims_host, float* normals_out,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- a5: batched closed-form symmetric 3x3 eigen-decomposition --- pose_estimation/sym_eig_3x3.py:246-307
 * A[n,3,3] -> vals[n,3] ascending, vecs[n,3,3] (eigenvectors as COLUMNS; nullable). */
int sixdgs_sym_eig3x3(const float* A, int64_t n, float eps, float* vals, float* vecs, void* stream);

/* ---- a6+a7+a8: candidate-ray generation -------- quadricell.py:191-386, sampling.py:116-124,176-251,
 *                                                  utils/sh_utils.py:55-118, general_utils.py:103-126
 * sel[m] = global Gaussian ids of the selected (valid, permuted) ellipsoids; normals[m,3] from a4.
 * Every ring's arc-length table is built once:
 *   1. sixdgs_raygen_cells:   cells_per_ell[m] (ring layout only; an upper bound of the ellipsoid's rays)
 *   2. sixdgs_exclusive_scan: slot_offset[m+1]; the caller sizes scratch arrays for slot_offset[m] rays
 *   3. sixdgs_raygen_fill:    writes each ellipsoid's surviving rays densely from slot_offset[e] into the scratch
 *                             ori/dir/rgb[., 3] and ell_id[.] (local index into sel; nullable) in ring-, cell-major
 *                             order, and rays_per_ell[m]
 *   4. sixdgs_exclusive_scan: ray_offset[m+1]  ->  5. sixdgs_raygen_compact: the gap-free final arrays.
 * mode 0 = rays (rotate, hemisphere quirk n_x*p'_x > 0, normalise, +mu, SH colour);
 * mode 1 = raw quadricell cells (a6 only: un-rotated points -> ori, no mask; dir/rgb untouched). */
int sixdgs_raygen_cells(const float* scaling_raw, const int64_t* sel, int64_t m, int target_points,
                        int32_t* cells_per_ell, void* stream);
int sixdgs_raygen_fill(const float* xyz, const float* scaling_raw, const float* rotation_raw,
                       const float* features /* [N,sh_coeffs,3] = get_features */, int sh_degree,
                       int sh_coeffs /* stored coefficients per Gaussian, >= (sh_degree+1)^2; 16 for degree-3 storage */,
                       const int64_t* sel, int64_t m, const float* normals, int target_points,
                       int resolution, int mode, const int64_t* slot_offset, float* ori, float* dir,
                       float* rgb, int64_t* ell_id, int32_t* rays_per_ell, void* stream);
int sixdgs_raygen_compact(const int64_t* slot_offset, const int64_t* ray_offset, const int32_t* rays_per_ell, int64_t m,
                          const float* tmp_ori, const float* tmp_dir, const float* tmp_rgb, const int64_t* tmp_ell,
                          float* ori, float* dir, float* rgb, int64_t* ell_id, void* stream);
/* exclusive scan int32 -> int64 with the total in out[n] (out has n+1 entries).  Single block. */
int sixdgs_exclusive_scan(const int32_t* in, int64_t n, int64_t* out, void* stream);

/* ---- a10 (+ k_proj of a11): ray features -> key cache -------- ray_preprocessor.py:3-46,
 *                                                               our_multihead_attention.py:75
 * Packed weights (built once by the host, zero padded so every K dim is a multiple of 32):
 *   w1p[512,160]  = mlp.0.weight  (141 -> 160)        b1[512]
 *   w2 [512,512]  = mlp.2.weight                      b2[512]
 *   w3p[512,672]  = mlp2.0.weight ([h512 | x141] -> 672) b3[512]
 *   w4 [384,512]  = mlp2.2.weight                     b4[384]
 *   wk [384,384]  = attention.k_proj.weight (nullable: skip projection, emit features) bk[384]
 * k_out[n,384] in k_dtype (SIXDGS_F32 | SIXDGS_BF16); feat_out (nullable) = pre-projection features.
 * impl: 0 = fp32 FMA GEMMs (exact path), 1 = TF32 tcgen05 GEMMs (TMA + TMEM, staged TMA-store epilogue; throughput
 *       path), 3 = the same GEMMs with a direct-store epilogue (bit-identical; kept as the comparison kernel).
 * workspace >= sixdgs_ray_features_workspace(n) bytes. */
size_t sixdgs_ray_features_workspace(int64_t n);
int sixdgs_ray_features(const float* ori, const float* dir, const float* rgb, int64_t n,
                        const float* w1p, const float* b1, const float* w2, const float* b2,
                        const float* w3p, const float* b3, const float* w4, const float* b4,
                        const float* wk, const float* bk, void* k_out, int k_dtype, float* feat_out,
                        int impl, void* workspace, size_t workspace_bytes, void* stream);

/* Exact tensor-core build of a SIXDGS_F16X2 key cache: the same five layers as three-term split-fp16 tcgen05 GEMMs
 * (every activation and weight carried as an fp16 pair hi + lo; fp32-grade keys, ~6x faster than impl 0).
 * Weights in "x2" layout, a row of width W stored as [hi(W) | lo(W)] fp16 with hi = fp16(w), lo = fp16(w - hi):
 *   w1 [512, 2*192] (mlp.0, 141 -> 192 zero padded), w2 [512, 2*512], w3 [512, 2*704] (mlp2.0: [h 512 | x 141 -> 192]),
 *   w4 [384, 2*512], wk [384, 2*384]; biases fp32.  k_out [n, 768] fp16; absmax as in sixdgs_split_keys (nullable).
 * workspace >= sixdgs_ray_features_x2_workspace(n) bytes. */
size_t sixdgs_ray_features_x2_workspace(int64_t n);
int sixdgs_ray_features_x2(const float* ori, const float* dir, const float* rgb, int64_t n, const void* w1,
                           const float* b1, const void* w2, const float* b2, const void* w3, const float* b3,
                           const void* w4, const float* b4, const void* wk, const float* bk, void* k_out,
                           float* absmax, void* workspace, size_t workspace_bytes, void* stream);

/* generic y[m,n] = act(x[m,k] w[n,k]^T + b[n]); k % 16 == 0, lda/ldc in elements (used for q_proj,
 * our_multihead_attention.py:74, with img features padded 398 -> 400; for the dk / dq products of the score backward;
 * and, in training mode, for the forward, dx and dW products of the ray MLP and the two projections -- the autograd of
 * ray_preprocessor.py:36-46 / our_multihead_attention.py:74-75 under train.py:146-176, see 6dgs_b200/train_ops.py). */
int sixdgs_linear(const float* x, int64_t m, int k, int lda, const float* w, const float* b, int n,
                  float* y, int ldc, int relu, void* stream);

/* ---- a11: softmax-over-rays attention score -------- our_multihead_attention.py:4-12,70-79;
 *                                                      identification_module.py:80-82
 * q[n_img,384] fp32 (already projected), K cache [n_rays,384] in k_dtype.
 * pass1: per-token running (max, sum-exp) of logits q.k/sqrt(384) over this K shard, written as
 *        `n_parts` partial rows part_m/part_z[n_parts, 256] (n_parts = sixdgs_score_parts()).
 * merge: log-sum-exp merge of partial rows (also used across ranks after an all-gather) -> m,z[256];
 *        tokens >= n_img or with token_valid == 0 get (m, z) = (+inf, +inf) so that pass 2 ignores them
 *        (lets a masked query run on all 256 grid tokens without a host-side compaction / sync).
 * pass2: scores[r] = sum_i exp(L_ir - m_i) / z_i; attn_map (nullable) [n_img, n_rays].
 * impl: 0 = SIMT fp32 (exact path, fp32 or bf16 K), 1 = tcgen05 tensor cores: bf16 K (throughput mode, one MMA
 *       term, scores to ~3e-2) or f16x2 K (exact mode: three fp16 MMA terms Qh.Kh + Qh.Kl + Ql.Kh in one fp32
 *       accumulator, scores to ~1e-5 of the fp32 reference). */
int sixdgs_score_parts(int impl);
/* bytes of scratch the chosen impl needs (impl 1: the bf16, pre-scaled copy of q the TMA reads; impl 0: 0) */
size_t sixdgs_score_workspace(int impl);
int sixdgs_score_pass1(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_img,
                       float* part_m, float* part_z, int impl, void* workspace, size_t workspace_bytes,
                       void* stream);
int sixdgs_score_merge(const float* part_m, const float* part_z, int n_parts, int n_groups,
                       int64_t group_stride /* rows between groups; groups of n_parts consecutive rows */, int n_img,
                       const uint8_t* token_valid /* nullable [256]: 0 = token masked out */, float* m,
                       float* z, void* stream);
int sixdgs_score_pass2(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_img,
                       const float* m, const float* z, float* scores, float* attn_map, int impl,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- a11 backward (training: autograd through our_multihead_attention.py:4-12 + identification_module.py:80-82,
 * driven by pose_estimation/train.py:146-176).  fp32 keys.  With g = dLoss/dscores [n_rays] and the forward's (m, z):
 *   gbar[t]       = sum_r A[t,r] g[r]                                   (part_gbar: sixdgs_score_backward_parts() x 256 scratch)
 *   dlogits[r,t]  = A[t,r] (g[r] - gbar[t]) / sqrt(384)                 row-major [n_rays,256]; dlogits_t (nullable) the
 *                   same values transposed, [256, ldt] with ldt >= n_rays
 * The parameter gradients are then two GEMMs (sixdgs_linear): dk = dlogits q, dq = dlogits^T k. */
int sixdgs_score_backward_parts(void);
int sixdgs_score_backward_gbar(const float* k_f32, int64_t n_rays, const float* q, int n_img, const float* m,
                               const float* z, const float* grad_scores, float* part_gbar, float* gbar, void* stream);
int sixdgs_score_backward_dlogits(const float* k_f32, int64_t n_rays, const float* q, int n_img, const float* m,
                                  const float* z, const float* grad_scores, const float* gbar, float* dlogits,
                                  float* dlogits_t, int64_t ldt, void* stream);

/* fp32 keys [n,384] -> SIXDGS_F16X2 rows [n,768]; absmax (nullable, device, zero-initialised by the caller) receives
 * max |16 k| over the converted rows so the caller can check the fp16 range (must stay < 65504). */
int sixdgs_split_keys(const float* k_f32, int64_t n, void* k_out, float* absmax, void* stream);

int sixdgs_keys_f16x2_to_f16f8(void* keys /* [n,768] fp16 in, [n,1536] bytes out, in place */, int64_t n, void* stream);

/* ---- a11, several queries per key sweep -------- our_multihead_attention.py:4-12,70-79;
 *                                                  identification_module.py:80-82 (one call per image there)
 * Same mathematics as score_pass1 / score_pass2 with impl 1, for n_queries <= sixdgs_score_batch_max() queries that
 * share the key cache: q[n_queries, 256, 384] fp32, part_m / part_z [n_queries, sixdgs_score_batch_parts(), 256],
 * m / z [n_queries, 256], scores[n_queries, score_stride] (score_stride >= n_rays).  The keys cross HBM once per
 * pass for the whole batch.  k_dtype SIXDGS_BF16, SIXDGS_F16X2 or SIXDGS_F16F8; workspace >= sixdgs_score_batch_workspace(n_queries). */
int sixdgs_score_batch_max(void);
int sixdgs_score_batch_parts(void);
size_t sixdgs_score_batch_workspace(int n_queries);
int sixdgs_score_pass1_batch(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries,
                             int n_img, float* part_m, float* part_z, void* workspace, size_t workspace_bytes,
                             void* stream);
int sixdgs_score_pass2_batch(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries,
                             int n_img, const float* m, const float* z, float* scores, int64_t score_stride,
                             void* workspace, size_t workspace_bytes, void* stream);

/* pass 2 with the all-ray weighted least-squares system of least_squared_loss.py:47-64 / line_intersection.py:75-154
 * accumulated in the epilogue (weights = this pass's scores): per query 13 doubles
 *   R = sum w (I - d d^T) as (xx,xy,xz,yy,yz,zz), q = sum w (I - d d^T) o (3), sum w d (3), sum w
 * ls_part [n_queries, sixdgs_ls_partial_rows(), 13] = per-CTA partial sums (scratch), ls_sys [n_queries, 13] = their
 * sum in a fixed order.  Across ray shards the systems simply add: one 13-double all-reduce per query.
 * sixdgs_ls_solve (n systems): centre[n,3] = solve(R, q) after scaling every sum by weight_scale (1 / n_img in the
 * reference); NaN x3 and status |= 1 when det(R) < 1e-7; watch[n,3] (nullable) = normalised sum w d; c2w[n,16]
 * (nullable, needs up[n,3]) = pose from (centre, watch, up) exactly like the tail of the top-k path (test.py:187-198);
 * aux[n,8] (nullable) = centre3, watch3, sum of weights, status. */
int sixdgs_ls_partial_rows(void);
int sixdgs_score_pass2_batch_ls(const void* k_cache, int k_dtype, int64_t n_rays, const float* q, int n_queries,
                                int n_img, const float* m, const float* z, float* scores, int64_t score_stride,
                                const float* rays_ori, const float* rays_dir, double* ls_part, double* ls_sys,
                                void* workspace, size_t workspace_bytes, void* stream);
int sixdgs_ls_solve(const double* ls_sys, int n, double weight_scale, const float* up, float* centre, float* watch,
                    float* c2w, float* aux, int32_t* status, void* stream);

/* ---- a12: top-k -------- identification_module.py:131 (torch.topk, sorted descending) -----------
 * workspace >= sixdgs_topk_workspace(n, k).  idx int64, ties broken by lower index first. */
size_t sixdgs_topk_workspace(int64_t n, int k);
int sixdgs_topk(const float* scores, int64_t n, int k, float* vals, int64_t* idx, void* workspace,
                size_t workspace_bytes, void* stream);

/* ---- a13: least-squares line intersection -------- line_intersection.py:75-154 ------------------
 * R = sum w (I - d d^T), q = sum w (I - d d^T) o, centre = solve(R, q); NaN x3 and *status |= 1 when
 * det(R) < 1e-7.  weights nullable (the live call sites pass none, test.py:169-179).  Any n (the
 * weighted all-ray form is the one least_squared_loss.py:62-64 specifies). */
int sixdgs_line_intersect(const float* points, const float* dirs, const float* weights, int64_t n,
                          float* centre, int32_t* status, void* workspace /* 12 doubles */, void* stream);

/* ---- a14: pose tail -------- pose_estimation/test.py:157-198, line_intersection.py:5-34 --------
 * top-k candidates (idx into rays_ori/rays_dir, vals = scores) + camera up -> c2w[16] row-major.
 * aux (nullable, 8 floats): centre[3], watch[3], n_kept, status (bit0 NaN centre, bit1 singular
 * rotation -> identity, bit2 NaN c2w -> identity).  k <= 1024. */
int sixdgs_pose_tail(const float* rays_ori, const float* rays_dir, int64_t ray_stride /* floats between rays, >= 3 */,
                     const int64_t* idx, const float* vals, int k, const float* up, float* c2w, float* aux,
                     void* stream);

/* ---- multi-GPU candidate exchange (no reference counterpart): out[k,7] = (score, ori3, dir3) of the local
 * top-k_local rays, rows >= k_local padded with score -inf. */
int sixdgs_gather_candidates(const float* vals, const int64_t* idx, int k_local, int k, const float* rays_ori,
                             const float* rays_dir, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIXDGS_H */
