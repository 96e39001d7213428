"""Tensor-level wrappers over the C ABI (one function per kernel family).  torch is used for device
memory and streams only; every computation below happens in libsixdgs.so."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import BF16, F16F8, F16X2, F32, FEAT, MAX_TOKENS, call, dptr, f32c, stream_ptr

SCORE_SIMT = 0
SCORE_TC = 1


def degrade_mask(scaling_raw: torch.Tensor, target_points: int = 50) -> Tuple[torch.Tensor, torch.Tensor]:
    s = f32c(scaling_raw)
    n = s.shape[0]
    valid = torch.empty(n, dtype=torch.uint8, device=s.device)
    rings = torch.empty(n, dtype=torch.int32, device=s.device)
    if n == 0:  # empty tensors have a null data pointer, which the C ABI rejects by design
        return valid.bool(), rings
    call("sixdgs_degrade_mask", dptr(s), n, target_points, dptr(valid, torch.uint8), dptr(rings, torch.int32), stream_ptr())
    return valid.bool(), rings


def _knn_grid_layout(c: torch.Tensor, points_per_cell: float = 8.0):
    """uniform grid over a robust box of the cloud: (lo[3], cell edge, dims[3]); one host read of the statistics.
    The box is the bounding box clipped to the 1 %..99 % quantile range (of a <= 65536-point strided subsample)
    widened by half that range on each side, so a few far outliers (floaters of a real 3DGS scene) cannot inflate the
    cells -- mean/sigma would not do: ten points at 1e6 drag sigma itself.  Points outside the box are clamped into
    border cells and the search stays exact."""
    lo_t, hi_t = torch.aminmax(c, dim=0)
    sub = c[:: max(1, c.shape[0] // 65536)]
    k_lo = max(1, int(0.01 * sub.shape[0]))
    k_hi = min(sub.shape[0], sub.shape[0] - k_lo + 1)
    q_lo = torch.kthvalue(sub, k_lo, dim=0).values
    q_hi = torch.kthvalue(sub, k_hi, dim=0).values
    pad = 0.5 * (q_hi - q_lo)
    box = torch.stack((torch.maximum(lo_t, q_lo - pad), torch.minimum(hi_t, q_hi + pad))).cpu().double()
    lo, hi = box[0], box[1]
    ext = (hi - lo).clamp_min(1e-12)
    m = c.shape[0]
    h = float((ext.prod() / max(m / points_per_cell, 1.0)) ** (1.0 / 3.0))
    h = max(h, float(ext.max()) / 1024.0, 1e-12)
    while True:
        dims = [int(e // h) + 1 for e in ext.tolist()]
        if max(dims) <= 1024 and dims[0] * dims[1] * dims[2] <= max(4 * m, 4096):
            break
        h *= 1.25
    return [float(x) for x in lo.tolist()], h, dims


def knn_normals(cloud: torch.Tensor, k: int = 20, q_begin: int = 0, q_count: Optional[int] = None,
                method: str = "auto") -> torch.Tensor:
    """Exact k-NN normals of rows [q_begin, q_begin+q_count) of `cloud`.  method: "brute" (O(m^2), no setup),
    "grid" (uniform-grid shell search, same result, O(m)) or "auto" (grid from 4096 points up)."""
    c = f32c(cloud)
    m = c.shape[0]
    q_count = m - q_begin if q_count is None else q_count
    out = torch.empty(q_count, 3, dtype=torch.float32, device=c.device)
    if q_count == 0:
        return out
    if method == "brute" or (method == "auto" and m < 4096):
        call("sixdgs_knn_normals", dptr(c), m, q_begin, q_count, k, dptr(out), stream_ptr())
        return out
    lo, h, dims = _knn_grid_layout(c)
    n_cells = dims[0] * dims[1] * dims[2]
    wsz = int(_lib.load().sixdgs_knn_grid_workspace(m, n_cells))
    ws = torch.empty(wsz, dtype=torch.uint8, device=c.device)
    lo_arr = (ctypes.c_float * 3)(*lo)
    dims_arr = (ctypes.c_int * 3)(*dims)
    call("sixdgs_knn_normals_grid", dptr(c), m, q_begin, q_count, k, ctypes.cast(lo_arr, ctypes.c_void_p),
         ctypes.c_float(h), ctypes.cast(dims_arr, ctypes.c_void_p), dptr(out), dptr(ws, torch.uint8), wsz, stream_ptr())
    return out


def sym_eig3x3(A: torch.Tensor, eigenvectors: bool = True, eps: Optional[float] = None):
    a = f32c(A).reshape(-1, 3, 3)
    n = a.shape[0]
    vals = torch.empty(n, 3, dtype=torch.float32, device=a.device)
    vecs = torch.empty(n, 3, 3, dtype=torch.float32, device=a.device) if eigenvectors else None
    if n == 0:
        return vals, vecs
    call("sixdgs_sym_eig3x3", dptr(a), n, ctypes.c_float(eps or 0.0), dptr(vals), dptr(vecs), stream_ptr())
    return vals, vecs


def exclusive_scan(counts: torch.Tensor) -> torch.Tensor:
    n = counts.shape[0]
    out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    call("sixdgs_exclusive_scan", dptr(counts, torch.int32), n, dptr(out, torch.int64), stream_ptr())
    return out


def raygen(xyz, scaling_raw, rotation_raw, features, sh_degree: int, sel: torch.Tensor, normals: Optional[torch.Tensor],
           target_points: int = 50, resolution: int = 1000, mode: int = 0):
    """cells per ellipsoid (cheap) -> scan -> ONE table pass that writes each ellipsoid's rays densely at its cell
    offset -> scan of the surviving counts -> block copy to the gap-free arrays.  Two small device->host reads (total
    cells, total rays); the reference has the same kind of sync at sampling.py:145."""
    dev = scaling_raw.device
    m = sel.shape[0]
    s = stream_ptr()
    cells_per = torch.empty(m, dtype=torch.int32, device=dev)
    call("sixdgs_raygen_cells", dptr(scaling_raw), dptr(sel, torch.int64), m, target_points, dptr(cells_per, torch.int32), s)
    slot = exclusive_scan(cells_per)
    n_slots = int(slot[-1].item())
    rays_per = torch.zeros(m, dtype=torch.int32, device=dev)
    full = mode == 0
    t_ori = torch.empty(max(n_slots, 1), 3, dtype=torch.float32, device=dev)
    t_dir = torch.empty(max(n_slots, 1), 3, dtype=torch.float32, device=dev) if full else None
    t_rgb = torch.empty(max(n_slots, 1), 3, dtype=torch.float32, device=dev) if full else None
    t_ell = torch.empty(max(n_slots, 1), dtype=torch.int64, device=dev)
    sh_coeffs = 16
    if features is not None:
        if features.dim() != 3 or features.shape[2] != 3 or features.shape[1] < (sh_degree + 1) ** 2:
            raise _lib.SixdgsError(f"features must be [N, >= (sh_degree+1)^2, 3] (get_features layout); got "
                                   f"{tuple(features.shape)} for sh_degree {sh_degree}")
        sh_coeffs = int(features.shape[1])
    if n_slots:
        call("sixdgs_raygen_fill", dptr(xyz), dptr(scaling_raw), dptr(rotation_raw), dptr(features), sh_degree, sh_coeffs,
             dptr(sel, torch.int64), m, dptr(normals), target_points, resolution, mode, dptr(slot, torch.int64),
             dptr(t_ori), dptr(t_dir), dptr(t_rgb), dptr(t_ell, torch.int64), dptr(rays_per, torch.int32), s)
    offs = exclusive_scan(rays_per)
    n_rays = int(offs[-1].item())
    ori = torch.empty(n_rays, 3, dtype=torch.float32, device=dev)
    ell = torch.empty(n_rays, dtype=torch.int64, device=dev)
    dirs = torch.empty(n_rays, 3, dtype=torch.float32, device=dev) if full else None
    rgb = torch.empty(n_rays, 3, dtype=torch.float32, device=dev) if full else None
    if n_rays:
        call("sixdgs_raygen_compact", dptr(slot, torch.int64), dptr(offs, torch.int64), dptr(rays_per, torch.int32), m,
             dptr(t_ori), dptr(t_dir), dptr(t_rgb), dptr(t_ell, torch.int64), dptr(ori), dptr(dirs), dptr(rgb),
             dptr(ell, torch.int64), s)
    return ori, dirs, rgb, ell, cells_per


def pack_ray_mlp_weights(sd: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    """Zero-pad the reference state-dict matrices so every reduction dim is a multiple of 32
    (layout documented in include/sixdgs.h)."""
    def g(k):
        return sd[k].detach().to(device=device, dtype=torch.float32)

    w1 = torch.zeros(512, 160, device=device)
    w1[:, :141] = g("ray_preprocessor.mlp.0.weight")
    w3 = torch.zeros(512, 672, device=device)
    w3[:, :653] = g("ray_preprocessor.mlp2.0.weight")
    wq = torch.zeros(FEAT, 400, device=device)
    wq[:, :398] = g("attention.q_proj.weight")

    def tf32(t):
        """round-to-nearest onto the TF32 grid (10 mantissa bits): the tensor-core GEMMs truncate fp32 operands,
        pre-rounding makes that truncation exact (features_tc.cu)"""
        i = t.contiguous().view(torch.int32)
        return ((i + 0x1000) & -8192).view(torch.float32)

    pw = {
        "w1p": w1.contiguous(), "b1": g("ray_preprocessor.mlp.0.bias").contiguous(),
        "w2": g("ray_preprocessor.mlp.2.weight").contiguous(), "b2": g("ray_preprocessor.mlp.2.bias").contiguous(),
        "w3p": w3.contiguous(), "b3": g("ray_preprocessor.mlp2.0.bias").contiguous(),
        "w4": g("ray_preprocessor.mlp2.2.weight").contiguous(), "b4": g("ray_preprocessor.mlp2.2.bias").contiguous(),
        "wk": g("attention.k_proj.weight").contiguous(), "bk": g("attention.k_proj.bias").contiguous(),
        "wqp": wq.contiguous(), "bq": g("attention.q_proj.bias").contiguous(),
    }
    for k in ("w1p", "w2", "w3p", "w4", "wk"):
        pw[k + "_tf32"] = tf32(pw[k])

    def x2(t, kpad):
        """[n,k] fp32 -> [n, 2*kpad] fp16 = [hi | lo] (zero padded): the operand format of the exact tensor-core build"""
        wpad = torch.zeros(t.shape[0], kpad, device=device)
        wpad[:, :t.shape[1]] = t
        hi = wpad.to(torch.float16)
        lo = (wpad - hi.float()).to(torch.float16)
        return torch.cat((hi, lo), 1).contiguous()

    w3x = torch.zeros(512, 704, device=device)  # [h 512 | x 141 -> 192]
    w3x[:, :653] = g("ray_preprocessor.mlp2.0.weight")
    pw["w1_x2"] = x2(g("ray_preprocessor.mlp.0.weight"), 192)
    pw["w2_x2"] = x2(pw["w2"], 512)
    pw["w3_x2"] = x2(w3x, 704)
    pw["w4_x2"] = x2(pw["w4"], 512)
    pw["wk_x2"] = x2(pw["wk"], 384)
    return pw


FEATURES_SIMT = 0  # fp32 FMA GEMMs (exact)
FEATURES_TC = 1    # TF32 tcgen05 GEMMs, shared-memory staged TMA-store epilogue
FEATURES_TC_DIRECT = 3  # same GEMMs, direct-store epilogue (bit-identical, 8 % slower; the comparison kernel)


def ray_features(ori, dirs, rgb, pw: Dict[str, torch.Tensor], k_dtype: Optional[int] = F32, want_features: bool = False,
                 project: bool = True, impl: int = FEATURES_SIMT):
    """-> (K cache [n,384] in k_dtype or None, features [n,384] fp32 or None)."""
    ori, dirs, rgb = f32c(ori), f32c(dirs), f32c(rgb)
    n = ori.shape[0]
    dev = ori.device
    k_out = None
    if k_dtype is not None:
        k_out = torch.empty(n, FEAT, dtype=torch.float32 if k_dtype == F32 else torch.bfloat16, device=dev)
    feat = torch.empty(n, FEAT, dtype=torch.float32, device=dev) if want_features else None
    if n == 0:
        return k_out, feat
    wsz = int(_lib.load().sixdgs_ray_features_workspace(n))
    ws = torch.empty(wsz, dtype=torch.uint8, device=dev)
    sfx = "_tf32" if impl in (FEATURES_TC, FEATURES_TC_DIRECT) else ""
    call("sixdgs_ray_features", dptr(ori), dptr(dirs), dptr(rgb), n, dptr(pw["w1p" + sfx]), dptr(pw["b1"]),
         dptr(pw["w2" + sfx]), dptr(pw["b2"]), dptr(pw["w3p" + sfx]), dptr(pw["b3"]), dptr(pw["w4" + sfx]), dptr(pw["b4"]),
         dptr(pw["wk" + sfx]) if project else None, dptr(pw["bk"]) if project else None,
         dptr(k_out, None), k_dtype if k_dtype is not None else F32, dptr(feat), impl, dptr(ws, torch.uint8), wsz,
         stream_ptr())
    return k_out, feat


def ray_features_x2(ori, dirs, rgb, pw: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None,
                    absmax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rays -> f16x2 key cache [n,768] on the exact tensor-core path (three-term split-fp16 GEMMs, csrc/features_x2.cu)"""
    ori, dirs, rgb = f32c(ori), f32c(dirs), f32c(rgb)
    n = ori.shape[0]
    if out is None:
        out = torch.empty(n, 2 * FEAT, dtype=torch.float16, device=ori.device)
    if n == 0:
        return out
    wsz = int(_lib.load().sixdgs_ray_features_x2_workspace(n))
    ws = torch.empty(wsz, dtype=torch.uint8, device=ori.device)
    h = torch.float16
    call("sixdgs_ray_features_x2", dptr(ori), dptr(dirs), dptr(rgb), n, dptr(pw["w1_x2"], h), dptr(pw["b1"]),
         dptr(pw["w2_x2"], h), dptr(pw["b2"]), dptr(pw["w3_x2"], h), dptr(pw["b3"]), dptr(pw["w4_x2"], h), dptr(pw["b4"]),
         dptr(pw["wk_x2"], h), dptr(pw["bk"]), dptr(out, h), dptr(absmax), dptr(ws, torch.uint8), wsz, stream_ptr())
    return out


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool = False) -> torch.Tensor:
    """y = act(x w^T + b); x [m,k] with k % 16 == 0 (pad on the caller side)."""
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty(m, n, dtype=torch.float32, device=x.device)
    call("sixdgs_linear", dptr(x), m, k, k, dptr(w), dptr(b), n, dptr(y), n, int(relu), stream_ptr())
    return y


def project_queries(img_fea: torch.Tensor, pw: Dict[str, torch.Tensor]) -> torch.Tensor:
    """q = Wq [f_img | pe] + bq (our_multihead_attention.py:74); 398 -> 400 zero pad."""
    n_img = img_fea.shape[0]
    x = torch.zeros(n_img, 400, dtype=torch.float32, device=img_fea.device)
    x[:, :398] = img_fea
    return linear(x, pw["wqp"], pw["bq"])


def _kdtype(k: torch.Tensor) -> int:
    if k.dtype == torch.float32 and k.shape[-1] == FEAT:
        return F32
    if k.dtype == torch.bfloat16 and k.shape[-1] == FEAT:
        return BF16
    if k.dtype == torch.float16 and k.shape[-1] == 2 * FEAT:
        return F16X2
    if k.dtype == torch.uint8 and k.shape[-1] == 4 * FEAT:
        return F16F8
    raise _lib.SixdgsError(f"key cache must be float32 / bfloat16 [n,{FEAT}], float16 [n,{2 * FEAT}] (f16x2) or uint8 "
                           f"[n,{4 * FEAT}] (f16f8), "
                           f"got {k.dtype} {tuple(k.shape)}")


def keys_to_f16f8(keys_f16x2: torch.Tensor) -> torch.Tensor:
    """f16x2 cache [n,768] fp16 -> f16f8 cache [n,1536] uint8, IN PLACE (the returned tensor aliases the input, whose
    lo half is overwritten by the two e4m3 planes)"""
    if _kdtype(keys_f16x2) != F16X2 or not keys_f16x2.is_contiguous():
        raise _lib.SixdgsError("keys_to_f16f8 needs a contiguous f16x2 cache")
    n = keys_f16x2.shape[0]
    if n:
        call("sixdgs_keys_f16x2_to_f16f8", dptr(keys_f16x2, torch.float16), n, stream_ptr())
    return keys_f16x2.view(torch.uint8)


F16X2_KEY_SCALE = 16.0   # stored f16x2 key = 16 k  (csrc/score_tc_mq.cu)
F16_MAX = 65504.0


def split_keys(k_f32: torch.Tensor, out: Optional[torch.Tensor] = None, absmax: Optional[torch.Tensor] = None):
    """fp32 keys [n,384] -> exact tensor-core format [n,768] fp16 = [hi | lo] of 16 k.  `absmax` (device float[1],
    zero-initialised by the caller) accumulates max |16 k| so the caller can check the fp16 range once at the end."""
    k = f32c(k_f32)
    n = k.shape[0]
    if out is None:
        out = torch.empty(n, 2 * FEAT, dtype=torch.float16, device=k.device)
    if n:
        call("sixdgs_split_keys", dptr(k), n, dptr(out, torch.float16), dptr(absmax), stream_ptr())
    return out


_score_ws = {}


def _score_workspace(impl: int, device) -> torch.Tensor:
    """per-(impl, device, stream) scratch for the score kernels (impl 1: bf16 copy of q for the TMA)."""
    key = (impl, str(device), stream_ptr())
    if key not in _score_ws:
        n = int(_lib.load().sixdgs_score_workspace(impl))
        _score_ws[key] = torch.empty(max(n, 16), dtype=torch.uint8, device=device)
    return _score_ws[key]


def score_pass1(k_cache: torch.Tensor, q: torch.Tensor, impl: int = SCORE_SIMT):
    parts = int(_lib.load().sixdgs_score_parts(impl))
    ws = _score_workspace(impl, q.device)
    pm = torch.empty(parts, MAX_TOKENS, dtype=torch.float32, device=q.device)
    pz = torch.empty(parts, MAX_TOKENS, dtype=torch.float32, device=q.device)
    call("sixdgs_score_pass1", dptr(k_cache, None), _kdtype(k_cache), k_cache.shape[0], dptr(q), q.shape[0], dptr(pm),
         dptr(pz), impl, dptr(ws, torch.uint8), ws.numel(), stream_ptr())
    return pm, pz


def score_merge(pm: torch.Tensor, pz: torch.Tensor, n_img: int, token_valid: Optional[torch.Tensor] = None,
                rows: Optional[int] = None, groups: int = 1, group_stride: int = 0, first_row: int = 0):
    """merge `groups` groups of `rows` consecutive partial rows (group g starts at first_row + g*group_stride);
    default: all rows of pm/pz."""
    m = torch.empty(MAX_TOKENS, dtype=torch.float32, device=pm.device)
    z = torch.empty(MAX_TOKENS, dtype=torch.float32, device=pm.device)
    rows = pm.shape[0] if rows is None else rows
    off = first_row * MAX_TOKENS * 4
    call("sixdgs_score_merge", dptr(pm) + off, dptr(pz) + off, rows, groups, group_stride, n_img,
         dptr(token_valid, torch.uint8), dptr(m), dptr(z), stream_ptr())
    return m, z


def score_pass2(k_cache: torch.Tensor, q: torch.Tensor, m: torch.Tensor, z: torch.Tensor, impl: int = SCORE_SIMT,
                want_map: bool = False, out: Optional[torch.Tensor] = None):
    n = k_cache.shape[0]
    scores = out if out is not None else torch.empty(n, dtype=torch.float32, device=q.device)
    amap = torch.empty(q.shape[0], n, dtype=torch.float32, device=q.device) if want_map else None
    ws = _score_workspace(impl, q.device)
    call("sixdgs_score_pass2", dptr(k_cache, None), _kdtype(k_cache), n, dptr(q), q.shape[0], dptr(m), dptr(z),
         dptr(scores), dptr(amap), impl, dptr(ws, torch.uint8), ws.numel(), stream_ptr())
    return scores, amap


# ---- several queries per key sweep (csrc/score_tc_mq.cu; bf16, f16x2 and f16f8 key caches) --------------------
def score_batch_max() -> int:
    return int(_lib.load().sixdgs_score_batch_max())


def _score_batch_workspace(device) -> torch.Tensor:
    key = ("mq", str(device), stream_ptr())
    if key not in _score_ws:
        n = int(_lib.load().sixdgs_score_batch_workspace(score_batch_max()))
        _score_ws[key] = torch.empty(n, dtype=torch.uint8, device=device)
    return _score_ws[key]


def _check_batch_q(q: torch.Tensor):
    if q.dim() != 3 or q.shape[1] != MAX_TOKENS or q.shape[2] != FEAT:
        raise _lib.SixdgsError(f"batched scoring needs q [B, {MAX_TOKENS}, {FEAT}], got {tuple(q.shape)}")


def score_pass1_batch(k_cache: torch.Tensor, q: torch.Tensor, n_img: int = MAX_TOKENS):
    """q [B,256,384] -> partial rows pm, pz [B * parts, 256] (query-major, the layout of B score_pass1 calls)"""
    _check_batch_q(q)
    lib = _lib.load()
    nb, parts, step = q.shape[0], int(lib.sixdgs_score_batch_parts()), score_batch_max()
    ws = _score_batch_workspace(q.device)
    pm = torch.empty(nb * parts, MAX_TOKENS, dtype=torch.float32, device=q.device)
    pz = torch.empty(nb * parts, MAX_TOKENS, dtype=torch.float32, device=q.device)
    for b0 in range(0, nb, step):
        nq = min(step, nb - b0)
        call("sixdgs_score_pass1_batch", dptr(k_cache, None), _kdtype(k_cache), k_cache.shape[0], dptr(q[b0:b0 + nq]), nq,
             n_img, dptr(pm[b0 * parts:(b0 + nq) * parts]), dptr(pz[b0 * parts:(b0 + nq) * parts]), dptr(ws, torch.uint8),
             ws.numel(), stream_ptr())
    return pm, pz


def score_pass2_batch(k_cache: torch.Tensor, q: torch.Tensor, m: torch.Tensor, z: torch.Tensor, n_img: int = MAX_TOKENS,
                      out: Optional[torch.Tensor] = None, ls_rays=None):
    """q [B,256,384], m / z [B,256] -> scores [B, n_rays].
    ls_rays = (rays_ori, rays_dir): additionally accumulate, in the same epilogue, the all-ray weighted least-squares
    system of least_squared_loss.py:62-64 with weights = scores -> returns (scores, ls_sys [B,13] float64)."""
    _check_batch_q(q)
    nb, n, step = q.shape[0], k_cache.shape[0], score_batch_max()
    scores = out if out is not None else torch.empty(nb, n, dtype=torch.float32, device=q.device)
    ws = _score_batch_workspace(q.device)
    ls_sys = ls_part = None
    if ls_rays is not None:
        rows = int(_lib.load().sixdgs_ls_partial_rows())
        ls_part = torch.empty(step, rows, 13, dtype=torch.float64, device=q.device)
        ls_sys = torch.empty(nb, 13, dtype=torch.float64, device=q.device)
    for b0 in range(0, nb, step):
        nq = min(step, nb - b0)
        if ls_rays is None:
            call("sixdgs_score_pass2_batch", dptr(k_cache, None), _kdtype(k_cache), n, dptr(q[b0:b0 + nq]), nq, n_img,
                 dptr(m[b0:b0 + nq]), dptr(z[b0:b0 + nq]), dptr(scores[b0:b0 + nq]), scores.stride(0),
                 dptr(ws, torch.uint8), ws.numel(), stream_ptr())
        else:
            call("sixdgs_score_pass2_batch_ls", dptr(k_cache, None), _kdtype(k_cache), n, dptr(q[b0:b0 + nq]), nq, n_img,
                 dptr(m[b0:b0 + nq]), dptr(z[b0:b0 + nq]), dptr(scores[b0:b0 + nq]), scores.stride(0),
                 dptr(f32c(ls_rays[0])), dptr(f32c(ls_rays[1])), dptr(ls_part, torch.float64),
                 dptr(ls_sys[b0:b0 + nq], torch.float64), dptr(ws, torch.uint8), ws.numel(), stream_ptr())
    return scores if ls_rays is None else (scores, ls_sys)


def ls_solve(ls_sys: torch.Tensor, weight_scale: float = 1.0, up: Optional[torch.Tensor] = None):
    """ls_sys [B,13] float64 (summed over shards) -> (centre [B,3], watch [B,3], status [B] int32); the sums are scaled
    by weight_scale first (1 / n_img: weights = score / n_img, least_squared_loss.py:62-64).
    With up [B,3] (unit camera-up vectors): -> (c2w [B,4,4], aux [B,8]) like the tail of the top-k path."""
    nb = ls_sys.shape[0]
    dev = ls_sys.device
    if up is not None:
        c2w = torch.empty(nb, 4, 4, dtype=torch.float32, device=dev)
        aux = torch.empty(nb, 8, dtype=torch.float32, device=dev)
        call("sixdgs_ls_solve", dptr(ls_sys, torch.float64), nb, ctypes.c_double(weight_scale), dptr(f32c(up)), None, None,
             dptr(c2w), dptr(aux), None, stream_ptr())
        return c2w, aux
    centre = torch.empty(nb, 3, dtype=torch.float32, device=dev)
    watch = torch.empty(nb, 3, dtype=torch.float32, device=dev)
    status = torch.zeros(nb, dtype=torch.int32, device=dev)
    call("sixdgs_ls_solve", dptr(ls_sys, torch.float64), nb, ctypes.c_double(weight_scale), None, dptr(centre), dptr(watch),
         None, None, dptr(status, torch.int32), stream_ptr())
    return centre, watch, status


def score_backward(k_f32: torch.Tensor, q: torch.Tensor, m: torch.Tensor, z: torch.Tensor, grad_scores: torch.Tensor,
                   chunk: int = 1 << 18):
    """Backward of scores = sum_i softmax_r(q k^T / sqrt(384)) w.r.t. q [n_img,384] and k [n,384] (fp32), given
    dLoss/dscores [n] and the forward statistics (m, z): two streaming passes over the keys (csrc/score_simt.cu:
    score_bwd_kernel) + two GEMMs per chunk of rays (sixdgs_linear).  -> (dq, dk)"""
    k, q, g = f32c(k_f32), f32c(q), f32c(grad_scores)
    n, n_img = k.shape[0], q.shape[0]
    dev = k.device
    parts = int(_lib.load().sixdgs_score_backward_parts())
    part = torch.empty(parts, MAX_TOKENS, dtype=torch.float32, device=dev)
    gbar = torch.empty(MAX_TOKENS, dtype=torch.float32, device=dev)
    call("sixdgs_score_backward_gbar", dptr(k), n, dptr(q), n_img, dptr(m), dptr(z), dptr(g), dptr(part), dptr(gbar),
         stream_ptr())
    qt = torch.zeros(FEAT, MAX_TOKENS, dtype=torch.float32, device=dev)  # q^T, token axis padded to 256
    qt[:, :n_img] = q.t()
    dq = torch.zeros(MAX_TOKENS, FEAT, dtype=torch.float32, device=dev)
    dk = torch.empty(n, FEAT, dtype=torch.float32, device=dev)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        c = hi - lo
        ldt = -(-c // 16) * 16
        dl = torch.empty(c, MAX_TOKENS, dtype=torch.float32, device=dev)
        dlt = torch.zeros(MAX_TOKENS, ldt, dtype=torch.float32, device=dev)
        call("sixdgs_score_backward_dlogits", dptr(k[lo:hi]), c, dptr(q), n_img, dptr(m), dptr(z), dptr(g[lo:hi]), dptr(gbar),
             dptr(dl), dptr(dlt), ldt, stream_ptr())
        call("sixdgs_linear", dptr(dl), c, MAX_TOKENS, MAX_TOKENS, dptr(qt), None, FEAT, dptr(dk[lo:hi]), FEAT, 0,
             stream_ptr())                                     # dk = dlogits q
        kt = torch.zeros(FEAT, ldt, dtype=torch.float32, device=dev)
        kt[:, :c] = k[lo:hi].t()
        dq += linear(dlt, kt, None)                            # dq += dlogits^T k
    return dq[:n_img].contiguous(), dk


def topk(scores: torch.Tensor, k: int):
    n = scores.shape[0]
    vals = torch.empty(k, dtype=torch.float32, device=scores.device)
    idx = torch.empty(k, dtype=torch.int64, device=scores.device)
    wsz = int(_lib.load().sixdgs_topk_workspace(n, k))
    ws = torch.empty(wsz, dtype=torch.uint8, device=scores.device)
    call("sixdgs_topk", dptr(scores), n, k, dptr(vals), dptr(idx, torch.int64), dptr(ws, torch.uint8), wsz, stream_ptr())
    return vals, idx


def line_intersect(points: torch.Tensor, dirs: torch.Tensor, weights: Optional[torch.Tensor] = None):
    p, d = f32c(points), f32c(dirs)
    w = f32c(weights) if weights is not None else None
    if p.shape[0] == 0:  # no rays: R = 0, det < 1e-7 -> the reference's NaN vector (line_intersection.py:139-142)
        return torch.full((3,), float("nan"), device=p.device), torch.ones(1, dtype=torch.int32, device=p.device)
    centre = torch.empty(3, dtype=torch.float32, device=p.device)
    status = torch.zeros(1, dtype=torch.int32, device=p.device)
    ws = torch.empty(12, dtype=torch.float64, device=p.device)
    call("sixdgs_line_intersect", dptr(p), dptr(d), dptr(w), p.shape[0], dptr(centre), dptr(status, torch.int32),
         dptr(ws, torch.float64), stream_ptr())
    return centre, status


def pose_tail(rays_ori, rays_dir, idx, vals, up):
    c2w = torch.empty(4, 4, dtype=torch.float32, device=rays_ori.device)
    aux = torch.empty(8, dtype=torch.float32, device=rays_ori.device)
    call("sixdgs_pose_tail", dptr(rays_ori), dptr(rays_dir), 3, dptr(idx, torch.int64), dptr(vals), idx.shape[0],
         dptr(f32c(up)), dptr(c2w), dptr(aux), stream_ptr())
    return c2w, aux


def pose_tail_candidates(cand: torch.Tensor, idx, vals, up):
    """pose tail straight from a packed candidate table [n,7] = (score, ori3, dir3) (no slicing copies)"""
    c2w = torch.empty(4, 4, dtype=torch.float32, device=cand.device)
    aux = torch.empty(8, dtype=torch.float32, device=cand.device)
    base = dptr(cand)
    call("sixdgs_pose_tail", base + 4, base + 16, 7, dptr(idx, torch.int64), dptr(vals), idx.shape[0], dptr(f32c(up)),
         dptr(c2w), dptr(aux), stream_ptr())
    return c2w, aux


def gather_candidates(vals, idx, rays_ori, rays_dir, k: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = out if out is not None else torch.empty(k, 7, dtype=torch.float32, device=rays_ori.device)
    call("sixdgs_gather_candidates", dptr(vals), dptr(idx, torch.int64), idx.shape[0], k, dptr(rays_ori), dptr(rays_dir),
         dptr(out), stream_ptr())
    return out
