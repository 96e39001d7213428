"""``IdentificationModule`` with the reference API (pose_estimation/identification_module.py:10-133)
on top of the sm_100a kernels.

Same sub-module and parameter names as the reference, so ``id_module.th`` state dicts load
unchanged (``ray_preprocessor.mlp.{0,2}``, ``ray_preprocessor.mlp2.{0,2}``, ``attention.{q,k}_proj``,
``camera_direction_prediction_network.*``, ``backbone_wrapper.*``).

What changes underneath: the ray-feature MLP and the key projection do not depend on the
query image (identification_module.py:79-80) but the reference recomputes them for every query.
Here they are computed once per (rays, weights) pair into a key cache K[n_rays, 384] (fp32 or bf16);
a query is then two streaming passes over K (per-token softmax statistics, then scores), a radix
top-k and one fused pose-tail launch.  The [n_img, n_rays] attention map is only materialised when it
is small (``attention_map_bytes_limit``); a 1M-Gaussian scene would need 30 GB.

Training (SURVEY §8b "Autograd", §8f-3): when gradients are required (``torch.is_grad_enabled()`` and trainable
hot-path parameters) ``run_attention`` / ``forward`` take ``_run_attention_autograd``: the ray MLP and the two
projections are differentiable torch ops (autograd owns their activations), while the softmax-over-rays score runs
forward AND backward on the kernels (``_RayScoreFunction``: two streaming passes each way, no [n_img, n_rays] map saved),
so the reference's ``train_id_module`` keeps working.  This is the training-mode implementation, not a fallback for a
missing extension -- it refuses CPU tensors and requires libsixdgs.so.  The query path (``test_image`` /
``query_pose`` / ``ShardedPoseEstimator``) never takes it.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops, train_ops
from ._lib import BF16, F16F8, F16X2, F32, SixdgsError
from .camera_up import CameraDirectionPredictor
from .image_tokens import BackboneWrapper


class RayPreprocessor(torch.nn.Module):
    """Parameter container + kernel call for the ray MLP (reference ray_preprocessor.py:11-46)."""

    def __init__(self, viewpe: int = 8, pospe: int = 8, rgbpe: int = 6, featureC: int = 128, fea_output: int = 128):
        """signature and defaults of the reference (ray_preprocessor.py:12-19).  The kernels are specialised for the one
        configuration the reference ever builds (identification_module.py:16-18: width 512, output 384), so the
        reference's own defaults (128, 128) are refused rather than silently widened."""
        super().__init__()
        if (viewpe, pospe, rgbpe, featureC, fea_output) != (8, 8, 6, 512, 384):
            raise NotImplementedError("the kernels are specialised for the configuration IdentificationModule builds "
                                      "(PE 8/8/6, featureC=512, fea_output=384; identification_module.py:16-18)")
        self.in_mlpC = 2 * viewpe * 3 + 3 + 2 * pospe * 3 + 3 + 2 * rgbpe * 3 + 3
        relu = lambda: torch.nn.ReLU(inplace=True)  # noqa: E731
        self.mlp = torch.nn.Sequential(torch.nn.Linear(self.in_mlpC, featureC), relu(),
                                       torch.nn.Linear(featureC, featureC), relu())
        self.mlp2 = torch.nn.Sequential(torch.nn.Linear(featureC + self.in_mlpC, featureC), relu(),
                                        torch.nn.Linear(featureC, fea_output))
        self.viewpe, self.pospe, self.rgbpe = viewpe, pospe, rgbpe

    def _packed(self):
        sd = {"ray_preprocessor." + k: v for k, v in self.state_dict().items()}
        dev = self.mlp[0].weight.device
        # identity q/k projections are never used by this module on its own
        sd["attention.q_proj.weight"] = torch.zeros(384, 398, device=dev)
        sd["attention.q_proj.bias"] = torch.zeros(384, device=dev)
        sd["attention.k_proj.weight"] = torch.zeros(384, 384, device=dev)
        sd["attention.k_proj.bias"] = torch.zeros(384, device=dev)
        return ops.pack_ray_mlp_weights(sd, dev)

    @torch.no_grad()
    def forward(self, pts, viewdirs, rgb):
        _, feat = ops.ray_features(pts, viewdirs, rgb, self._packed(), k_dtype=None, want_features=True, project=False)
        return feat


class MultiHeadAttention(torch.nn.Module):
    """1-head attention WEIGHTS only (no V): softmax over rays of (Wq img)(Wk ray)^T / sqrt(384)
    (reference our_multihead_attention.py:45-79)."""

    def __init__(self, ray_fea_size: int, img_fea_size: int, embed_dim: int, num_heads: int = 1):
        super().__init__()
        assert embed_dim % num_heads == 0, "Embedding dimension must be 0 modulo number of heads."
        if (ray_fea_size, img_fea_size, embed_dim, num_heads) != (384, 398, 384, 1):
            raise NotImplementedError("kernels are specialised for 384-d single-head attention")
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.q_proj = torch.nn.Linear(img_fea_size, embed_dim)
        self.k_proj = torch.nn.Linear(ray_fea_size, embed_dim)
        torch.nn.init.xavier_uniform_(self.q_proj.weight)
        self.q_proj.bias.data.fill_(0)
        torch.nn.init.xavier_uniform_(self.k_proj.weight)
        self.k_proj.bias.data.fill_(0)

    @torch.no_grad()
    def forward(self, img_features, ray_features, mask=None):
        if mask is not None:
            raise NotImplementedError("attention masks are never passed on the pose path (identification_module.py:80)")
        dev = img_features.device
        x = torch.zeros(img_features.shape[0], 400, device=dev)
        x[:, :398] = img_features
        wq = torch.zeros(384, 400, device=dev)
        wq[:, :398] = self.q_proj.weight
        q = ops.linear(x, wq, self.q_proj.bias.detach().contiguous())
        k = ops.linear(ray_features.contiguous(), self.k_proj.weight.detach().contiguous(),
                       self.k_proj.bias.detach().contiguous())
        pm, pz = ops.score_pass1(k, q, ops.SCORE_SIMT)
        m, z = ops.score_merge(pm, pz, q.shape[0])
        _, amap = ops.score_pass2(k, q, m, z, ops.SCORE_SIMT, want_map=True)
        return amap


class _RayScoreFunction(torch.autograd.Function):
    """scores = sum_i softmax_r(q k^T / sqrt(384)) with the forward AND the backward on the libsixdgs kernels (exact fp32
    path): forward = the two streaming passes of the query path, backward = two more passes that form
    dlogits = A (g - gbar) / sqrt(384) plus the GEMMs dk = dlogits q, dq = dlogits^T k (ops.score_backward).  The
    [n_img, n_rays] attention map is returned (non-differentiable) only when asked -- the reference's training loop
    reads nothing but its shape (train.py:159)."""

    @staticmethod
    def forward(ctx, q, k, want_map):
        qd, kd = q.detach().float().contiguous(), k.detach().float().contiguous()
        pm, pz = ops.score_pass1(kd, qd, ops.SCORE_SIMT)
        m, z = ops.score_merge(pm, pz, qd.shape[0])
        scores, amap = ops.score_pass2(kd, qd, m, z, ops.SCORE_SIMT, want_map=want_map)
        ctx.save_for_backward(qd, kd, m, z)
        if amap is None:
            amap = torch.empty(qd.shape[0], 0, device=qd.device)
        ctx.mark_non_differentiable(amap)
        return scores, amap

    @staticmethod
    def backward(ctx, g, _g_map):
        qd, kd, m, z = ctx.saved_tensors
        dq, dk = ops.score_backward(kd, qd, m, z, g)
        return dq, dk, None


@dataclass
class RayKeyCache:
    """K[n_rays, 384] (fp32 / bf16) or K[n_rays, 768] (f16x2: fp16 hi | lo) for one (ray set, weight version); the
    per-scene state of a query.  `rays` keeps the three ray tensors alive: the implicit cache of
    IdentificationModule compares them by identity, so a freed tensor's recycled address can never alias it."""
    keys: torch.Tensor
    n_rays: int
    tag: tuple
    scores: Optional[torch.Tensor] = None  # reusable output buffer
    rays: Optional[tuple] = None


class IdentificationModule(torch.nn.Module):
    def __init__(self, backbone_type: str = "dino", camera_up_output_augmentation=None,
                 target_rays_dirs: Optional[torch.Tensor] = None, augmentation_channels: int = 10, *,
                 backbone: Optional[torch.nn.Module] = None, score_impl: Optional[str] = None,
                 attention_map_bytes_limit: int = 1 << 31):
        super().__init__()
        aug = getattr(camera_up_output_augmentation, "name", camera_up_output_augmentation)
        if aug not in (None, "NONE"):
            raise NotImplementedError("camera-up output augmentations are unused by the entry points "
                                      "(pretrain_eval_attention.py:55-59 passes the default)")
        self.backbone_wrapper = BackboneWrapper(backbone_type=backbone_type, backbone=backbone)
        self.ray_preprocessor = RayPreprocessor(featureC=512, fea_output=self.backbone_wrapper.img_num_features)
        self.camera_up_out_augmentation = None
        self.camera_direction_prediction_network = CameraDirectionPredictor(
            self.backbone_wrapper.img_num_features, self.backbone_wrapper.backbone_wh, fea_output=3)
        self.attention = MultiHeadAttention(self.backbone_wrapper.img_num_features,
                                            self.backbone_wrapper.img_num_features + 14,
                                            self.backbone_wrapper.img_num_features, 1)
        self.score_impl = score_impl or os.environ.get("SIXDGS_SCORE_IMPL", "simt_fp32")
        if self.score_impl not in ("simt_fp32", "simt_bf16", "tc_bf16", "tc_f16x2", "tc_f16f8"):
            raise ValueError("score_impl must be simt_fp32 | simt_bf16 | tc_bf16 | tc_f16x2 | tc_f16f8")
        self.attention_map_bytes_limit = attention_map_bytes_limit
        self.features_impl = os.environ.get("SIXDGS_FEATURES_IMPL", "auto")  # auto | simt | direct (TF32, direct-store epilogue)
        self._packed_cache = None
        self._key_cache: Optional[RayKeyCache] = None

    # ------------------------------------------------------------------ scene-side (per ray set)
    def _hot_params(self):
        return [self.ray_preprocessor.mlp[0], self.ray_preprocessor.mlp[2], self.ray_preprocessor.mlp2[0],
                self.ray_preprocessor.mlp2[2], self.attention.q_proj, self.attention.k_proj]

    def _weights_tag(self):
        return tuple((p.data_ptr(), p._version) for lin in self._hot_params() for p in (lin.weight, lin.bias))

    def packed_weights(self):
        tag = self._weights_tag()
        if self._packed_cache is None or self._packed_cache[0] != tag:
            sd = {k: v for k, v in self.state_dict().items() if k.startswith(("ray_preprocessor.", "attention."))}
            self._packed_cache = (tag, ops.pack_ray_mlp_weights(sd, self.attention.q_proj.weight.device))
        return self._packed_cache[1]

    @property
    def _impl(self) -> int:
        return ops.SCORE_TC if self.score_impl in ("tc_bf16", "tc_f16x2", "tc_f16f8") else ops.SCORE_SIMT

    @property
    def _k_dtype(self) -> int:
        return {"simt_fp32": F32, "tc_f16x2": F16X2, "tc_f16f8": F16F8}.get(self.score_impl, BF16)

    SPLIT_CHUNK = 1 << 20  # rays per fp32 staging chunk of the f16x2 build (1.6 GB of fp32 keys)

    @torch.no_grad()
    def build_key_cache(self, rays_ori, rays_dir, rays_rgb) -> RayKeyCache:
        """rays -> PE -> MLP -> k_proj -> K (once per scene / weight update; the reference redoes this per query)."""
        if self.score_impl in ("tc_f16x2", "tc_f16f8"):
            # exact mode (the TF32 build is only good to 1.5e-3 on the keys): three-term split-fp16 tensor-core GEMMs
            # (csrc/features_x2.cu); SIXDGS_FEATURES_IMPL=simt keeps the fp32 FMA GEMMs, staged through fp32 chunks and
            # split into fp16 hi | lo rows (the cross-check of the tensor-core build)
            n = rays_ori.shape[0]
            keys = torch.empty(n, 2 * 384, dtype=torch.float16, device=rays_ori.device)
            absmax = torch.zeros(1, dtype=torch.float32, device=rays_ori.device)
            pw = self.packed_weights()
            if self.features_impl != "simt":
                ops.ray_features_x2(rays_ori, rays_dir, rays_rgb, pw, keys, absmax)
            else:
                for lo in range(0, n, self.SPLIT_CHUNK):
                    hi = min(n, lo + self.SPLIT_CHUNK)
                    kf, _ = ops.ray_features(rays_ori[lo:hi], rays_dir[lo:hi], rays_rgb[lo:hi], pw, k_dtype=F32,
                                             impl=ops.FEATURES_SIMT)
                    ops.split_keys(kf, keys[lo:hi], absmax)
            if n and not float(absmax.item()) < ops.F16_MAX:  # one host read per scene build
                raise SixdgsError(f"f16x2 key cache: max |16 k| = {float(absmax.item()):.4g} exceeds the fp16 range; "
                                  "use score_impl='simt_fp32' for keys of this magnitude")
            if self.score_impl == "tc_f16f8":  # fast variant: e4m3 copies for the two cross terms, converted in place
                if n and not float(absmax.item()) / 64.0 < 448.0:
                    raise SixdgsError(f"f16f8 key cache: max |k| / 4 = {float(absmax.item()) / 64.0:.4g} exceeds the e4m3 "
                                      "range; use score_impl='tc_f16x2'")
                keys = ops.keys_to_f16f8(keys)
            return RayKeyCache(keys, n, ())
        # the throughput (bf16-key) modes build the cache with TF32 tensor-core GEMMs; the exact mode keeps fp32 FMA
        impl = ops.FEATURES_TC if (self.score_impl == "tc_bf16" and self.features_impl != "simt") else ops.FEATURES_SIMT
        if impl == ops.FEATURES_TC and self.features_impl == "direct":
            impl = ops.FEATURES_TC_DIRECT
        keys, _ = ops.ray_features(rays_ori, rays_dir, rays_rgb, self.packed_weights(), k_dtype=self._k_dtype, impl=impl)
        return RayKeyCache(keys, keys.shape[0], ())

    def invalidate_key_cache(self):
        """drop the implicit key cache (call after rewriting a ray buffer in place through raw pointers, which does
        not bump the tensor version the cache watches)"""
        self._key_cache = None

    def _cache_for(self, rays_ori, rays_dir, rays_rgb) -> RayKeyCache:
        """implicit per-module cache: valid while the caller passes the SAME three tensor objects (identity, not
        address -- the cache holds references, so their storage cannot be freed and recycled under it), unmodified
        (tensor version counters) and the hot-path weights are unchanged."""
        tag = (rays_ori._version, rays_dir._version, rays_rgb._version, self.score_impl, self._weights_tag())
        kc = self._key_cache
        if (kc is None or kc.rays is None or kc.rays[0] is not rays_ori or kc.rays[1] is not rays_dir
                or kc.rays[2] is not rays_rgb or kc.tag != tag):
            self._key_cache = None  # release the old keys before building the new ones
            kc = self.build_key_cache(rays_ori, rays_dir, rays_rgb)
            kc.tag, kc.rays = tag, (rays_ori, rays_dir, rays_rgb)
            self._key_cache = kc
        return kc

    # ------------------------------------------------------------------ query-side
    @torch.no_grad()
    def score_tokens(self, tokens_pe: torch.Tensor, cache: RayKeyCache, want_map: bool = False):
        """[n_img,398] image tokens -> (scores[n_rays], attention_map or None, (m, z))."""
        q = ops.project_queries(tokens_pe, self.packed_weights())
        pm, pz = ops.score_pass1(cache.keys, q, self._impl)
        m, z = ops.score_merge(pm, pz, q.shape[0])
        scores, amap = ops.score_pass2(cache.keys, q, m, z, self._impl, want_map=want_map)
        return scores, amap, (m, z)

    def _camera_up(self, grid: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.normalize(self.camera_direction_prediction_network(grid), dim=-1)

    def run_attention(self, img, mask, rays_ori, rays_dir, rays_rgb):
        """-> (score[n], attention_map[n_img,n] or None, features_img_flat[n_img,384], camera_up_dir[3])
        (identification_module.py:77-92)."""
        if torch.is_grad_enabled() and any(p.requires_grad for lin in self._hot_params() for p in lin.parameters()):
            return self._run_attention_autograd(img, mask, rays_ori, rays_dir, rays_rgb)
        with torch.no_grad():
            tok_pe, tok, grid = self.backbone_wrapper(img, mask)
            cache = self._cache_for(rays_ori, rays_dir, rays_rgb)
            want_map = tok_pe.shape[0] * cache.n_rays * 4 <= self.attention_map_bytes_limit and self._impl == ops.SCORE_SIMT
            score, amap, _ = self.score_tokens(tok_pe, cache, want_map)
            up = self._camera_up(grid)
        return score, amap, tok, up

    def _run_attention_autograd(self, img, mask, rays_ori, rays_dir, rays_rgb):
        """Differentiable training-mode evaluation of run_attention (identification_module.py:77-92) with torch
        ops: PE -> MLP -> k_proj, q_proj, softmax over rays, sum over tokens.  CUDA tensors only."""
        from . import _lib

        _lib.load()  # the package never runs without its extension, training mode included
        if not rays_ori.is_cuda:
            raise _lib.SixdgsError("rays must be CUDA tensors (this package has no CPU path)")
        lin = torch.nn.functional.linear
        tok_pe, tok, grid = self.backbone_wrapper(img, mask)
        rp = self.ray_preprocessor
        if os.environ.get("SIXDGS_TRAIN_MLP", "torch") == "kernels":
            # every GEMM of the ray MLP and of the two projections, forward and backward, on sixdgs_linear
            # (train_ops._LinearFunction); torch keeps the PE, the concat, the ReLU masks and the bias sums
            k = train_ops.ray_keys(rp, self.attention, rays_ori, rays_dir, rays_rgb)
            q = train_ops.image_queries(self.attention, tok_pe)
        else:
            x = train_ops.ray_mlp_input(rp, rays_ori, rays_dir, rays_rgb)
            fea = rp.mlp2(torch.cat((rp.mlp(x), x), -1))
            q = lin(tok_pe, self.attention.q_proj.weight, self.attention.q_proj.bias)
            k = lin(fea, self.attention.k_proj.weight, self.attention.k_proj.bias)
        if os.environ.get("SIXDGS_TRAIN_SCORE", "kernels") == "torch":  # pure torch-op route (kept as the cross-check)
            amap = torch.softmax((q @ k.t()) / (q.shape[-1] ** 0.5), dim=-1)
            return amap.sum(0), amap, tok, self._camera_up(grid)
        # softmax-over-rays score: forward and backward on the kernels (no [n_img, n] map kept for the backward)
        want_map = q.shape[0] * k.shape[0] * 4 <= self.attention_map_bytes_limit
        scores, amap = _RayScoreFunction.apply(q, k, want_map)
        if not want_map:
            amap = torch.empty(q.shape[0], k.shape[0], device="meta")  # the shape is all train.py:159 reads
        return scores, amap, tok, self._camera_up(grid)

    def forward(self, img, mask, rays_ori, rays_dir, rays_rgb, rays_to_test: int = -1):
        """Training-shaped variant (identification_module.py:94-115): random ray subset first."""
        used = torch.randperm(rays_ori.shape[0], device=img.device, dtype=torch.long)
        if rays_to_test != -1:
            used = used[:rays_to_test]
        scores, amap, tok, up = self.run_attention(img, mask, rays_ori[used], rays_dir[used], rays_rgb[used])
        return scores, amap, tok, up, used

    @torch.no_grad()
    def test_image(self, img, mask, rays_ori, rays_dir, rays_rgb, rays_to_output: int = 100):
        """-> (idx[k] int64, vals[k], scores[n], camera_up_dir[3], attention_map) (identification_module.py:117-133)."""
        scores, amap, _, up = self.run_attention(img, mask, rays_ori, rays_dir, rays_rgb)
        vals, idx = ops.topk(scores, rays_to_output)
        return idx, vals, scores, up, amap

    @torch.no_grad()
    def query_pose(self, img, mask, rays_ori, rays_dir, rays_rgb, rays_to_output: int = 100,
                   cache: Optional[RayKeyCache] = None):
        """Fused query: image -> c2w[4,4] with no host synchronisation (test.py:85-198 in six launches
        + backbone + up head).  Returns (c2w, aux) with aux = [centre3, watch3, n_kept, status]."""
        cache = cache or self._cache_for(rays_ori, rays_dir, rays_rgb)
        tok_pe, tok, keep = self.backbone_wrapper.tokens_dense(img, mask)
        q = ops.project_queries(tok_pe.reshape(-1, tok_pe.shape[-1]).contiguous(), self.packed_weights())
        pm, pz = ops.score_pass1(cache.keys, q, self._impl)
        m, z = ops.score_merge(pm, pz, q.shape[0], keep.reshape(-1).to(torch.uint8))
        scores, _ = ops.score_pass2(cache.keys, q, m, z, self._impl)
        vals, idx = ops.topk(scores, rays_to_output)
        up = self._camera_up(tok.permute(2, 0, 1))
        return ops.pose_tail(rays_ori, rays_dir, idx, vals, up)
