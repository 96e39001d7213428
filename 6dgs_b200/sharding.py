"""Multi-GPU pose query: the path shards along the ellipsoid / ray axis.

Each rank owns a contiguous block of the selected ellipsoids, hence its own rays and its own slice
of the key cache; image tokens, weights and the camera-up head are replicated (they are tiny).
Only two things couple the shards, each one small all-gather per query:
  1. softmax statistics: every rank's partial (max, sum-exp) rows -> log-sum-exp merge -> (m, z)[256]
  2. candidates: every rank's local top-k (score, origin, direction) -> global top-k -> pose tail,
     computed redundantly on every rank (exactly the reference's top-100 semantics, test.py:85-198).
No ray data ever moves between GPUs.  The collectives go through torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests); the compute goes through a small backend object so
the host logic can be exercised without a GPU (tests inject an oracle-backed backend; the package
itself only ships the CUDA backend).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops


class CudaBackend:
    """The product backend: every call lands in libsixdgs.so."""

    def __init__(self, idm):
        self.idm = idm
        self.impl = idm._impl
        self.parts = int(_lib.load().sixdgs_score_parts(self.impl))

    def tokens(self, img, mask):
        """dense tokens: all 256 grid tokens + validity bytes, so a masked query needs no host sync"""
        tok_pe, tok, keep = self.idm.backbone_wrapper.tokens_dense(img, mask)
        return tok_pe.reshape(-1, tok_pe.shape[-1]).contiguous(), tok.permute(2, 0, 1), keep.reshape(-1).to(torch.uint8)

    def project(self, tok_pe):
        return ops.project_queries(tok_pe, self.idm.packed_weights())

    def pass1(self, keys, q):
        return ops.score_pass1(keys, q, self.impl)

    def merge(self, pm, pz, n_img, valid=None):
        return ops.score_merge(pm, pz, n_img, valid)

    def pass2(self, keys, q, m, z, out):
        return ops.score_pass2(keys, q, m, z, self.impl, out=out)[0]

    def topk(self, scores, k):
        return ops.topk(scores, k)

    def camera_up(self, grid):
        return self.idm._camera_up(grid)

    def pose_tail(self, ori, dirs, idx, vals, up):
        return ops.pose_tail(ori, dirs, idx, vals, up)


class ShardedPoseEstimator:
    def __init__(self, idm, rays_ori: torch.Tensor, rays_dir: torch.Tensor, cache, rank: int = 0, world: int = 1,
                 backend=None, group=None):
        self.backend = backend or CudaBackend(idm)
        self.ori, self.dirs, self.cache = rays_ori, rays_dir, cache
        self.rank, self.world, self.group = rank, world, group
        self.parts = getattr(self.backend, "parts", 1)
        if cache.scores is None:
            cache.scores = torch.empty(cache.n_rays, dtype=torch.float32, device=rays_ori.device)
        # q_proj 1, (qprep + pass) x2 on the tensor-core path, merge 1, top-k 11 (+11 global), pose tail 1
        tc = 2 if getattr(self.backend, "impl", 0) == ops.SCORE_TC else 0
        self.launches_per_query = 1 + 2 + tc + 1 + 11 + 1 + (11 if world > 1 else 0)

    def _all_gather(self, t: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        t = t.contiguous()
        if dist.get_backend(self.group) == "nccl":  # one collective kernel, no per-rank copies
            out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t, group=self.group)
            return out
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t, group=self.group)
        return torch.cat(outs, 0)

    @torch.no_grad()
    def query(self, img: torch.Tensor, mask: torch.Tensor, k: int = 100):
        """-> (c2w[4,4], aux[8]); identical on every rank."""
        b = self.backend
        tok_pe, grid, valid = b.tokens(img, mask)
        n_img = tok_pe.shape[0]
        q = b.project(tok_pe)
        pm, pz = b.pass1(self.cache.keys, q)
        if self.world > 1:
            pm, pz = self._all_gather(pm), self._all_gather(pz)
        m, z = b.merge(pm, pz, n_img, valid)
        scores = b.pass2(self.cache.keys, q, m, z, self.cache.scores)
        k_local = min(k, self.cache.n_rays)
        vals, idx = b.topk(scores, k_local)
        up = b.camera_up(grid)
        if self.world == 1:
            return b.pose_tail(self.ori, self.dirs, idx, vals, up)
        cand = torch.full((k, 7), float("-inf"), dtype=torch.float32, device=scores.device)
        cand[:k_local, 0] = vals
        cand[:k_local, 1:4] = self.ori[idx]
        cand[:k_local, 4:7] = self.dirs[idx]
        allc = self._all_gather(cand)
        gvals, gidx = b.topk(allc[:, 0].contiguous(), k)
        return b.pose_tail(allc[:, 1:4].contiguous(), allc[:, 4:7].contiguous(), gidx, gvals, up)
