"""Multi-GPU pose query: the path shards along the ellipsoid / ray axis.

Each rank owns a contiguous block of the selected ellipsoids, hence its own rays and its own slice
of the key cache; image tokens, weights and the camera-up head are replicated (they are tiny).
Only two things couple the shards, each one small all-gather per batch of queries:
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

    def tokens(self, imgs, masks):
        """dense tokens for a batch of images [B,H,W,3]: all 256 grid tokens + validity bytes per image, so a masked
        query needs no host sync.  -> (tok_pe [B,256,398], grid [B,384,16,16], valid [B,256] uint8)"""
        tok_pe, tok, keep = self.idm.backbone_wrapper.tokens_dense_batch(imgs, masks)
        b = imgs.shape[0]
        return (tok_pe.reshape(b, -1, tok_pe.shape[-1]).contiguous(), tok.permute(0, 3, 1, 2),
                keep.reshape(b, -1).to(torch.uint8).contiguous())

    def project(self, tok_pe):
        return ops.project_queries(tok_pe, self.idm.packed_weights())

    def pass1(self, keys, q):
        return ops.score_pass1(keys, q, self.impl)

    def merge(self, pm, pz, n_img, valid=None, rows=None, groups=1, group_stride=0, first_row=0):
        return ops.score_merge(pm, pz, n_img, valid, rows, groups, group_stride, first_row)

    def candidates(self, vals, idx, ori, dirs, k, out):
        return ops.gather_candidates(vals, idx, ori, dirs, k, out)

    def pose_tail_candidates(self, cand, idx, vals, up):
        return ops.pose_tail_candidates(cand, idx, vals, up)

    def pass2(self, keys, q, m, z, out):
        return ops.score_pass2(keys, q, m, z, self.impl, out=out)[0]

    # several queries per key sweep (ShardedPoseEstimator(multi_query=True); bf16 or f16x2 key cache)
    def pass1_batch(self, keys, q):
        return ops.score_pass1_batch(keys, q)

    def pass2_batch(self, keys, q, m, z, out):
        return ops.score_pass2_batch(keys, q, m, z, out=out)

    def pass2_batch_ls(self, keys, q, m, z, out, ori, dirs):
        """pass 2 + the all-ray weighted least-squares system in its epilogue -> (scores [B,n], ls_sys [B,13] float64)"""
        return ops.score_pass2_batch(keys, q, m, z, out=out, ls_rays=(ori, dirs))

    def ls_solve(self, ls_sys, weight_scale, up):
        return ops.ls_solve(ls_sys, weight_scale, up)

    def topk(self, scores, k):
        return ops.topk(scores, k)

    def camera_up(self, grid):
        """[B,384,16,16] -> unit up vectors [B,3]"""
        return torch.nn.functional.normalize(self.idm.camera_direction_prediction_network.forward_batch(grid), dim=-1)

    def pose_tail(self, ori, dirs, idx, vals, up):
        return ops.pose_tail(ori, dirs, idx, vals, up)


class ShardedPoseEstimator:
    """One rank's view of a (possibly sharded) scene: its rays, its key-cache slice and the query pipeline.

    ``query`` launches eagerly.  ``enable_cuda_graphs`` captures the pipeline with static buffers: a single
    graph on one GPU; on several GPUs three graph segments with the two NCCL all-gathers issued eagerly
    between them (a collective inside a captured graph ties the graph to NCCL's internal state and proved
    fragile; two eager launches per query cost ~10 us).

    ``front_end``: what the ranks do with the image front end (resize, backbone, q projection, up head) of a
    batch.  "sharded" (default; a no-op on one rank): when the batch size is a multiple of the world size each rank
    runs it for its B/world images only and one more all-gather (q rows, up vector and token validity, ~394 KB per
    query) hands every rank the whole batch; other batch sizes fall back to "replicated": every rank runs it for the
    whole batch -- no extra collective, but the replicated milliseconds do not shrink with the world size.  With ``query_batch(..., local=True)`` the caller passes just this rank's images, so only
    those cross PCIe.

    ``multi_query`` (tensor-core path only; default: on for it): score the whole batch in one sweep over the key cache
    per pass (csrc/score_tc_mq.cu) instead of one sweep per query -- bit-identical scores, the keys cross HBM once per
    pass per batch.

    ``solve``: "topk" (default) is the reference's evaluation path (test.py:85-198): top-100 rays, dedup, unweighted LS,
    watch direction from the re-weighted winners; across shards the local winners are all-gathered.  "weighted_ls" is the
    all-ray weighted least squares of least_squared_loss.py:47-64 (weights = score / n_img over EVERY ray): the 3x3
    system, sum w d and sum w are accumulated in the pass-2 epilogue, the shards' 13-double systems are summed by ONE
    all-reduce, and the pose comes from (centre, normalised sum w d, camera up) -- no top-k at all.  Needs the batched
    tensor-core kernel (multi_query)."""

    def __init__(self, idm, rays_ori: torch.Tensor, rays_dir: torch.Tensor, cache, rank: int = 0, world: int = 1,
                 backend=None, group=None, front_end: str = "sharded", multi_query: Optional[bool] = None,
                 solve: str = "topk"):
        if front_end not in ("replicated", "sharded"):
            raise ValueError(f"front_end must be 'replicated' or 'sharded', got {front_end!r}")
        if solve not in ("topk", "weighted_ls"):
            raise ValueError(f"solve must be 'topk' or 'weighted_ls', got {solve!r}")
        self.backend = backend or CudaBackend(idm)
        self.ori, self.dirs, self.cache = rays_ori, rays_dir, cache
        self.rank, self.world, self.group = rank, world, group
        self.front_end = front_end
        if multi_query is None:
            multi_query = getattr(self.backend, "impl", ops.SCORE_SIMT) == ops.SCORE_TC
        self.multi_query = bool(multi_query)
        if self.multi_query and getattr(self.backend, "impl", ops.SCORE_TC) != ops.SCORE_TC:
            raise ValueError("multi_query needs the tensor-core score path (score_impl='tc_bf16')")
        self.solve = solve
        if solve == "weighted_ls" and not self.multi_query:
            raise ValueError("solve='weighted_ls' needs the batched tensor-core score kernel (multi_query)")
        self._scores_b = None  # [B, n_rays] score rows of a batch (multi_query)
        self.parts = getattr(self.backend, "parts", 1)
        if cache.scores is None:
            cache.scores = torch.empty(cache.n_rays, dtype=torch.float32, device=rays_ori.device)
        # libsixdgs launches per query: pass 1 + merge + pass 2 (3, +2 q-prep on the tensor-core path), radix top-k
        # (12; 1 when the shard has <= 4096 rays), pose tail 1, and when sharded candidate pack 1 + global top-k 1.
        # The q projection is one launch per batch (launches_per_batch).
        tc = 2 if getattr(self.backend, "impl", 0) == ops.SCORE_TC else 0
        radix = 12  # init, 4 x (histogram, select), gather, tie fix-up, final
        self.launches_per_query = 3 + tc + (radix if cache.n_rays > 4096 else 1) + 1 + (2 if world > 1 else 0)
        self.launches_per_batch = 1
        if self.multi_query:  # per query only the merge remains; per batch (of <= 8) q-prep + kernel for each pass
            self.launches_per_query -= 2 + tc
            self.launches_per_batch += 4
        if self.solve == "weighted_ls":  # no top-k / candidate exchange / pose tail: merge per query; per batch the
            self.launches_per_query = 1  # partial-system reduction and one solve for all queries
            self.launches_per_batch += 2
        self._g = None  # captured graphs + static buffers

    # ------------------------------------------------------------------ collectives
    def _all_gather(self, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        import torch.distributed as dist

        t = t.contiguous()
        if dist.get_backend(self.group) == "nccl":  # one collective kernel, no per-rank copies
            if out is None:
                out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t, group=self.group)
            return out
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t, group=self.group)
        res = torch.cat(outs, 0)
        if out is not None:
            out.copy_(res)
            return out
        return res

    # ------------------------------------------------------------------ pipeline stages (no collectives inside)
    # Every stage works on a BATCH of B queries: the image front end (resize, backbone, q projection, up head)
    # runs once for the batch -- it is latency-bound, so B images cost about as much as one -- while the key
    # cache is streamed per query (two passes each).  The collectives are per batch as well.
    def _front(self, imgs, masks):
        """images -> tokens -> (q [B,n_img,384], unit up vectors [>=B,3], token validity [B,n_img] uint8 or None)"""
        b = self.backend
        tok_pe, grid, valid = b.tokens(imgs, masks)
        nb, n_img = tok_pe.shape[0], tok_pe.shape[1]
        q = b.project(tok_pe.reshape(nb * n_img, -1)).reshape(nb, n_img, -1)
        return q, b.camera_up(grid), valid

    def _pass1_all(self, q, up, valid):
        """this shard's partial softmax rows for every query of the batch"""
        b = self.backend
        nb, n_img = q.shape[0], q.shape[1]
        if self.cache.n_rays == 0:
            # a shard without rays (more ranks than blocks of ellipsoids): neutral statistics, (max, sum) = (-inf, 0),
            # in the row layout the other ranks use, so the all-gather stays regular
            rows = self.parts
            pm = torch.full((nb * rows, _lib.MAX_TOKENS), float("-inf"), dtype=torch.float32, device=q.device)
            return {"q": q, "valid": valid, "up": up, "pmz": torch.cat((pm, torch.zeros_like(pm)), 0), "n_img": n_img,
                    "nb": nb, "rows": rows}
        if self.multi_query:
            q = q.contiguous()  # the batched kernel reads the batch as one [B*256, 384] matrix
            pm, pz = b.pass1_batch(self.cache.keys, q)  # [B * parts, 256], query-major like the loop below
            rows = pm.shape[0] // nb
        else:
            parts = [b.pass1(self.cache.keys, q[i]) for i in range(nb)]
            pm = torch.cat([p[0] for p in parts], 0)  # [B * parts, 256]
            pz = torch.cat([p[1] for p in parts], 0)
            rows = parts[0][0].shape[0]
        # one buffer for both statistics, [2 * B * rows, 256] = pm rows then pz rows: ONE all-gather per batch
        return {"q": q, "valid": valid, "up": up, "pmz": torch.cat((pm, pz), 0), "n_img": n_img, "nb": nb, "rows": rows}

    def _stage1(self, imgs, masks):
        """images -> tokens -> q, camera up, and this shard's partial softmax rows for every query of the batch"""
        return self._pass1_all(*self._front(imgs, masks))

    # ---- sharded front end: each rank's (q, up, valid) rows travel as one fp32 record per query, padded to a
    # multiple of 64 floats so that every q block of the gathered buffer stays 256-byte aligned for the kernels
    @staticmethod
    def _record_len(n_img, d):
        return -(-(n_img * d + 3 + n_img) // 64) * 64

    def _pack_front(self, q, up, valid):
        bl, n_img, d = q.shape
        rec = torch.zeros((bl, self._record_len(n_img, d)), dtype=torch.float32, device=q.device)
        rec[:, :n_img * d] = q.reshape(bl, -1)
        rec[:, n_img * d:n_img * d + 3] = up[:bl]
        if valid is None:
            rec[:, n_img * d + 3:n_img * d + 3 + n_img] = 1.0
        else:
            rec[:, n_img * d + 3:n_img * d + 3 + n_img] = valid.to(torch.float32)
        return rec

    @staticmethod
    def _unpack_front(rec, n_img, d):
        q = rec[:, :n_img * d].unflatten(1, (n_img, d))  # a view: q[i] is contiguous and aligned
        up = rec[:, n_img * d:n_img * d + 3]
        valid = rec[:, n_img * d + 3:n_img * d + 3 + n_img].to(torch.uint8)
        return q, up, valid

    def _shards_front(self, n_batch: int, local: bool) -> bool:
        if local:
            if self.front_end != "sharded" or self.world == 1:
                raise ValueError("local=True needs front_end='sharded' and more than one rank")
            return True
        return self.front_end == "sharded" and self.world > 1 and n_batch % self.world == 0

    def _local_chunk(self, imgs, masks, local):
        if local:
            return imgs, masks
        bl = imgs.shape[0] // self.world
        return imgs[self.rank * bl:(self.rank + 1) * bl], masks[self.rank * bl:(self.rank + 1) * bl]

    def _stage2(self, pmz, st, k):
        """merged statistics -> scores -> local top-k per query (+ packed candidates when sharded).
        pmz is [world * 2 * B * rows, 256]: per rank the pm rows of the batch, then its pz rows; query i owns rows
        [r*2*B*rows + i*rows, +rows) (pm) and the same shifted by B*rows (pz) of every rank r."""
        b = self.backend
        nb, rows = st["nb"], st["rows"]
        groups = pmz.shape[0] // (2 * nb * rows)  # == world
        pm, pz = pmz, pmz[nb * rows:]
        k_local = min(k, self.cache.n_rays)
        vals, idxs = [], []
        cand = None
        if self.world > 1:
            cand = torch.empty((nb, k, 7), dtype=torch.float32, device=self.ori.device)
        if self.cache.n_rays == 0:  # empty shard: no scores, no local winners; its candidate rows are all -inf
            ev = torch.empty(0, dtype=torch.float32, device=self.ori.device)
            ei = torch.empty(0, dtype=torch.int64, device=self.ori.device)
            if cand is not None:
                for i in range(nb):
                    b.candidates(ev, ei, self.ori, self.dirs, k, cand[i])
            return [ev] * nb, [ei] * nb, cand

        def merged(i):
            # query i's rows: [rank g][query i][0..rows) -> group stride nb*rows, first row i*rows (no copies)
            return b.merge(pm, pz, st["n_img"], st["valid"][i] if st["valid"] is not None else None,
                           rows=rows, groups=groups, group_stride=2 * nb * rows, first_row=i * rows)

        scores_b = None
        if self.multi_query:
            mz = [merged(i) for i in range(nb)]
            if self._scores_b is None or self._scores_b.shape[0] < nb:
                self._scores_b = torch.empty((nb, self.cache.n_rays), dtype=torch.float32, device=self.ori.device)
            scores_b = b.pass2_batch(self.cache.keys, st["q"], torch.stack([x[0] for x in mz]),
                                     torch.stack([x[1] for x in mz]), self._scores_b[:nb])
        for i in range(nb):
            if scores_b is not None:
                scores = scores_b[i]
            else:
                m, z = merged(i)
                scores = b.pass2(self.cache.keys, st["q"][i], m, z, self.cache.scores)
            v, ix = b.topk(scores, k_local)
            vals.append(v)
            idxs.append(ix)
            if cand is not None:
                b.candidates(v, ix, self.ori, self.dirs, k, cand[i])
        return vals, idxs, cand

    def _all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _stage2_weighted(self, pmz, st):
        """merged statistics -> scores + this shard's weighted least-squares system per query, ls_sys [B,13] float64"""
        b = self.backend
        nb, rows = st["nb"], st["rows"]
        if self.cache.n_rays == 0:  # empty shard: contributes nothing to the sums
            return torch.zeros((nb, 13), dtype=torch.float64, device=self.ori.device)
        groups = pmz.shape[0] // (2 * nb * rows)
        pm, pz = pmz, pmz[nb * rows:]
        mz = [b.merge(pm, pz, st["n_img"], st["valid"][i] if st["valid"] is not None else None, rows=rows, groups=groups,
                      group_stride=2 * nb * rows, first_row=i * rows) for i in range(nb)]
        if self._scores_b is None or self._scores_b.shape[0] < nb:
            self._scores_b = torch.empty((nb, self.cache.n_rays), dtype=torch.float32, device=self.ori.device)
        _, ls_sys = b.pass2_batch_ls(self.cache.keys, st["q"], torch.stack([x[0] for x in mz]),
                                     torch.stack([x[1] for x in mz]), self._scores_b[:nb], self.ori, self.dirs)
        return ls_sys

    def _stage3_weighted(self, ls_sys, st):
        """summed systems -> poses.  weights = score / n_img with n_img the number of valid tokens of the query; the
        scale only matters for the det(R) < 1e-7 guard, the solution is scale-free"""
        return self.backend.ls_solve(ls_sys, 1.0 / st["n_img"], st["up"][:st["nb"]].contiguous())

    def _stage3(self, allc, up, k, nb):
        """allc [world * B, k, 7] (rank-major) -> per-query global top-k -> pose"""
        b = self.backend
        allc = allc.reshape(-1, nb, allc.shape[-2], 7)
        byq = allc.transpose(0, 1).contiguous()  # [B, world, k, 7]: one copy per batch
        outs = []
        for i in range(nb):
            c = byq[i].reshape(-1, 7)
            gvals, gidx = b.topk(c[:, 0].contiguous(), k)
            outs.append(b.pose_tail_candidates(c, gidx, gvals, up[i]))
        return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])

    def _query_eager(self, imgs, masks, k, local=False):
        if self._shards_front(imgs.shape[0], local):
            q, up, valid = self._front(*self._local_chunk(imgs, masks, local))
            rec = self._all_gather(self._pack_front(q, up, valid))
            st = self._pass1_all(*self._unpack_front(rec, q.shape[1], q.shape[2]))
        else:
            st = self._stage1(imgs, masks)
        pmz = st["pmz"]
        if self.world > 1:
            pmz = self._all_gather(pmz)
        if self.solve == "weighted_ls":
            ls_sys = self._stage2_weighted(pmz, st)
            if self.world > 1:
                ls_sys = self._all_reduce_sum(ls_sys)
            return self._stage3_weighted(ls_sys, st)
        vals, idxs, cand = self._stage2(pmz, st, k)
        if self.world == 1:
            outs = [self.backend.pose_tail(self.ori, self.dirs, idxs[i], vals[i], st["up"][i]) for i in range(st["nb"])]
            return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])
        return self._stage3(self._all_gather(cand), st["up"], k, st["nb"])

    # ------------------------------------------------------------------ CUDA graphs
    def enable_cuda_graphs(self, img: torch.Tensor, mask: torch.Tensor, k: int = 100, local: bool = False) -> bool:
        """Capture the pipeline for image batches of this shape ([B,H,W,3] / [B,H,W]; a single [H,W,3] image is
        treated as B = 1; ``local`` as in ``query_batch``).  Returns False (and stays eager) if capture fails."""
        if img.dim() == 3:
            img, mask = img[None], mask[None]
        if self.cache.n_rays == 0:
            return False  # an empty shard has nothing worth capturing; it stays on the eager path
        cur = torch.cuda.current_stream()
        shard_front = self._shards_front(img.shape[0], local)
        g = {"shape": tuple(img.shape), "local": local, "k": k}
        if shard_front:  # the static input is this rank's chunk only
            img, mask = self._local_chunk(img, mask, local)
        g["img"], g["mask"] = img.clone(), mask.clone()
        try:
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._query_eager(g["img"], g["mask"], k, local=shard_front)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            if self.world == 1:
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    g["out"] = self._query_eager(g["img"], g["mask"], k)
                g["graphs"] = [g1]
            else:
                g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                if shard_front:
                    g0 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g0):
                        q, up, valid = self._front(g["img"], g["mask"])
                        g["rec"] = self._pack_front(q, up, valid)
                    g["rec_all"] = self._all_gather(g["rec"])
                    g["front_graph"] = g0
                    with torch.cuda.graph(g1):
                        g["st"] = self._pass1_all(*self._unpack_front(g["rec_all"], q.shape[1], q.shape[2]))
                else:
                    with torch.cuda.graph(g1):
                        g["st"] = self._stage1(g["img"], g["mask"])
                g["pmz_all"] = self._all_gather(g["st"]["pmz"])
                if self.solve == "weighted_ls":
                    with torch.cuda.graph(g2):
                        g["ls_sys"] = self._stage2_weighted(g["pmz_all"], g["st"])
                    self._all_reduce_sum(g["ls_sys"])
                    with torch.cuda.graph(g3):
                        g["out"] = self._stage3_weighted(g["ls_sys"], g["st"])
                else:
                    with torch.cuda.graph(g2):
                        _, _, g["cand"] = self._stage2(g["pmz_all"], g["st"], k)
                    g["allc"] = self._all_gather(g["cand"])
                    with torch.cuda.graph(g3):
                        g["out"] = self._stage3(g["allc"], g["st"]["up"], k, g["st"]["nb"])
                g["graphs"] = [g1, g2, g3]
            torch.cuda.synchronize()
            # buffers that were allocated by the warm-up calls (outside the capture) but are baked into the graphs:
            # keep them alive with the graphs, or a later eager call with a larger batch would free them under a replay
            g["keepalive"] = (self._scores_b, self.cache.scores, self.cache.keys, self.ori, self.dirs)
            self._g = g
            return True
        except Exception as e:  # noqa: BLE001
            import sys
            print(f"[sixdgs] CUDA graph capture unavailable ({type(e).__name__}: {e}); staying eager", file=sys.stderr)
            self._g = None
            torch.cuda.synchronize()
            return False

    def _query_graphs(self, img, mask):
        g = self._g
        if "front_graph" in g and not g["local"]:
            img, mask = self._local_chunk(img, mask, False)
        if img.data_ptr() != g["img"].data_ptr():
            g["img"].copy_(img)
        if mask.data_ptr() != g["mask"].data_ptr():
            g["mask"].copy_(mask)
        if self.world == 1:
            g["graphs"][0].replay()
        else:
            if "front_graph" in g:
                g["front_graph"].replay()
                self._all_gather(g["rec"], g["rec_all"])
            g["graphs"][0].replay()
            self._all_gather(g["st"]["pmz"], g["pmz_all"])
            g["graphs"][1].replay()
            if self.solve == "weighted_ls":
                self._all_reduce_sum(g["ls_sys"])  # in place on the graph's static buffer (rewritten by every replay)
            else:
                self._all_gather(g["cand"], g["allc"])
            g["graphs"][2].replay()
        return g["out"]

    @torch.no_grad()
    def query_batch(self, imgs: torch.Tensor, masks: torch.Tensor, k: int = 100, local: bool = False):
        """imgs [B,H,W,3], masks [B,H,W] -> (c2w[B,4,4], aux[B,8]); identical on every rank.  With graphs enabled
        (for this batch shape) the returned tensors are the graph's static outputs, overwritten by the next call.
        ``local=True`` (front_end="sharded" only): imgs/masks hold just this rank's B/world images, in rank order;
        the result still covers the whole batch."""
        g = self._g
        if g is not None and g["k"] == k and tuple(imgs.shape) == g["shape"] and local == g["local"]:
            return self._query_graphs(imgs, masks)
        return self._query_eager(imgs, masks, k, local)

    @torch.no_grad()
    def query(self, img: torch.Tensor, mask: torch.Tensor, k: int = 100):
        """single query: img [H,W,3], mask [H,W] -> (c2w[4,4], aux[8])"""
        c2w, aux = self.query_batch(img[None], mask[None], k)
        return c2w[0], aux[0]
