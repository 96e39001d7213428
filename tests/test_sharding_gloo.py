"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and one world_size-4) gloo processes run
ShardedPoseEstimator with an oracle-backed compute backend (tests may use the oracle; the package
only ships the CUDA backend) and must reproduce the single-process oracle pose exactly."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleBackend:
    """torch-CPU stand-in for CudaBackend built from oracle functions (test infrastructure)."""
    parts = 1
    impl = 0

    def __init__(self, oracle, weights, tok_pe, up):
        self.o, self.w, self.tok_pe, self.up = oracle, weights, tok_pe, up
        self.front_batches = []  # how many images each front-end call saw

    def tokens(self, imgs, masks):
        # an image filled with the value i sees the fixture tokens scaled by (1 + 0.05 i): distinct queries, same
        # rays, and the tokens depend on the image content only (not on its position in whatever batch a rank holds)
        nb = imgs.shape[0]
        self.front_batches.append(nb)
        return torch.stack([self.tok_pe * (1.0 + 0.05 * float(imgs[i].mean())) for i in range(nb)]), None, None

    def project(self, tok_pe):
        return torch.nn.functional.linear(tok_pe, self.w["attention.q_proj.weight"], self.w["attention.q_proj.bias"])

    def pass1(self, keys, q):
        L = (q @ keys.t()) / (384 ** 0.5)
        m = L.max(1).values
        z = torch.exp(L - m[:, None]).sum(1)
        pad = 256 - m.shape[0]
        return (torch.cat((m, torch.full((pad,), -float("inf"))))[None], torch.cat((z, torch.zeros(pad)))[None])

    def merge(self, pm, pz, n_img, valid=None, rows=None, groups=1, group_stride=0, first_row=0):
        if rows is not None:
            sel = torch.cat([torch.arange(first_row + g * group_stride, first_row + g * group_stride + rows) for g in range(groups)])
            pm, pz = pm[sel], pz[sel]
        m = pm.max(0).values
        z = (pz * torch.exp(pm - m[None])).nan_to_num(0.0).sum(0)
        return m[:n_img], z[:n_img]

    def pass2(self, keys, q, m, z, out):
        L = (q @ keys.t()) / (384 ** 0.5)
        out.copy_((torch.exp(L - m[:, None]) / z[:, None]).sum(0))
        return out

    def topk(self, scores, k):
        t = torch.topk(scores, k)
        return t.values, t.indices

    def camera_up(self, grid):
        return self.up[None].expand(8, -1)

    def candidates(self, vals, idx, ori, dirs, k, out):
        out.fill_(float("-inf"))
        n = idx.shape[0]
        out[:n, 0], out[:n, 1:4], out[:n, 4:7] = vals, ori[idx], dirs[idx]
        return out

    def pose_tail_candidates(self, cand, idx, vals, up):
        return self.pose_tail(cand[:, 1:4].contiguous(), cand[:, 4:7].contiguous(), idx, vals, up)

    def pose_tail(self, ori, dirs, idx, vals, up):
        c2w, aux = self.o.pose_tail(idx, vals, ori, dirs, up)
        return c2w, torch.zeros(8)


class OracleBatchBackend(OracleBackend):
    """adds the several-queries-per-sweep entry points (same results as the per-query calls, one call per batch)"""
    impl = 1  # what ShardedPoseEstimator(multi_query=True) requires of its backend

    def __init__(self, *a):
        super().__init__(*a)
        self.batch_calls = 0

    def pass1_batch(self, keys, q):
        assert q.is_contiguous()
        self.batch_calls += 1
        parts = [self.pass1(keys, q[i]) for i in range(q.shape[0])]
        return torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])

    def pass2_batch(self, keys, q, m, z, out):
        self.batch_calls += 1
        for i in range(q.shape[0]):
            self.pass2(keys, q[i], m[i], z[i], out[i])
        return out


    # all-ray weighted least squares (solve="weighted_ls"): the 13 sums of the pass-2 epilogue and their solve
    def pass2_batch_ls(self, keys, q, m, z, out, ori, dirs):
        self.pass2_batch(keys, q, m, z, out)
        o, d = ori.double(), dirs.double()
        P = torch.eye(3, dtype=torch.float64)[None] - d[:, :, None] * d[:, None, :]
        Po = (P @ o[:, :, None])[:, :, 0]
        sys_ = []
        for i in range(q.shape[0]):
            w = out[i].double()
            R = (w[:, None, None] * P).sum(0)
            sys_.append(torch.cat((torch.stack((R[0, 0], R[0, 1], R[0, 2], R[1, 1], R[1, 2], R[2, 2])),
                                   (w[:, None] * Po).sum(0), (w[:, None] * d).sum(0), w.sum()[None])))
        return out, torch.stack(sys_)

    def ls_solve(self, ls_sys, weight_scale, up):
        c2w = []
        for s, u in zip(ls_sys, up):
            R = torch.stack((s[[0, 1, 2]], s[[1, 3, 4]], s[[2, 4, 5]])) * weight_scale
            centre = torch.linalg.solve(R, s[6:9] * weight_scale).float()
            watch = torch.nn.functional.normalize(s[9:12], dim=0).float()
            out = torch.eye(4)
            out[:3, :3] = torch.linalg.inv(self.o.make_rotation_mat(-watch, u))
            out[:3, 3] = centre
            c2w.append(out)
        return torch.stack(c2w), torch.zeros(len(c2w), 8)


def _worker(rank, world, port, out_path, nb=1, front_end="replicated", local=False, multi_query=False, solve="topk",
            empty_last=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sx = importlib.import_module("6dgs_b200")
    oracle = importlib.import_module("sixdgs_oracle")
    from conftest import load_golden
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    w = sx.synthetic.synth_id_weights(seed=g["weight_seed"])
    ori, dirs, rgb = r["ori"], r["dirs"], r["rgb"]
    n = ori.shape[0]
    per = (n + world - 1) // world
    lo, hi = rank * per, min((rank + 1) * per, n)
    if empty_last:  # ragged sharding: the last rank owns no rays at all
        lo, hi = (0, n) if rank == 0 else (n, n)
    keys = torch.nn.functional.linear(oracle.ray_features(ori[lo:hi], dirs[lo:hi], rgb[lo:hi], w),
                                      w["attention.k_proj.weight"], w["attention.k_proj.bias"])
    cache = sx.RayKeyCache(keys, hi - lo, ())
    be = (OracleBatchBackend if multi_query else OracleBackend)(oracle, w, g["tok_pe"], g["up"])
    est = sx.ShardedPoseEstimator(None, ori[lo:hi].contiguous(), dirs[lo:hi].contiguous(), cache, rank, world, backend=be,
                                  front_end=front_end, multi_query=multi_query, solve=solve)
    imgs = torch.arange(nb, dtype=torch.float32)[:, None, None, None].expand(nb, 2, 2, 3).contiguous()
    masks = torch.ones(nb, 2, 2, dtype=torch.bool)
    if local:  # hand over this rank's images only
        bl = nb // world
        imgs, masks = imgs[rank * bl:(rank + 1) * bl], masks[rank * bl:(rank + 1) * bl]
    c2w, _ = est.query_batch(imgs, masks, k=100, local=local)
    shards = front_end == "sharded" and nb % world == 0
    assert be.front_batches == [nb // world if shards else nb]  # the front end ran once, on this rank's share
    if hi > lo:
        assert getattr(be, "batch_calls", 0) == (2 if multi_query else 0)  # one call per pass for the whole batch
    # every rank must hold the same poses
    gathered = [torch.empty_like(c2w) for _ in range(world)]
    dist.all_gather(gathered, c2w)
    assert all(torch.equal(gathered[0], x) for x in gathered)
    if rank == 0:
        torch.save(c2w, out_path)
    dist.destroy_process_group()


# replicated front end, batch 3; sharded front end with an odd batch (falls back to replicated), with an even batch
# given whole, and with each rank given only its own images
@pytest.mark.timeout(300)
@pytest.mark.parametrize("nb,front_end,local,multi_query", [(3, "replicated", False, False), (3, "sharded", False, False),
                                                            (4, "sharded", False, False), (4, "sharded", True, False),
                                                            (3, "replicated", False, True), (4, "sharded", True, True)])
def test_two_rank_sharded_query_matches_single_process(oracle, synthetic, tmp_path, nb, front_end, local, multi_query):
    from conftest import load_golden
    out = str(tmp_path / "c2w.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, nb, front_end, local, multi_query), nprocs=2, join=True)
    c2w = torch.load(out)
    assert c2w.shape == (nb, 4, 4)
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    ref, _ = oracle.pose_tail(g["topk_idx"], g["topk_vals"], r["ori"], r["dirs"], g["up"])
    torch.testing.assert_close(c2w[0], ref, rtol=1e-5, atol=1e-5)
    # the other queries of the batch (scaled tokens) against the unsharded oracle
    w = synthetic.synth_id_weights(seed=g["weight_seed"])
    fea = oracle.ray_features(r["ori"], r["dirs"], r["rgb"], w)
    for i in range(1, nb):
        score, _ = oracle.attention_scores(g["tok_pe"] * (1.0 + 0.05 * i), fea, w, return_map=False)
        top = torch.topk(score, 100)
        ref_i, _ = oracle.pose_tail(top.indices, top.values, r["ori"], r["dirs"], g["up"])
        torch.testing.assert_close(c2w[i], ref_i, rtol=1e-4, atol=1e-4)
        assert not torch.allclose(c2w[i], c2w[0])


@pytest.mark.timeout(300)
@pytest.mark.parametrize("solve", ["topk", "weighted_ls"])
def test_four_rank_bench_configuration(oracle, synthetic, tmp_path, solve):
    """the shape of `bench.py --gpus 4`: 8 queries per batch, sharded front end with each rank handed its own 2 images,
    batched sweeps -- every rank ends with the 8 poses of the unsharded oracle"""
    from conftest import load_golden
    out = str(tmp_path / "c2w.pt")
    nb = 8
    mp.spawn(_worker, args=(4, _free_port(), out, nb, "sharded", True, True, solve), nprocs=4, join=True)
    c2w = torch.load(out)
    assert c2w.shape == (nb, 4, 4)
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    w = synthetic.synth_id_weights(seed=g["weight_seed"])
    fea = oracle.ray_features(r["ori"], r["dirs"], r["rgb"], w)
    for i in range(nb):
        score, _ = oracle.attention_scores(g["tok_pe"] * (1.0 + 0.05 * i), fea, w, return_map=False)
        if solve == "topk":
            top = torch.topk(score, 100)
            ref, _ = oracle.pose_tail(top.indices, top.values, r["ori"], r["dirs"], g["up"])
            torch.testing.assert_close(c2w[i], ref, rtol=1e-4, atol=1e-4)
        else:
            centre = oracle.line_intersection(r["ori"], -r["dirs"], score / 256)
            torch.testing.assert_close(c2w[i, :3, 3], centre, rtol=1e-4, atol=1e-4)


@pytest.mark.timeout(300)
def test_two_rank_weighted_least_squares_one_allreduce(oracle, synthetic, tmp_path):
    """solve="weighted_ls": each rank accumulates the weighted LS system of ITS rays (weights = its scores), one
    all-reduce sums the two 13-double systems, every rank solves -> the unsharded oracle's all-ray weighted LS
    (least_squared_loss.py:62-64: compute_line_intersection_impl2(ori, -dir, score / n_img)) and watch direction"""
    from conftest import load_golden
    out = str(tmp_path / "c2w.pt")
    nb = 3
    mp.spawn(_worker, args=(2, _free_port(), out, nb, "replicated", False, True, "weighted_ls"), nprocs=2, join=True)
    c2w = torch.load(out)
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    w = synthetic.synth_id_weights(seed=g["weight_seed"])
    fea = oracle.ray_features(r["ori"], r["dirs"], r["rgb"], w)
    for i in range(nb):
        score, _ = oracle.attention_scores(g["tok_pe"] * (1.0 + 0.05 * i), fea, w, return_map=False)
        wts = score / 256
        centre = oracle.line_intersection(r["ori"], -r["dirs"], wts)
        watch = torch.nn.functional.normalize((wts[:, None] * r["dirs"]).sum(0), dim=0)
        torch.testing.assert_close(c2w[i, :3, 3], centre, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(c2w[i, :3, :3], torch.linalg.inv(oracle.make_rotation_mat(-watch, g["up"])), rtol=1e-4, atol=1e-4)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("multi_query,solve", [(False, "topk"), (True, "topk"), (True, "weighted_ls")])
def test_rank_without_rays_is_neutral(oracle, synthetic, tmp_path, multi_query, solve):
    """ragged sharding: rank 1 owns no rays (more ranks than blocks of ellipsoids).  Its statistics are (-inf, 0), its
    candidates all -inf, its least-squares system zero -- the poses equal the single-process oracle's"""
    from conftest import load_golden
    out = str(tmp_path / "c2w.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, 2, "replicated", False, multi_query, solve, True), nprocs=2, join=True)
    c2w = torch.load(out)
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    w = synthetic.synth_id_weights(seed=g["weight_seed"])
    fea = oracle.ray_features(r["ori"], r["dirs"], r["rgb"], w)
    for i in range(2):
        score, _ = oracle.attention_scores(g["tok_pe"] * (1.0 + 0.05 * i), fea, w, return_map=False)
        if solve == "topk":
            top = torch.topk(score, 100)
            ref, _ = oracle.pose_tail(top.indices, top.values, r["ori"], r["dirs"], g["up"])
            torch.testing.assert_close(c2w[i], ref, rtol=1e-4, atol=1e-4)
        else:
            centre = oracle.line_intersection(r["ori"], -r["dirs"], score / 256)
            torch.testing.assert_close(c2w[i, :3, 3], centre, rtol=1e-4, atol=1e-4)


def test_single_rank_path_uses_no_collective(sx, oracle, synthetic):
    from conftest import load_golden
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    w = synthetic.synth_id_weights(seed=g["weight_seed"])
    keys = torch.nn.functional.linear(oracle.ray_features(r["ori"], r["dirs"], r["rgb"], w),
                                      w["attention.k_proj.weight"], w["attention.k_proj.bias"])
    est = sx.ShardedPoseEstimator(None, r["ori"], r["dirs"], sx.RayKeyCache(keys, keys.shape[0], ()), 0, 1,
                                  backend=OracleBackend(oracle, w, g["tok_pe"], g["up"]))
    c2w, _ = est.query(torch.zeros(2, 2, 3), torch.ones(2, 2, dtype=torch.bool))
    with pytest.raises(ValueError):  # nothing to shard the front end over
        est.query_batch(torch.zeros(1, 2, 2, 3), torch.ones(1, 2, 2, dtype=torch.bool), local=True)
    with pytest.raises(ValueError):
        sx.ShardedPoseEstimator(None, r["ori"], r["dirs"], est.cache, 0, 1, backend=est.backend, front_end="rank0")
    with pytest.raises(ValueError):  # the batched sweep exists on the tensor-core path only
        sx.ShardedPoseEstimator(None, r["ori"], r["dirs"], est.cache, 0, 1, backend=est.backend, multi_query=True)
    with pytest.raises(ValueError):  # ... and so does the fused weighted least squares
        sx.ShardedPoseEstimator(None, r["ori"], r["dirs"], est.cache, 0, 1, backend=est.backend, solve="weighted_ls")
    ref, _ = oracle.pose_tail(g["topk_idx"], g["topk_vals"], r["ori"], r["dirs"], g["up"])
    torch.testing.assert_close(c2w, ref, rtol=1e-5, atol=1e-5)
