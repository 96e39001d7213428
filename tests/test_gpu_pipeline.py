"""GPU tests of the batch / multi-GPU pipeline pieces that were written without a GPU at the end of round 1 and
validated on a B200 at the start of round 2 (profiles/validate_experimental_r2.log): the multi-query score kernel, the
sharded image front end, the staged TMA-store GEMM epilogue, the heavy-tailed ray-generation fixture; plus the
deterministic tie handling of the top-k and the key-cache identity rule."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_staged_tma_store_epilogue_matches_the_direct_store_kernel(sx, synthetic):
    """features_tc.cu: linear_tc_staged_kernel (smem-staged TMA tensor stores, the default TF32 build) must be
    bit-identical to linear_tc_kernel (same MMAs, same epilogue math; only the way the tile reaches HBM differs)"""
    dev = "cuda"
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone())
    idm.load_state_dict(synthetic.synth_id_weights(seed=3), strict=False)
    pw = idm.to(dev).packed_weights()
    gen = torch.Generator().manual_seed(4)
    for n in (1, 127, 128, 129, 200_000):
        ori = (torch.randn(n, 3, generator=gen) * 3).to(dev)
        dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1).to(dev)
        rgb = torch.rand(n, 3, generator=gen).to(dev)
        for kd in (sx._lib.F32, sx._lib.BF16):
            k1, f1 = sx.ops.ray_features(ori, dirs, rgb, pw, k_dtype=kd, want_features=True, impl=sx.ops.FEATURES_TC_DIRECT)
            k3, f3 = sx.ops.ray_features(ori, dirs, rgb, pw, k_dtype=kd, want_features=True, impl=sx.ops.FEATURES_TC)
            torch.cuda.synchronize()
            assert torch.equal(k1, k3) and torch.equal(f1, f3), (n, kd)


def test_sharded_front_end_two_shards_in_one_process(sx, synthetic, oracle):
    """ShardedPoseEstimator(front_end="sharded") on the CUDA backend: each of two shards runs the image front end
    for ONE of the two images, the all-gather of the packed (q, up, validity) records is emulated by concatenation
    in rank order, and the rest of the pipeline must reproduce the unsharded poses (the host logic alone is
    covered on CPU by tests/test_sharding_gloo.py)."""
    from conftest import load_golden
    dev = "cuda"
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    ori, dirs, rgb = r["ori"].to(dev), r["dirs"].to(dev), r["rgb"].to(dev)
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone(), score_impl="simt_fp32")
    idm.load_state_dict(synthetic.synth_id_weights(seed=g["weight_seed"]), strict=False)
    idm = idm.to(dev).eval().requires_grad_(False)
    img = g["img"].to(dev)
    imgs = torch.stack((img, img.flip(1)))
    masks = torch.stack((torch.ones(64, 64, dtype=torch.bool, device=dev), g["mask2"].to(dev)))
    full = sx.ShardedPoseEstimator(idm, ori, dirs, idm.build_key_cache(ori, dirs, rgb))
    ref, _ = full.query_batch(imgs, masks)
    n = ori.shape[0]
    cut = n // 2 + 37
    shards = []
    for rank, (lo, hi) in enumerate(((0, cut), (cut, n))):
        o, d, c = ori[lo:hi].contiguous(), dirs[lo:hi].contiguous(), rgb[lo:hi].contiguous()
        shards.append(sx.ShardedPoseEstimator(idm, o, d, idm.build_key_cache(o, d, c), rank, 2, front_end="sharded"))
    recs = []
    for s in shards:
        assert s._shards_front(2, False)
        q, up, valid = s._front(*s._local_chunk(imgs, masks, False))
        assert q.shape[0] == 1
        recs.append(s._pack_front(q, up, valid))
    rec = torch.cat(recs)
    assert rec.shape[1] % 64 == 0
    k = 100
    sts = [s._pass1_all(*s._unpack_front(rec, 256, 384)) for s in shards]
    assert all(st["q"][i].is_contiguous() and st["q"][i].data_ptr() % 256 == 0 for st in sts for i in range(2))
    pmz = torch.cat([st["pmz"] for st in sts])
    allc = torch.cat([s._stage2(pmz, st, k)[2] for s, st in zip(shards, sts)])
    for s, st in zip(shards, sts):
        c2w, _ = s._stage3(allc, st["up"], k, st["nb"])
        torch.testing.assert_close(c2w, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n_rays,n_img", [(1, 256), (100, 256), (5513, 201), (74 * 256 * 3 + 77, 256)])
def test_multi_query_score_kernel_equals_per_query_kernel(sx, n_rays, n_img):
    """score_tc_mq.cu (several queries per key sweep) against score_tc.cu (validated): same tiles per CTA pair, same
    MMA order per accumulator, same epilogue arithmetic -> the partial softmax rows and the scores must be
    bit-identical, for batch sizes below, at and above one launch's capacity."""
    dev = "cuda"
    gen = torch.Generator().manual_seed(n_rays)
    keys = (torch.randn(n_rays, 384, generator=gen) * 0.7).to(dev).to(torch.bfloat16)
    parts = int(sx._lib.load().sixdgs_score_batch_parts())
    assert parts == int(sx._lib.load().sixdgs_score_parts(sx.ops.SCORE_TC))
    for nb in (1, 3, sx.ops.score_batch_max(), sx.ops.score_batch_max() + 3):
        q = (torch.randn(nb, 256, 384, generator=gen) * 1.5).to(dev)
        pm_b, pz_b = sx.ops.score_pass1_batch(keys, q, n_img)
        ms, zs = [], []
        for i in range(nb):
            pm, pz = sx.ops.score_pass1(keys, q[i, :n_img].contiguous(), sx.ops.SCORE_TC)
            torch.cuda.synchronize()
            assert torch.equal(pm_b[i * parts:(i + 1) * parts, :n_img], pm[:, :n_img]), (nb, i)
            assert torch.equal(pz_b[i * parts:(i + 1) * parts, :n_img], pz[:, :n_img]), (nb, i)
            m, z = sx.ops.score_merge(pm, pz, n_img)
            ms.append(m)
            zs.append(z)
        scores_b = sx.ops.score_pass2_batch(keys, q, torch.stack(ms), torch.stack(zs), n_img)
        for i in range(nb):
            s, _ = sx.ops.score_pass2(keys, q[i, :n_img].contiguous(), ms[i], zs[i], sx.ops.SCORE_TC)
            torch.cuda.synchronize()
            assert torch.equal(scores_b[i], s), (nb, i)
        assert abs(scores_b.double().sum(1) - n_img).max().item() < 1e-2 * n_img


def test_multi_query_pipeline_equals_per_query_pipeline(sx, synthetic):
    """ShardedPoseEstimator(multi_query=True) end to end on one GPU: same poses as the per-query pipeline"""
    from conftest import load_golden
    dev = "cuda"
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    ori, dirs, rgb = r["ori"].to(dev), r["dirs"].to(dev), r["rgb"].to(dev)
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone(), score_impl="tc_bf16")
    idm.load_state_dict(synthetic.synth_id_weights(seed=g["weight_seed"]), strict=False)
    idm = idm.to(dev).eval().requires_grad_(False)
    img = g["img"].to(dev)
    imgs = torch.stack((img, img.flip(1), img * 0.7))
    masks = torch.stack((torch.ones(64, 64, dtype=torch.bool, device=dev), g["mask2"].to(dev),
                         torch.ones(64, 64, dtype=torch.bool, device=dev)))
    cache = idm.build_key_cache(ori, dirs, rgb)
    ref, _ = sx.ShardedPoseEstimator(idm, ori, dirs, cache, multi_query=False).query_batch(imgs, masks)
    est = sx.ShardedPoseEstimator(idm, ori, dirs, cache)
    assert est.multi_query  # the default on the tensor-core path
    out, _ = est.query_batch(imgs, masks)
    torch.testing.assert_close(out, ref, rtol=0, atol=0)
    assert est.enable_cuda_graphs(imgs, masks)
    out_g, _ = est.query_batch(imgs, masks)
    torch.testing.assert_close(out_g, ref, rtol=0, atol=0)


def test_topk_heavy_ties_signed_zero_and_specials(sx):
    """values always equal torch.topk; ties at the k-th key resolve to the LOWEST indices deterministically, also when
    there are more ties than the kernel's tie table (4096) holds; -0.0 and +0.0 are one value"""
    dev = "cuda"
    gen = torch.Generator().manual_seed(5)
    cases = []
    for n in (4097, 5000, 100_000, 3_600_000):
        cases.append(torch.rand(n, generator=gen))
        cases.append(torch.randn(n, generator=gen) * 1e-3)
        cases.append(torch.randint(0, 7, (n,), generator=gen).float())          # heavy ties (>> 4096 per value)
        x = torch.randn(n, generator=gen)
        x[::1001] = float("inf")
        x[5::1003] = -float("inf")
        x[7::1009] = 1e-42                                                        # denormals
        cases.append(x)
        z = torch.zeros(n)                                                        # an (almost) all-masked query
        z[n - 50:] = 1.0
        z[1::2] = -0.0
        cases.append(z)
    for x in cases:
        xd = x.to(dev)
        for k in (1, 100, 1024):
            v, i = sx.ops.topk(xd, k)
            v2, i2 = sx.ops.topk(xd, k)
            torch.cuda.synchronize()
            assert torch.equal(v, torch.topk(xd, k).values), (x.shape[0], k)
            assert torch.equal(i, i2) and torch.equal(xd[i], v)
            # contract: sorted by (value desc, index asc) over the whole input
            order = torch.sort(x, descending=True, stable=True).indices[:k]  # stable: equal values keep index order
            assert torch.equal(i.cpu(), order), (x.shape[0], k)


def test_no_grad_forward_twice_does_not_reuse_a_stale_key_cache(sx, synthetic):
    """forward() gathers a fresh random ray subset per call; the implicit key cache must follow the
    tensors it was built from, not a recycled address"""
    from conftest import load_golden
    dev = "cuda"
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    ori, dirs, rgb = r["ori"].to(dev), r["dirs"].to(dev), r["rgb"].to(dev)
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone(), score_impl="simt_fp32")
    idm.load_state_dict(synthetic.synth_id_weights(seed=g["weight_seed"]), strict=False)
    idm = idm.to(dev).eval().requires_grad_(False)
    img = g["img"].to(dev)
    mask = torch.ones(64, 64, dtype=torch.bool, device=dev)
    with torch.no_grad():
        for _ in range(4):
            scores, _, _, _, used = idm(img, mask, ori, dirs, rgb, rays_to_test=2000)
            o, d, c = ori[used], dirs[used], rgb[used]
            want, _, _, _ = idm.run_attention(img, mask, o, d, c)
            torch.testing.assert_close(scores, want, rtol=1e-5, atol=0)
            ref = idm.score_tokens(idm.backbone_wrapper(img, mask)[0], idm.build_key_cache(o, d, c))[0]
            torch.testing.assert_close(scores, ref, rtol=1e-5, atol=0)
    # the same three tensors again hit the cache; an in-place update invalidates it
    k1 = idm._cache_for(ori, dirs, rgb)
    assert idm._cache_for(ori, dirs, rgb) is k1
    ori.add_(0.0)
    assert idm._cache_for(ori, dirs, rgb) is not k1


def test_generate_rays_heavy_tail_vs_reference_fixture(sx, synthetic):
    """GPU ray generation against the reference-generated heavy-tailed fixture (tests/golden/rays_heavy.npz: 40-6444
    cells per ellipsoid, the load-imbalance case; the oracle is pinned to it on the CPU)"""
    from conftest import load_golden
    g = load_golden("rays_heavy.npz")
    sc = synthetic.synth_scene(int(g["scene_n"]), seed=int(g["scene_seed"]), heavy_tail=True)
    sc["scaling"] = g["scaling"]
    scene = sx.GaussianScene.from_dict(sc, device="cuda")
    ori, dirs, rgb = sx.generate_all_possible_rays(scene, ellipsoid_idx=g["perm"])
    n_ref = int(g["n_rays"])
    assert abs(ori.shape[0] - n_ref) <= max(3, int(0.002 * n_ref))
    sums = torch.stack((ori.double().sum(0), dirs.double().sum(0), rgb.double().sum(0))).cpu()
    assert torch.allclose(sums, g["sums"], rtol=2e-3, atol=0.002 * n_ref)
    if ori.shape[0] == n_ref:
        for mine, ref in ((ori, g["ori_s"]), (dirs, g["dirs_s"]), (rgb, g["rgb_s"])):
            frac = ((mine[::8].cpu() - ref).abs().max(dim=1).values <= 1e-5).float().mean().item()
            assert frac >= 0.985, frac


def test_weighted_least_squares_solve_mode_single_gpu_and_two_shards(sx, synthetic, oracle):
    """ShardedPoseEstimator(solve="weighted_ls"): all-ray weighted LS fused into the pass-2 epilogue, against the oracle's
    compute_line_intersection_impl2(ori, -dir, score / n_img) (least_squared_loss.py:62-64) on the kernel's own scores;
    two emulated shards (systems summed, as the all-reduce does) give the unsharded pose; CUDA graphs replay it."""
    from conftest import load_golden
    dev = "cuda"
    g, r = load_golden("id_module_peaked.npz"), load_golden("rays_small.npz")
    ori, dirs, rgb = r["ori"].to(dev), r["dirs"].to(dev), r["rgb"].to(dev)
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone(), score_impl="tc_f16x2")
    idm.load_state_dict(synthetic.synth_id_weights(seed=3, q_gain=float(g["q_gain"])), strict=False)
    idm = idm.to(dev).eval().requires_grad_(False)
    img = g["img"].to(dev)
    imgs = torch.stack((img, img.flip(1), img * 0.7))
    masks = torch.ones(3, 64, 64, dtype=torch.bool, device=dev)
    cache = idm.build_key_cache(ori, dirs, rgb)
    est = sx.ShardedPoseEstimator(idm, ori, dirs, cache, solve="weighted_ls")
    c2w, aux = est.query_batch(imgs, masks)
    for i in range(3):
        _, _, scores, up, _ = idm.test_image(imgs[i], masks[i], ori, dirs, rgb)
        w = scores.cpu() / 256
        centre = oracle.line_intersection(r["ori"].double(), -r["dirs"].double(), w.double()).float()
        watch = torch.nn.functional.normalize((w[:, None].double() * r["dirs"].double()).sum(0), dim=0).float()
        torch.testing.assert_close(c2w[i, :3, 3].cpu(), centre, rtol=1e-4, atol=1e-4)
        rot = torch.linalg.inv(oracle.make_rotation_mat(-watch, up.cpu()))
        torch.testing.assert_close(c2w[i, :3, :3].cpu(), rot, rtol=1e-4, atol=1e-4)
        assert aux[i, 7].item() == 0
    # two shards: per-shard systems add up to the whole (what the single all-reduce computes)
    n = ori.shape[0]
    cut = n // 2 + 11
    st = est._stage1(imgs, masks)
    sys_full = est._stage2_weighted(st["pmz"], st)
    shards = []
    for rank, (lo, hi) in enumerate(((0, cut), (cut, n))):
        o, d, c = ori[lo:hi].contiguous(), dirs[lo:hi].contiguous(), rgb[lo:hi].contiguous()
        shards.append(sx.ShardedPoseEstimator(idm, o, d, idm.build_key_cache(o, d, c), rank, 2, solve="weighted_ls"))
    sts = [s._stage1(imgs, masks) for s in shards]
    pmz = torch.cat([s_["pmz"] for s_ in sts])
    sys_sum = sum(s._stage2_weighted(pmz, s_) for s, s_ in zip(shards, sts))
    torch.testing.assert_close(sys_sum, sys_full, rtol=1e-6, atol=1e-5)
    c2w2, _ = shards[0]._stage3_weighted(sys_sum, sts[0])
    torch.testing.assert_close(c2w2, c2w, rtol=1e-5, atol=1e-5)
    assert est.enable_cuda_graphs(imgs, masks)
    c2w_g, _ = est.query_batch(imgs, masks)
    torch.testing.assert_close(c2w_g, c2w, rtol=0, atol=0)


@pytest.mark.parametrize("stored_deg,active_deg", [(0, 0), (1, 1), (2, 2), (3, 1), (3, 3)])
def test_ray_colours_for_low_sh_degrees(sx, synthetic, oracle, stored_deg, active_deg):
    """SH features are [N, (deg+1)^2, 3]; the fill kernel must stride by the STORED coefficient
    count and evaluate the ACTIVE degree (reference: get_features + eval_sh(active_sh_degree), sampling.py:116-124,236-251)"""
    sc = synthetic.synth_scene(150, seed=31)
    nc = (stored_deg + 1) ** 2
    sc["features_rest"] = sc["features_rest"][:, :nc - 1].contiguous()
    sc["sh_degree"] = active_deg
    scene = sx.GaussianScene.from_dict(sc, device="cuda")
    assert scene.max_sh_degree == stored_deg and scene.get_features.shape[1] == nc
    valid = oracle.mask_degraded_ellipsoids(*torch.exp(sc["scaling"]).unbind(-1))
    idx = torch.arange(int(valid.sum()))
    ori, dirs, rgb = sx.generate_all_possible_rays(scene, ellipsoid_idx=idx)
    feats = torch.cat((sc["features_dc"], sc["features_rest"]), 1)
    o_ref, d_ref, c_ref = oracle.generate_rays(sc["xyz"], sc["scaling"], sc["rotation"], feats, sh_degree=active_deg,
                                               ellipsoid_idx=idx)
    assert ori.shape[0] == o_ref.shape[0]
    # colour of OUR rays evaluated by the oracle's SH on our directions: isolates the SH stage from cell-level flips
    frac = ((rgb.cpu() - c_ref).abs().max(dim=1).values <= 1e-5).float().mean().item()
    assert frac >= 0.985, frac
    with pytest.raises(ValueError):
        sx.GaussianScene(sc["xyz"], sc["scaling"], sc["rotation"], sc["features_dc"], sc["features_rest"], sh_degree=stored_deg + 1)


def test_scene_loaded_from_the_reference_ply_generates_the_same_rays(sx, synthetic):
    """PLY written by the reference's save_ply -> load_ply -> rays == rays of the scene built from the same tensors"""
    import os
    from conftest import GOLDEN, load_golden
    g = load_golden("ply_ref.npz")
    a = sx.GaussianScene.load_ply(os.path.join(GOLDEN, "point_cloud_ref.ply"), device="cuda")
    b = sx.GaussianScene(g["xyz"], g["scaling"], g["rotation"], g["features_dc"], g["features_rest"], 3, device="cuda")
    ra = sx.generate_all_possible_rays(a, max_ellipsoids=None)
    rb = sx.generate_all_possible_rays(b, max_ellipsoids=None)
    assert ra[0].shape[0] > 1000
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)


def test_offline_eval_driver_end_to_end(sx, synthetic, tmp_path):
    """tools/eval_pose.py on a 3DGS-style experiment directory: discovers the PLY written by the reference's save_ply,
    reads cameras.json / PNGs / id_module.th, runs test_pose_estimation and writes the reference's JSON schema; the ground
    truth poses are the fixture's (reference test_pose_estimation), the predictions are finite rigid poses"""
    import importlib
    import json
    from test_abi_and_host import _write_experiment
    drv = importlib.import_module("6dgs_b200.eval_driver")
    exp, img_dir, p = _write_experiment(tmp_path, synthetic)
    out = tmp_path / "results.json"
    res = drv.main(["--exp_path", str(exp), "--images", str(img_dir), "--out", str(out), "--every", "1",
                    "--backbone", "synthetic", "--max_ellipsoids", "0"])
    saved = json.load(open(out))
    assert len(saved["results"]) == 3 and saved["trained_weights"] is True and saved["n_rays"] == res["n_rays"] > 1000
    gt = torch.tensor([r["gt_c2w"] for r in saved["results"]])
    torch.testing.assert_close(gt, p["gt_c2w"], rtol=1e-4, atol=1e-4)
    pred = torch.tensor([r["pred_c2w"] for r in saved["results"]])
    assert torch.isfinite(pred).all()
    rot = pred[:, :3, :3]
    torch.testing.assert_close(rot @ rot.transpose(1, 2), torch.eye(3).expand(3, 3, 3), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("n,n_img,q_scale", [(1, 256, 1.0), (1000, 201, 1.0), (3017, 256, 8.0), (70_003, 256, 3.0)])
def test_score_backward_kernels_vs_fp64_autograd(sx, n, n_img, q_scale):
    """d(sum_i softmax_r(q k^T / sqrt(384))) / d(q, k) from the backward kernels (two streaming passes + two GEMMs)
    against torch autograd in fp64 (what train.py:176 `loss.backward()` computes through the reference's attention)"""
    from importlib import import_module
    fn = import_module("6dgs_b200.identification")._RayScoreFunction
    dev = "cuda"
    gen = torch.Generator().manual_seed(n)
    q = (torch.randn(n_img, 384, generator=gen) * q_scale).to(dev).requires_grad_(True)
    k = (torch.randn(n, 384, generator=gen) * 0.7).to(dev).requires_grad_(True)
    g = torch.randn(n, generator=gen).to(dev)
    scores, amap = fn.apply(q, k, n <= 3017)
    (scores * g).sum().backward()
    q64, k64 = q.detach().double().requires_grad_(True), k.detach().double().requires_grad_(True)
    a64 = torch.softmax((q64 @ k64.t()) / 384 ** 0.5, dim=-1)
    (a64.sum(0) * g.double()).sum().backward()
    torch.testing.assert_close(scores.detach().double(), a64.sum(0).detach(), rtol=1e-3, atol=0)
    if n <= 3017:
        torch.testing.assert_close(amap.double(), a64.detach(), rtol=1e-3, atol=1e-12)
    for mine, ref, name in ((q.grad, q64.grad, "dq"), (k.grad, k64.grad, "dk")):
        err = (mine.double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        assert err < 2e-4, (name, err)


def test_training_step_gradients_kernels_vs_torch_route(sx, synthetic, monkeypatch):
    """forward() under autograd (train_id_module's call, train.py:146-176): parameter gradients with the score
    forward/backward on the kernels equal those of the pure torch-op route"""
    from conftest import load_golden
    dev = "cuda"
    r, g = load_golden("rays_small.npz"), load_golden("id_module.npz")
    ori, dirs, rgb = r["ori"][:3000].to(dev), r["dirs"][:3000].to(dev), r["rgb"][:3000].to(dev)
    img, mask = g["img"].to(dev), torch.ones(64, 64, dtype=torch.bool, device=dev)
    target = torch.rand(2500, generator=torch.Generator().manual_seed(1)).to(dev) * 0.1
    grads = {}
    for route in ("kernels", "torch"):
        monkeypatch.setenv("SIXDGS_TRAIN_SCORE", route)
        idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone(), score_impl="simt_fp32")
        idm.load_state_dict(synthetic.synth_id_weights(seed=3, q_gain=8.0), strict=False)
        idm = idm.to(dev).train()
        torch.manual_seed(0)
        scores, amap, tok, up, used = idm(img, mask, ori, dirs, rgb, rays_to_test=2500)
        assert amap.shape == (256, 2500)
        loss = torch.square(scores - target).mean() + 0.1 * (0.5 - 0.5 * up[2])  # train.py:161-170 shape of the loss
        loss.backward()
        grads[route] = {n_: p.grad.clone() for n_, p in idm.named_parameters()
                        if p.grad is not None and n_.startswith(("ray_preprocessor", "attention"))}
    assert set(grads["kernels"]) == set(grads["torch"]) and len(grads["torch"]) == 12
    # (the softmax over rays is invariant to a constant added to every key, so the gradients of mlp2.2.bias and
    # k_proj.bias are exactly zero in exact arithmetic -- rounding noise in both routes: absolute floor below)
    floor = 1e-5 * max(gt.abs().max().item() for gt in grads["torch"].values())
    for n_, gt in grads["torch"].items():
        gk = grads["kernels"][n_]
        assert (gk - gt).abs().max().item() <= 2e-3 * gt.abs().max().item() + floor, n_


@pytest.mark.parametrize("impl,solve", [("simt_fp32", "topk"), ("tc_f16x2", "topk"), ("tc_f16x2", "weighted_ls")])
def test_shard_without_rays_is_neutral_on_the_cuda_backend(sx, synthetic, impl, solve):
    """ragged sharding on the CUDA backend (two shards emulated in one process, the second one EMPTY): neutral statistics,
    all -inf candidate rows, a zero least-squares system -- the poses equal the unsharded ones; the empty shard declines
    CUDA-graph capture"""
    from conftest import load_golden
    dev = "cuda"
    g, r = load_golden("id_module.npz"), load_golden("rays_small.npz")
    ori, dirs, rgb = r["ori"].to(dev), r["dirs"].to(dev), r["rgb"].to(dev)
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone(), score_impl=impl)
    idm.load_state_dict(synthetic.synth_id_weights(seed=g["weight_seed"]), strict=False)
    idm = idm.to(dev).eval().requires_grad_(False)
    img = g["img"].to(dev)
    imgs = torch.stack((img, img.flip(1)))
    masks = torch.ones(2, 64, 64, dtype=torch.bool, device=dev)
    ref, _ = sx.ShardedPoseEstimator(idm, ori, dirs, idm.build_key_cache(ori, dirs, rgb), solve=solve).query_batch(imgs, masks)
    e3 = torch.empty(0, 3, device=dev)
    shards = [sx.ShardedPoseEstimator(idm, ori, dirs, idm.build_key_cache(ori, dirs, rgb), 0, 2, solve=solve),
              sx.ShardedPoseEstimator(idm, e3, e3.clone(), idm.build_key_cache(e3, e3.clone(), e3.clone()), 1, 2, solve=solve)]
    assert shards[1].cache.n_rays == 0 and not shards[1].enable_cuda_graphs(imgs, masks)
    sts = [s._stage1(imgs, masks) for s in shards]
    assert sts[0]["pmz"].shape == sts[1]["pmz"].shape
    pmz = torch.cat([st["pmz"] for st in sts])
    if solve == "weighted_ls":
        sys_sum = sum(s._stage2_weighted(pmz, st) for s, st in zip(shards, sts))
        for s, st in zip(shards, sts):
            c2w, _ = s._stage3_weighted(sys_sum, st)
            torch.testing.assert_close(c2w, ref, rtol=1e-5, atol=1e-5)
    else:
        cands = [s._stage2(pmz, st, 100)[2] for s, st in zip(shards, sts)]
        assert torch.isinf(cands[1][..., 0]).all() and (cands[1][..., 0] < 0).all()
        allc = torch.cat(cands)
        for s, st in zip(shards, sts):
            c2w, _ = s._stage3(allc, st["up"], 100, st["nb"])
            torch.testing.assert_close(c2w, ref, rtol=1e-5, atol=1e-5)
