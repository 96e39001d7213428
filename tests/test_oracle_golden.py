"""Pin the CPU oracle (oracle/sixdgs_oracle.py) against fixtures produced by the unmodified
reference (oracle/gen_golden.py).  CPU only."""
import torch

from conftest import load_golden


def test_degrade_mask(oracle):
    g = load_golden("quadricell.npz")
    s = g["mask_scales"]
    valid = oracle.mask_degraded_ellipsoids(s[:, 0], s[:, 1], s[:, 2])
    assert 0 < int(valid.sum()) < valid.numel()  # fixture exercises both branches
    assert torch.equal(valid, g["mask_valid"])


def test_quadricell_centers(oracle):
    g = load_golden("quadricell.npz")
    abc = g["abc"]
    pts, eid = oracle.quadricell_centers(abc[:, 0], abc[:, 1], abc[:, 2], 50)
    assert torch.equal(eid, g["ellipsoid_id"])
    assert torch.equal(pts, g["points"])  # same torch ops in the same order: bit exact on CPU


def test_sym_eig(oracle):
    g = load_golden("sym_eig.npz")
    vals, vecs = oracle.sym_eig_3x3(g["A"])
    torch.testing.assert_close(vals, g["vals"], rtol=0, atol=0)
    torch.testing.assert_close(vecs, g["vecs"], rtol=0, atol=0)
    try:
        oracle.sym_eig_3x3(torch.zeros(2, 2))
        assert False
    except ValueError:
        pass


def test_normals(oracle):
    g = load_golden("normals.npz")
    n = oracle.knn_normals(g["cloud"][:300], g["cloud"], 20)
    torch.testing.assert_close(n, g["normals"], rtol=0, atol=0)


def _scene(g):
    return dict(xyz=g["xyz"], scaling_raw=g["scaling"], rotation_raw=g["rotation"],
                features=torch.cat((g["features_dc"], g["features_rest"]), 1))


def test_generate_rays_small(oracle):
    g = load_golden("rays_small.npz")
    ori, dirs, rgb = oracle.generate_rays(**_scene(g), ellipsoid_idx=g["perm"])
    assert ori.shape == g["ori"].shape
    torch.testing.assert_close(ori, g["ori"], rtol=0, atol=0)
    torch.testing.assert_close(dirs, g["dirs"], rtol=0, atol=0)
    torch.testing.assert_close(rgb, g["rgb"], rtol=0, atol=1e-7)


def test_generate_rays_capped(oracle, synthetic):
    g = load_golden("rays_capped.npz")
    sc = synthetic.synth_scene(g["scene_n"], seed=g["scene_seed"])
    ori, dirs, rgb = oracle.generate_rays(sc["xyz"], sc["scaling"], sc["rotation"],
                                          torch.cat((sc["features_dc"], sc["features_rest"]), 1),
                                          ellipsoid_idx=g["perm"])
    assert ori.shape[0] == g["n_rays"]
    torch.testing.assert_close(ori[::16], g["ori_s"], rtol=0, atol=0)
    torch.testing.assert_close(dirs[::16], g["dirs_s"], rtol=0, atol=0)
    torch.testing.assert_close(rgb[::16], g["rgb_s"], rtol=0, atol=1e-7)


def test_generate_rays_heavy_tail(oracle, synthetic):
    """Log-normal scales: ~10x spread of cells per ellipsoid and a share of degraded (needle) ellipsoids."""
    g = load_golden("rays_heavy.npz")
    sc = synthetic.synth_scene(g["scene_n"], seed=g["scene_seed"], heavy_tail=True)
    sc["scaling"] = g["scaling"]
    assert int(g["n_valid"]) < int(g["scene_n"])  # the fixture does exercise the degrade mask
    ori, dirs, rgb = oracle.generate_rays(sc["xyz"], sc["scaling"], sc["rotation"],
                                          torch.cat((sc["features_dc"], sc["features_rest"]), 1),
                                          ellipsoid_idx=g["perm"])
    assert ori.shape[0] == g["n_rays"]
    torch.testing.assert_close(ori[::8], g["ori_s"], rtol=0, atol=0)
    torch.testing.assert_close(dirs[::8], g["dirs_s"], rtol=0, atol=0)
    torch.testing.assert_close(rgb[::8], g["rgb_s"], rtol=0, atol=1e-7)


def test_ray_features_and_scores(oracle, synthetic):
    g = load_golden("id_module.npz")
    r = load_golden("rays_small.npz")
    w = synthetic.synth_id_weights(seed=g["weight_seed"])
    chk = torch.stack([v.double().abs().sum() for _, v in sorted(w.items())])
    torch.testing.assert_close(chk, g["weight_checksum"], rtol=1e-12, atol=0)  # RNG drift guard
    fea = oracle.ray_features(r["ori"], r["dirs"], r["rgb"], w)
    torch.testing.assert_close(fea[g["fea_sel"]], g["fea"], rtol=1e-5, atol=1e-5)
    score, A = oracle.attention_scores(g["tok_pe"], fea, w)
    torch.testing.assert_close(A[[0, 100, 255]], g["A_rows"], rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(score, g["scores"], rtol=1e-4, atol=1e-8)
    assert abs(float(score.sum()) - 256.0) < 1e-2
    top = torch.topk(score, 100)
    assert set(top.indices.tolist()) == set(g["topk_idx"].tolist())
    sc2, m, z = oracle.attention_scores_chunked(g["tok_pe"], lambda lo, hi: fea[lo:hi], fea.shape[0], w, chunk=1000)
    torch.testing.assert_close(sc2, g["scores"], rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(m, g["row_max"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(m + torch.log(z), g["row_lse"], rtol=1e-5, atol=1e-5)


def test_line_intersection(oracle):
    g = load_golden("line_intersection.npz")
    c = oracle.line_intersection(g["o"], g["d"])
    torch.testing.assert_close(c, g["c_unweighted"], rtol=1e-6, atol=1e-6)
    cw = oracle.line_intersection(g["o"], g["d"], g["w"])
    torch.testing.assert_close(cw, g["c_weighted"], rtol=1e-6, atol=1e-6)
    assert torch.isnan(oracle.line_intersection(g["o_par"], g["d_par"])).all()
    assert torch.isnan(g["c_parallel"]).all()
    assert torch.equal(oracle.exclude_negatives(c, g["o"], g["d"]), g["neg_mask"])
    torch.testing.assert_close(oracle.make_rotation_mat(g["rot_dir"], g["rot_up"]), g["rot"], rtol=1e-6, atol=1e-7)


def test_peaked_softmax_fixture(oracle, synthetic, sx):
    """The oracle against the reference on the PEAKED fixture (attention.q_proj x20: logit std 6.5): scores, row
    statistics, top-100 and the full pose of test_pose_estimation.  Also the arithmetic that decides the score format of
    the GPU path: the same logits with bf16 / fp16 operand rounding (what a single tensor-core term computes) miss the
    north_star bar, the fp16 hi+lo three-term form does not."""
    g = load_golden("id_module_peaked.npz")
    r = load_golden("rays_small.npz")
    i = load_golden("id_module.npz")
    w = synthetic.synth_id_weights(seed=g["weight_seed"], q_gain=float(g["q_gain"]))
    fea = oracle.ray_features(r["ori"], r["dirs"], r["rgb"], w)
    score, _ = oracle.attention_scores(i["tok_pe"], fea, w, return_map=False)
    torch.testing.assert_close(score, g["scores"], rtol=2e-4, atol=0)
    top = torch.topk(score, 100)
    assert set(top.indices.tolist()) == set(g["topk_idx"].tolist())
    # the pose of the fixture comes from test_pose_estimation on the uint8 image (test.py:69-73 divides by 255): same
    # route here -- torch front end (boundary components, CPU) -> oracle scores -> oracle pose tail
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone())
    idm.load_state_dict(w, strict=False)
    with torch.no_grad():
        tok_pe, _, grid = idm.backbone_wrapper(g["img_u8"].float() / 255.0, torch.ones(64, 64, dtype=torch.bool))
        up = idm._camera_up(grid)
    s8, _ = oracle.attention_scores(tok_pe, fea, w, return_map=False)
    t8 = torch.topk(s8, 100)
    c2w, _ = oracle.pose_tail(t8.indices, t8.values, r["ori"], r["dirs"], up)
    torch.testing.assert_close(c2w, g["pred_c2w"], rtol=1e-4, atol=1e-4)
    # operand-rounding study in fp64 (exact products and sums, only the operands are rounded)
    q = torch.nn.functional.linear(i["tok_pe"], w["attention.q_proj.weight"], w["attention.q_proj.bias"]).double()
    k = torch.nn.functional.linear(fea, w["attention.k_proj.weight"], w["attention.k_proj.bias"]).double()
    ref = torch.softmax(q @ k.t() / 384 ** 0.5, -1).sum(0)

    def err(logits):
        s = torch.softmax(logits / 384 ** 0.5, -1).sum(0)
        return ((s - ref).abs() / ref).max().item()

    def split(x, dt):
        hi = x.float().to(dt).double()
        return hi, (x - hi).float().to(dt).double()

    e_bf16 = err(q.float().to(torch.bfloat16).double() @ k.float().to(torch.bfloat16).double().t())
    e_fp16 = err(q.float().half().double() @ k.float().half().double().t())
    qh, ql = split(q, torch.float16)
    kh, kl = split(k, torch.float16)
    e_x2 = err(qh @ kh.t() + qh @ kl.t() + ql @ kh.t())
    assert e_bf16 > 1e-2 and e_fp16 > 2e-3 and e_x2 < 1e-5, (e_bf16, e_fp16, e_x2)
    # the fast variant (tc_f16f8) with the kernel's own scales: stored key 16 k, stored query 64 c q (c = log2e / sqrt(384)
    # cancels in this study), cross-term operands e4m3(hi / 64) and e4m3(64 lo) -- products land on the main term's scale
    def f8(x):
        return x.float().to(torch.float8_e4m3fn).double()

    Qh, Ql = split(q * 64.0, torch.float16)
    Kh, Kl = split(k * 16.0, torch.float16)
    acc = Qh @ Kh.t() + f8(Qh / 64.0) @ f8(Kl * 64.0).t() + f8(Ql * 64.0) @ f8(Kh / 64.0).t()
    e_f8 = err(acc / 1024.0)
    assert e_x2 < e_f8 < 1e-3, e_f8          # inside the bar here (logit std 6.5), 0.05 x the fp16 error, no margin at 3 x the spread
    assert e_f8 < 0.1 * e_fp16
