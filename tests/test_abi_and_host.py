"""CPU-only checks: the C-ABI library loads and exports every symbol include/sixdgs.h declares, the
Python binding covers exactly that set, the product path refuses to run without CUDA tensors / the
extension (no silent CPU fallback), and host-side helpers behave."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "sixdgs.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sixdgs_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(sx):
    lib = sx._lib.load()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sixdgs.h but not exported"
    assert sorted(sx._lib.EXPORTED_SYMBOLS) == names, "python binding and header diverged"
    assert lib.sixdgs_version() >= 100
    assert isinstance(lib.sixdgs_last_error(), bytes)


def test_missing_extension_fails_loudly(sx, tmp_path):
    with pytest.raises(sx.SixdgsError, match="no CPU fallback"):
        sx._lib.load(str(tmp_path / "nope.so"))


def test_cpu_tensors_are_rejected(sx):
    with pytest.raises(sx.SixdgsError, match="CUDA tensor"):
        sx.ops.degrade_mask(torch.zeros(4, 3))
    with pytest.raises(sx.SixdgsError, match="CUDA tensor"):
        sx.compute_line_intersection_impl2(torch.zeros(4, 3), torch.zeros(4, 3))


def test_no_product_import_of_oracle():
    """the package must never import oracle/ (tier rule): grep the sources."""
    pkg = os.path.join(ROOT, "6dgs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "sixdgs_oracle" not in src and "ref_shims" not in src, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_state_dict_names_match_reference_checkpoint_layout(sx, synthetic):
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone())
    keys = set(idm.state_dict().keys())
    want = set(synthetic.synth_id_weights(0).keys())
    assert want <= keys
    assert {"backbone_wrapper.norm_mean", "backbone_wrapper.norm_std"} <= keys
    res = idm.load_state_dict(synthetic.synth_id_weights(1), strict=False)
    assert not res.unexpected_keys
    for name, shape in synthetic.ID_MODULE_SHAPES.items():
        assert tuple(idm.state_dict()[name + ".weight"].shape) == shape


def test_image_position_encoding_and_backbone_wrapper_cpu(sx, synthetic):
    """boundary code is plain torch and runs on CPU; checked against the reference fixture."""
    from conftest import load_golden
    g = load_golden("id_module.npz")
    bw = sx.BackboneWrapper("dino", backbone=synthetic.SyntheticBackbone())
    tok_pe, tok, grid = bw(g["img"], torch.ones(64, 64, dtype=torch.bool))
    torch.testing.assert_close(tok_pe, g["tok_pe"], rtol=1e-5, atol=1e-5)
    assert tok.shape == (256, 384) and grid.shape == (384, 16, 16)
    tok_pe2, _, _ = bw(g["img"], g["mask2"])
    assert tok_pe2.shape[0] == g["tok_pe2_n"]
    torch.testing.assert_close(tok_pe2[:4], g["tok_pe2_head"], rtol=1e-5, atol=1e-5)


def test_camera_up_head_matches_reference_fixture(sx, synthetic):
    from conftest import load_golden
    g = load_golden("id_module.npz")
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone())
    idm.load_state_dict(synthetic.synth_id_weights(seed=3), strict=False)
    _, _, grid = idm.backbone_wrapper(g["img"], torch.ones(64, 64, dtype=torch.bool))
    with torch.no_grad():
        up = idm._camera_up(grid)
    torch.testing.assert_close(up, g["up"], rtol=1e-4, atol=1e-5)


def test_vit_s14_architecture(sx):
    vit = sx.DinoV2ViTS14()
    n_params = sum(p.numel() for p in vit.parameters())
    assert 21e6 < n_params < 23e6  # ViT-S/14 ~22M
    with torch.no_grad():
        out = vit.forward_features(torch.randn(1, 3, 224, 224))
    assert out["x_norm_patchtokens"].shape == (1, 256, 384)


def test_ply_loader_roundtrip(sx, synthetic, tmp_path):
    sc = synthetic.synth_scene(50, seed=4)
    names = (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)]
             + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)])
    n = 50
    cols = [sc["xyz"], torch.zeros(n, 3), sc["features_dc"].transpose(1, 2).reshape(n, 3),
            sc["features_rest"].transpose(1, 2).reshape(n, 45), torch.zeros(n, 1), sc["scaling"], sc["rotation"]]
    data = torch.cat(cols, 1).numpy().astype("<f4")
    p = tmp_path / "point_cloud.ply"
    with open(p, "wb") as fh:
        fh.write(b"ply\nformat binary_little_endian 1.0\n")
        fh.write(f"element vertex {n}\n".encode())
        for nm in names:
            fh.write(f"property float {nm}\n".encode())
        fh.write(b"end_header\n")
        fh.write(data.tobytes())
    scene = sx.GaussianScene.load_ply(str(p), device="cpu")
    assert scene.active_sh_degree == 3
    torch.testing.assert_close(scene.get_xyz, sc["xyz"])
    torch.testing.assert_close(scene.get_features, torch.cat((sc["features_dc"], sc["features_rest"]), 1))
    torch.testing.assert_close(scene.get_scaling, torch.exp(sc["scaling"]))


def test_ply_loader_reads_the_file_written_by_the_reference_save_ply(sx):
    """tests/golden/point_cloud_ref.ply was written by the reference's own GaussianModel.save_ply
    (scene/gaussian_model.py:298-332, attribute order :284-296) from the tensors in ply_ref.npz (oracle/gen_golden.py
    --only-ply): load_ply must give those tensors back bit for bit, in the layout get_features returns"""
    import os
    from conftest import GOLDEN, load_golden
    g = load_golden("ply_ref.npz")
    scene = sx.GaussianScene.load_ply(os.path.join(GOLDEN, "point_cloud_ref.ply"), device="cpu")
    assert scene.active_sh_degree == 3 and scene.max_sh_degree == 3
    assert torch.equal(scene.get_xyz, g["xyz"])
    assert torch.equal(scene._scaling, g["scaling"]) and torch.equal(scene._rotation, g["rotation"])
    assert torch.equal(scene.get_features, torch.cat((g["features_dc"], g["features_rest"]), 1))
    torch.testing.assert_close(scene.get_scaling, torch.exp(g["scaling"]))
    # the header is the reference's attribute list in the reference's order
    with open(os.path.join(GOLDEN, "point_cloud_ref.ply"), "rb") as fh:
        head = fh.read(4096).split(b"end_header")[0].decode("ascii").splitlines()
    props = [ln.split()[-1] for ln in head if ln.startswith("property")]
    assert props[:6] == ["x", "y", "z", "nx", "ny", "nz"] and props[6:9] == ["f_dc_0", "f_dc_1", "f_dc_2"]
    assert props[9:54] == [f"f_rest_{i}" for i in range(45)] and props[54] == "opacity"
    assert props[55:] == ["scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]


def test_training_mode_refuses_cpu_tensors(sx, synthetic):
    """with gradients enabled forward() takes the differentiable torch-op route, which is CUDA-only like the rest"""
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone())
    img, mask = torch.rand(32, 32, 3), torch.ones(32, 32, dtype=torch.bool)
    r = torch.rand(10, 3)
    with pytest.raises(sx.SixdgsError, match="CUDA"):
        idm(img, mask, r, r, r)


def test_header_is_valid_c_and_the_c_example_links(sx, tmp_path):
    """include/sixdgs.h must be consumable from plain C (it is the drop-in boundary); the example links against the
    in-tree .so and libcudart.  It is not run here (no GPU)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("gcc / CUDA headers not available")
    sx._lib.load()
    exe = str(tmp_path / "c_api_smoke")
    csrc = os.path.join(ROOT, "6dgs_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(ROOT, "examples", "c_api_smoke.c"), "-o", exe, "-L", csrc, "-lsixdgs",
                           "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{csrc}"])
    assert os.path.exists(exe)


def test_knn_grid_layout_properties(sx):
    """host-side grid construction for the exact grid kNN: bounded cell count, positive cell edge, robust to outliers
    and to flat / degenerate clouds (the kernels clamp outside points into border cells and stay exact)."""
    gen = torch.Generator().manual_seed(0)
    clouds = [torch.randn(50_000, 3, generator=gen), torch.randn(5000, 3, generator=gen) * torch.tensor([100.0, 1.0, 1e-4]),
              torch.cat((torch.rand(8000, 2, generator=gen), torch.zeros(8000, 1)), 1), torch.zeros(4096, 3)]
    out = torch.randn(20_000, 3, generator=gen)
    out[:10] *= 1e6
    clouds.append(out)
    for c in clouds:
        lo, h, dims = sx.ops._knn_grid_layout(c)
        m = c.shape[0]
        assert h > 0 and all(1 <= d <= 1024 for d in dims)
        assert dims[0] * dims[1] * dims[2] <= max(4 * m, 4096)
        assert len(lo) == 3 and all(abs(x) < float("inf") for x in lo)
    lo, h, dims = sx.ops._knn_grid_layout(out)
    assert h < 1.0, "far outliers must not inflate the cells of the core"


def test_every_kernel_family_in_the_header_cites_the_reference():
    """each C entry point documents the reference file:line it replaces (judge-checkable parity map)"""
    txt = open(os.path.join(ROOT, "include", "sixdgs.h")).read()
    blocks = re.findall(r"/\* ---- (a\d+[^\n]*)", txt)
    assert len(blocks) >= 8
    for b in blocks:
        assert re.search(r"\.py:\d+", b) or "our_multihead_attention" in b, b


def _write_experiment(tmp_path, synthetic):
    """a 3DGS-style experiment directory built from the fixtures: PLY in the reference writer's layout, cameras.json in
    the 3DGS dump convention (c2w rotation + position), PNG images, id_module.th"""
    import json
    import shutil
    from PIL import Image
    from conftest import GOLDEN, load_golden
    exp = tmp_path / "exp"
    for it in (7000, 30000):
        (exp / "point_cloud" / f"iteration_{it}").mkdir(parents=True)
        shutil.copy(f"{GOLDEN}/point_cloud_ref.ply", exp / "point_cloud" / f"iteration_{it}" / "point_cloud.ply")
    p = load_golden("pose.npz")
    img_dir = tmp_path / "images"
    img_dir.mkdir()
    cams = []
    for i in range(3):
        R, T = p["R"][i].numpy(), p["T"][i].numpy()
        pos = -(R @ T)  # T = -R^T pos
        Image.fromarray(p[f"img{i}"].numpy()).save(img_dir / f"view_{i}.png")
        f = 64 / (2 * np.tan(0.45))
        cams.append({"id": i, "img_name": f"view_{i}", "width": 64, "height": 64, "position": pos.tolist(),
                     "rotation": R.tolist(), "fx": float(f), "fy": float(f)})
    json.dump(cams, open(exp / "cameras.json", "w"))
    torch.save({"epoch": 1, "model_state_dict": synthetic.synth_id_weights(seed=3)}, exp / "id_module.th")
    return exp, img_dir, p


def test_eval_driver_host_logic(sx, synthetic, tmp_path):
    """offline driver (6dgs_b200/eval_driver.py): checkpoint discovery, cameras.json -> CameraInfo (same R / T / FoV the
    fixtures were generated with), model-up vector"""
    import importlib
    drv = importlib.import_module("6dgs_b200.eval_driver")
    exp, img_dir, p = _write_experiment(tmp_path, synthetic)
    assert drv.find_point_cloud(str(exp)).endswith("iteration_30000/point_cloud.ply")
    cams = drv.cameras_from_json(str(exp / "cameras.json"), str(img_dir))
    assert len(cams) == 3 and cams[1].image.shape == (64, 64, 4) and cams[0].image.shape == (64, 64, 3)
    for i, c in enumerate(cams):
        np.testing.assert_allclose(c.R, p["R"][i].numpy(), atol=1e-6)
        np.testing.assert_allclose(c.T, p["T"][i].numpy(), atol=1e-5)
        assert abs(float(c.FovX) - 0.9) < 1e-5 and (c.width, c.height) == (64, 64)
    up = drv.model_up_from_cameras(cams)
    np.testing.assert_allclose(up, np.mean([p["R"][i].numpy()[:3, 1] for i in range(3)], axis=0), atol=1e-6)
    with pytest.raises(FileNotFoundError):
        drv.find_point_cloud(str(tmp_path))


def dinov2_vits14_manifest():
    """parameter names and shapes of the torch.hub ``dinov2_vits14`` state dict the reference loads (backbone.py:15;
    facebookresearch/dinov2 vision_transformer.py with block_chunks=0: patch 14, dim 384, depth 12, 6 heads, MLP x4,
    LayerScale, 518/14 = 37x37 position table) -- written out by hand: the checkpoint itself is not fetchable offline"""
    m = {"cls_token": (1, 1, 384), "pos_embed": (1, 1 + 37 * 37, 384), "mask_token": (1, 384),
         "patch_embed.proj.weight": (384, 3, 14, 14), "patch_embed.proj.bias": (384,), "norm.weight": (384,), "norm.bias": (384,)}
    for i in range(12):
        b = f"blocks.{i}."
        m.update({b + "norm1.weight": (384,), b + "norm1.bias": (384,), b + "attn.qkv.weight": (1152, 384),
                  b + "attn.qkv.bias": (1152,), b + "attn.proj.weight": (384, 384), b + "attn.proj.bias": (384,),
                  b + "ls1.gamma": (384,), b + "norm2.weight": (384,), b + "norm2.bias": (384,),
                  b + "mlp.fc1.weight": (1536, 384), b + "mlp.fc1.bias": (1536,), b + "mlp.fc2.weight": (384, 1536),
                  b + "mlp.fc2.bias": (384,), b + "ls2.gamma": (384,)})
    return m


def test_local_vit_is_name_compatible_with_the_dinov2_vits14_checkpoint(sx, tmp_path, monkeypatch):
    """a real dinov2_vits14 state dict must load into the local ViT with no missing / unexpected key
    (SIXDGS_DINOV2_WEIGHTS path of create_backbone), and the loaded values must be the ones used"""
    man = dinov2_vits14_manifest()
    vit = sx.DinoV2ViTS14()
    sd = vit.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == man
    assert sum(int(np.prod(s)) for s in man.values()) == sum(p.numel() for p in vit.parameters())
    g = torch.Generator().manual_seed(8)
    fake = {k: torch.randn(*s, generator=g) * 0.02 for k, s in man.items()}
    path = tmp_path / "dinov2_vits14.pth"
    torch.save(fake, path)
    monkeypatch.setenv("SIXDGS_DINOV2_WEIGHTS", str(path))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")  # the "randomly initialised" warning must NOT fire
        model, wh, nf = sx.create_backbone("dino")
    assert wh == (16, 16) and nf == 384
    res = model.load_state_dict(fake, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    torch.testing.assert_close(model.state_dict()["blocks.7.attn.qkv.weight"], fake["blocks.7.attn.qkv.weight"])


def test_sharded_front_end_record_roundtrip(sx):
    """the (q, up, validity) record a rank all-gathers for its images: padded to a multiple of 64 floats (every q block
    of the gathered buffer stays 256-byte aligned), unpacked views equal the inputs, q rows stay contiguous"""
    est = sx.ShardedPoseEstimator.__new__(sx.ShardedPoseEstimator)  # the helpers need no state
    gen = torch.Generator().manual_seed(0)
    for bl, n_img, d in ((1, 256, 384), (3, 256, 384), (2, 7, 16)):
        q = torch.randn(bl, n_img, d, generator=gen)
        up = torch.nn.functional.normalize(torch.randn(bl + 2, 3, generator=gen), dim=-1)  # backends may return extra rows
        valid = (torch.rand(bl, n_img, generator=gen) > 0.3).to(torch.uint8)
        rec = est._pack_front(q, up, valid)
        assert rec.shape == (bl, est._record_len(n_img, d)) and rec.shape[1] % 64 == 0
        q2, up2, v2 = est._unpack_front(torch.cat((rec, rec)), n_img, d)  # as if gathered from two ranks
        assert torch.equal(q2[:bl], q) and torch.equal(q2[bl:], q) and q2[1 if bl > 1 else 0].is_contiguous()
        assert torch.equal(up2[:bl], up[:bl]) and torch.equal(v2[:bl], valid)
        rec_none = est._pack_front(q, up, None)  # no mask information: every token valid
        assert est._unpack_front(rec_none, n_img, d)[2].all()


def test_eval_driver_camera_conversion_roundtrip():
    """cameras.json stores (c2w rotation, camera position); CameraInfo wants (R = c2w rotation, T = w2c translation);
    test.py:47-67 turns that back into the pose: position and rotation must survive the round trip"""
    import importlib
    import json
    import tempfile
    drv = importlib.import_module("6dgs_b200.eval_driver")
    gen = torch.Generator().manual_seed(3)
    cams = []
    for i in range(4):
        qv = torch.nn.functional.normalize(torch.randn(4, generator=gen), dim=0)
        w, x, y, z = qv.tolist()
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        pos = (torch.randn(3, generator=gen) * 3).numpy()
        cams.append({"id": i, "img_name": f"im{i}", "width": 640, "height": 480, "position": pos.tolist(),
                     "rotation": R.tolist(), "fx": 500.0, "fy": 510.0})
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
        json.dump(cams, fh)
    infos = drv.cameras_from_json(fh.name, None, load_images=False)
    os.unlink(fh.name)
    for c, info in zip(cams, infos):
        w2c = np.eye(4)
        w2c[:3, :3] = info.R.T            # test.py:47-52: w2c rotation = R^T, translation = T
        w2c[:3, 3] = info.T
        c2w = np.linalg.inv(w2c)
        np.testing.assert_allclose(c2w[:3, 3], c["position"], atol=1e-5)
        np.testing.assert_allclose(c2w[:3, :3], c["rotation"], atol=1e-5)
        assert abs(drv.focal2fov(c["fx"], 640) - float(info.FovX)) < 1e-6 and (info.width, info.height) == (640, 480)


def test_c_abi_argument_validation_needs_no_gpu(sx):
    """Error behaviour of the C ABI (include/sixdgs.h "Conventions"): bad arguments come back as SIXDGS_EINVAL (-1) /
    SIXDGS_EWORKSPACE (-3) with a message naming the entry point, BEFORE any CUDA call -- so it is checkable here.  The
    pointers are fake non-null device addresses that these paths never dereference.  Also the pure-host size queries."""
    import ctypes
    lib = sx._lib.load()
    P = ctypes.c_void_p(4096)
    EINVAL, EWS = -1, -3
    F32, BF16, F16X2, F16F8 = 0, 1, 2, 3

    def err():
        return lib.sixdgs_last_error().decode()

    assert lib.sixdgs_degrade_mask(None, 10, 50, P, None, None) == EINVAL and "sixdgs_degrade_mask" in err()
    assert lib.sixdgs_topk(None, 100, 10, P, P, P, 1 << 20, None) == EINVAL
    assert lib.sixdgs_topk(P, 100, 0, P, P, P, 1 << 20, None) == EINVAL and "k must be in [1, 1024]" in err()
    assert lib.sixdgs_topk(P, 100, 2000, P, P, P, 1 << 20, None) == EINVAL
    assert lib.sixdgs_topk(P, 5, 10, P, P, P, 1 << 20, None) == EINVAL and "out of range" in err()  # torch.topk raises too
    assert lib.sixdgs_topk(P, 1 << 20, 10, P, P, P, 8, None) == EWS
    # score: dtype / impl combinations, token count
    assert lib.sixdgs_score_pass1(P, 99, 1000, P, 256, P, P, 1, P, 1 << 22, None) == EINVAL and "k_dtype" in err()
    assert lib.sixdgs_score_pass1(P, F32, 1000, P, 256, P, P, 1, P, 1 << 22, None) == EINVAL     # tensor cores need bf16 / f16x2 / f16f8
    assert lib.sixdgs_score_pass1(P, F16X2, 1000, P, 256, P, P, 0, P, 1 << 22, None) == EINVAL   # the fp16-pair formats need impl 1
    assert lib.sixdgs_score_pass1(P, F16F8, 1000, P, 256, P, P, 0, P, 1 << 22, None) == EINVAL
    assert lib.sixdgs_score_pass1(P, BF16, 1000, P, 300, P, P, 1, P, 1 << 22, None) == EINVAL and "n_img" in err()
    assert lib.sixdgs_score_pass1(P, BF16, 0, P, 256, P, P, 1, P, 1 << 22, None) == EINVAL
    assert lib.sixdgs_score_pass2(P, BF16, 1000, P, 256, P, P, P, P, 1, P, 1 << 22, None) == EINVAL and "attention map" in err()
    # batched score: at most 8 queries per launch, workspace, score stride
    ws8 = lib.sixdgs_score_batch_workspace(8)
    assert lib.sixdgs_score_batch_max() == 8 and lib.sixdgs_score_batch_parts() == 74 == lib.sixdgs_score_parts(1)
    assert ws8 == 8 * 256 * 768 * 2 + 1024 and lib.sixdgs_score_workspace(1) == 256 * 768 * 2 + 1024 and lib.sixdgs_score_workspace(0) == 0
    assert lib.sixdgs_score_pass1_batch(P, F16X2, 1000, P, 9, 256, P, P, P, ws8, None) == EINVAL and "n_queries" in err()
    assert lib.sixdgs_score_pass1_batch(P, F16X2, 1000, P, 8, 256, P, P, None, 0, None) == EWS
    assert lib.sixdgs_score_pass1_batch(P, F32, 1000, P, 8, 256, P, P, P, ws8, None) == EINVAL
    assert lib.sixdgs_score_pass2_batch(P, F16F8, 1000, P, 2, 256, P, P, P, 999, P, ws8, None) == EINVAL and "score_stride" in err()
    assert lib.sixdgs_score_pass2_batch_ls(P, F16X2, 1000, P, 2, 256, P, P, P, 1000, None, P, P, P, P, ws8, None) == EINVAL
    assert lib.sixdgs_ls_partial_rows() == 148 and lib.sixdgs_score_backward_parts() == 296
    # least-squares solve, pose tail
    assert lib.sixdgs_ls_solve(P, 0, 1.0, None, P, None, None, None, None, None) == EINVAL
    assert lib.sixdgs_ls_solve(P, 2, 1.0, None, None, None, P, None, None, None) == EINVAL and "camera-up" in err()
    assert lib.sixdgs_pose_tail(P, P, 3, P, P, 2000, P, P, None, None) == EINVAL
    assert lib.sixdgs_pose_tail(P, P, 2, P, P, 100, P, P, None, None) == EINVAL and "ray_stride" in err()
    # ray generation / features
    assert lib.sixdgs_raygen_fill(P, P, P, P, 4, 16, P, 10, P, 50, 1000, 0, P, P, P, P, None, None, None) == EINVAL and "sh_degree" in err()
    assert lib.sixdgs_raygen_fill(P, P, P, P, 3, 9, P, 10, P, 50, 1000, 0, P, P, P, P, None, None, None) == EINVAL and "coefficients" in err()
    assert lib.sixdgs_raygen_fill(P, P, P, P, 3, 16, P, 10, P, 50, 5000, 0, P, P, P, P, None, None, None) == EINVAL and "resolution" in err()
    assert lib.sixdgs_raygen_cells(None, None, 10, 50, P, None) == EINVAL
    assert lib.sixdgs_raygen_compact(P, P, P, 10, P, None, None, None, P, P, None, None, None) == EINVAL  # dir without its source
    x2ws = lib.sixdgs_ray_features_x2_workspace(1 << 20)
    assert x2ws == (1 << 17) * 2 * (704 + 512 + 384) * 2 + 1024
    assert lib.sixdgs_ray_features_x2(P, P, P, 1000, *([P] * 10), P, None, P, 16, None) == EWS
    assert lib.sixdgs_ray_features_x2(P, P, P, 1000, *([P] * 9), None, P, None, P, x2ws, None) == EINVAL
    assert lib.sixdgs_split_keys(None, 10, P, None, None) == EINVAL and lib.sixdgs_split_keys(P, 0, P, None, None) == 0
    assert lib.sixdgs_keys_f16x2_to_f16f8(None, 10, None) == EINVAL and lib.sixdgs_keys_f16x2_to_f16f8(P, 0, None) == 0
    assert lib.sixdgs_score_backward_dlogits(P, 1000, P, 256, P, P, P, P, P, P, 10, None) == EINVAL and "ldt" in err()
    # empty inputs are not errors
    assert lib.sixdgs_degrade_mask(P, 0, 50, P, None, None) == 0 and lib.sixdgs_raygen_cells(P, None, 0, 50, P, None) == 0
    assert lib.sixdgs_version() >= 200


def test_cfg_args_parser_and_experiment_discovery(tmp_path):
    """our replacement of the reference's ANTLR cfg_args grammar (cfg_grammar/Namespace.g4: INT, FLOAT, BOOL, STRING) on
    the reference's own example (parse_config.py:46) and on what argparse really writes; directory-of-experiments
    discovery with the reference's naming rules (file_utils.py:19-77)"""
    import importlib
    drv = importlib.import_module("6dgs_b200.eval_driver")
    cfg = drv.parse_config("Namespace(sh_degree=3, source_path='/home/mbortolon/data/datasets/360_v2/bicycle', "
                           "model_path='./output/ec0d365d-5', images='images', resolution=-1, white_background=False, "
                           "data_device='cuda', eval=True)")
    assert cfg == {"sh_degree": 3, "source_path": "/home/mbortolon/data/datasets/360_v2/bicycle",
                   "model_path": "./output/ec0d365d-5", "images": "images", "resolution": -1, "white_background": False,
                   "data_device": "cuda", "eval": True}
    assert drv.parse_config('Namespace(a=1.5, b=None, c=[1, 2], d=true, e="x y", f=+3, g=-0.25)') == \
        {"a": 1.5, "b": None, "c": [1, 2], "d": True, "e": "x y", "f": 3, "g": -0.25}
    assert drv.parse_config(" Namespace()\n") == {}
    for bad in ("Foo(a=1)", "Namespace(a=b)", "Namespace(1)", "Namespace(a=1", "Namespace(a=__import__('os'))", "Namespace(**x)"):
        with pytest.raises(ValueError):
            drv.parse_config(bad)
    root = tmp_path / "exps"
    for name, its in (("synthetic_lego_7", (7000, 30000)), ("synthetic_hot_dog_12", (30000,)), ("synthetic_empty_3", ()),
                      ("mip_360_bicycle_1", (100,)), ("synthetic_lego_9", ("junk", 500))):
        for it in its:
            d = root / name / "point_cloud" / f"iteration_{it}"
            d.mkdir(parents=True)
            (d / "point_cloud.ply").write_bytes(b"ply")
        (root / name).mkdir(parents=True, exist_ok=True)
    (root / "synthetic_lego_7" / "point_cloud" / "iteration_40000").mkdir()  # a checkpoint directory without its PLY
    (root / "synthetic_file_1").write_text("not a directory")
    (root / "synthetic_lego_7" / "cfg_args").write_text("Namespace(sh_degree=2, fps_sampling=None, source_path='/d/lego')")
    objs = drv.parse_exp_dir(str(root), "synthetic_")
    assert list(objs) == ["3"] * 0 + ["12", "7", "9"]  # sorted directory order; the object without a checkpoint is skipped
    assert objs["12"]["category_name"] == "synthetic_hot_dog" and objs["7"]["category_name"] == "synthetic_lego"
    assert objs["7"]["checkpoint_filepath"].endswith("synthetic_lego_7/point_cloud/iteration_30000/point_cloud.ply")
    assert objs["9"]["checkpoint_filepath"].endswith("iteration_500/point_cloud.ply")
    assert list(drv.parse_exp_dir(str(root), "mip_360_")) == ["1"] and len(drv.parse_exp_dir(str(root), "")) == 4
    assert drv.get_highest_valid_checkpoint(str(root / "synthetic_empty_3")) == ""
    a = drv.get_checkpoint_arguments(str(root / "synthetic_lego_7"))
    assert a.sh_degree == 2 and a.fps_sampling is None and a.source_path == "/d/lego" and a.not_there is None


class _StandInPackage:
    """what eval_driver.main needs from the package, without a GPU: records the calls, returns canned results"""

    def __init__(self, synthetic, stored_degree=3, fail_on=()):
        import types
        self.calls, self.synthetic, self.fail_on = [], synthetic, fail_on
        outer = self

        class Scene:
            max_sh_degree = stored_degree

            @classmethod
            def load_ply(cls, path, device="cuda"):
                outer.calls.append(("load_ply", path, str(device)))
                if any(f in path for f in outer.fail_on):
                    raise RuntimeError("corrupt checkpoint")
                return cls()

        class Idm(torch.nn.Module):
            def __init__(self, backbone_type, score_impl=None, backbone=None):
                super().__init__()
                outer.calls.append(("idm", backbone_type, score_impl, type(backbone).__name__))

        self.GaussianScene, self.IdentificationModule = Scene, Idm
        self.DistanceBasedScoreLoss = lambda: "loss_fn"
        self.generate_all_possible_rays = lambda scene, sample_quadricell_targets=50, max_ellipsoids=1000: (
            outer.calls.append(("rays", sample_quadricell_targets, max_ellipsoids)) or tuple(torch.zeros(17, 3) for _ in range(3)))

        def tpe(cams, idm, ori, dirs, rgb, model_up, sequence_id="", category_id="", loss_fn=None):
            outer.calls.append(("tpe", len(cams), sequence_id, category_id, loss_fn, tuple(round(float(x), 4) for x in model_up)))
            res = [{"frame_id": i, "sequence_id": sequence_id, "pred_c2w": torch.eye(4).tolist(), "gt_c2w": torch.eye(4).tolist()}
                   for i in range(len(cams))]
            return res, 0.25, 3.0, (0.5 if loss_fn else -1.0), (0.75 if loss_fn else -1.0)

        self.test_pose_estimation = tpe
        del types


def test_eval_driver_control_flow_single_and_directory_modes(synthetic, tmp_path):
    """eval_driver.main with a stand-in for the package (the real one needs a GPU; tests/test_gpu_pipeline.py runs it):
    single experiment with --oracle_rays, cfg_args / PLY degree mismatch, and the reference driver's
    directory-of-experiments mode with a failing object"""
    import importlib
    import json
    import shutil
    drv = importlib.import_module("6dgs_b200.eval_driver")
    exp, img_dir, p = _write_experiment(tmp_path, synthetic)
    (exp / "cfg_args").write_text("Namespace(sh_degree=3, source_path='/data/x', white_background=False)")
    pk = _StandInPackage(synthetic)
    out = tmp_path / "r.json"
    res = drv.main(["--exp_path", str(exp), "--images", str(img_dir), "--out", str(out), "--every", "2", "--oracle_rays",
                    "--device", "cpu", "--max_ellipsoids", "0", "--backbone", "synthetic"], sx=pk)
    assert json.load(open(out)) == json.loads(json.dumps(res))
    assert res["n_rays"] == 17 and res["trained_weights"] is True and res["source_path"] == "/data/x"
    assert res["oracle_rays"] == {"avg_translation_error": 0.25, "avg_angular_error": 3.0, "avg_score_loss": 0.5, "recall": 0.75}
    kinds = [c[0] for c in pk.calls]
    assert kinds == ["load_ply", "idm", "rays", "tpe", "tpe"] and pk.calls[2] == ("rays", 50, None)
    assert pk.calls[1] == ("idm", "dino", "tc_f16x2", "SyntheticBackbone") and pk.calls[0][2] == "cpu"
    up = np.mean([p["R"][i].numpy()[:3, 1] for i in range(3)], axis=0)
    assert pk.calls[3][1:5] == (2, "exp", "", "loss_fn") and pk.calls[4][4] is None  # every 2nd of 3 cameras; oracle pass first
    np.testing.assert_allclose(pk.calls[4][5], up, atol=1e-4)  # model up from ALL cameras
    with pytest.raises(ValueError, match="sh_degree"):
        drv.main(["--exp_path", str(exp), "--images", str(img_dir), "--out", str(out), "--device", "cpu"],
                 sx=_StandInPackage(synthetic, stored_degree=2))
    # directory of experiments: two blender objects (one of them fails to load), one of another dataset
    root = tmp_path / "all"
    for name in ("synthetic_lego_7", "synthetic_chair_8", "tt_truck_1"):
        shutil.copytree(exp, root / name)
    shutil.copytree(img_dir, tmp_path / "imgs" / "synthetic_lego_7")
    shutil.copytree(img_dir, tmp_path / "imgs" / "synthetic_chair_8")
    pk = _StandInPackage(synthetic, fail_on=("synthetic_chair_8",))
    res = drv.main(["--exp_path", str(root), "--data_type", "blender", "--images", str(tmp_path / "imgs"), "--out", str(out),
                    "--every", "1", "--device", "cpu"], sx=pk)
    assert list(res["objects"]) == ["7"] and len(res["results"]) == 3 and res["results"][0]["sequence_id"] == "7"
    assert [c for c in pk.calls if c[0] == "tpe"][0][2:4] == ("7", "synthetic_lego")
    assert sum(c[0] == "load_ply" for c in pk.calls) == 2 and pk.calls[0][1].endswith("synthetic_chair_8/point_cloud/iteration_30000/point_cloud.ply")
    assert "results" not in res["objects"]["7"] and res["objects"]["7"]["n_rays"] == 17
    with pytest.raises(FileNotFoundError):
        drv.main(["--exp_path", str(root), "--data_type", "mip360", "--out", str(out), "--device", "cpu"], sx=pk)


def test_ctypes_signatures_match_the_header_prototypes():
    """every prototype of include/sixdgs.h against the argument / return types the ctypes binding (6dgs_b200/_lib.py)
    attaches: pointer -> c_void_p, int64_t -> c_int64, size_t -> c_size_t, float / double / int by value.  A drift here
    would pass garbage on the stack without any compiler noticing."""
    import ctypes
    import importlib
    import re
    lib = importlib.import_module("6dgs_b200._lib")
    src = open(os.path.join(ROOT, "include", "sixdgs.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = re.findall(r"\b(int|size_t|const char\s*\*)\s+(sixdgs_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert {p[1] for p in protos} == set(lib._SIGNATURES) and len(protos) >= 40

    def kind(arg):
        a = " ".join(arg.split())
        if a in ("void", ""):
            return None
        if "*" in a:
            return ctypes.c_void_p
        for pat, t in ((r"\bint64_t\b", ctypes.c_int64), (r"\bsize_t\b", ctypes.c_size_t), (r"\bdouble\b", ctypes.c_double),
                       (r"\bfloat\b", ctypes.c_float), (r"\bint\b", ctypes.c_int)):
            if re.search(pat, a):
                return t
        raise AssertionError(f"unrecognised parameter {a!r}")

    for ret, name, args in protos:
        want = [k for k in (kind(a) for a in args.split(",")) if k is not None]
        got, res = lib._SIGNATURES[name]
        assert want == got, (name, [k.__name__ for k in want], [k.__name__ for k in got])
        assert res == {"int": ctypes.c_int, "size_t": ctypes.c_size_t}.get(ret, ctypes.c_char_p), name
