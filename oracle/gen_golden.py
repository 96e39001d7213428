"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (``python oracle/gen_golden.py``); the GPU
box never has /root/reference -- it consumes the committed fixtures.  Inputs come from
``6dgs_b200/synthetic.py`` (seeded) and are stored in the fixtures next to the reference outputs,
so tests never depend on RNG reproducibility across machines.

Stage boundaries follow SURVEY.md §8c:
  scales -> valid | (a,b,c) -> (points, ellipsoid_id) | pts -> normals | cov -> (vals, vecs) |
  scene -> (ori, dir, rgb) | (ori,dir,rgb,W) -> ray_fea | (img_fea, ray_fea, W) -> (A, score) |
  score -> topk | (o,d[,w]) -> centre | test_pose_estimation -> pred_c2w
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shims  # noqa: E402

ref_shims.install()
synthetic = importlib.import_module("6dgs_b200.synthetic")

from pose_estimation import quadricell as ref_q  # noqa: E402
from pose_estimation import sampling as ref_s  # noqa: E402
from pose_estimation.identification_module import IdentificationModule  # noqa: E402
from pose_estimation.line_intersection import (compute_line_intersection_impl2, exclude_negatives,  # noqa: E402
                                               make_rotation_mat)
from pose_estimation.sym_eig_3x3 import sym_eig_3x3  # noqa: E402
from pose_estimation.test import test_pose_estimation  # noqa: E402
from scene.scene_structure import CameraInfo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(max(1, os.cpu_count() or 1))


def npz(name, **kw):
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()}
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print(f"{name}: " + ", ".join(f"{k}{tuple(a.shape)}" for k, a in arrs.items()))


def ref_model(sc):
    return ref_shims.make_gaussian_model(sc["xyz"], sc["scaling"], sc["rotation"], sc["features_dc"],
                                         sc["features_rest"], sc["sh_degree"])


def gen_quadricell():
    # a2: degrade mask on a set that contains both valid and degraded (needle-like) ellipsoids
    g = torch.Generator().manual_seed(11)
    scales = torch.exp(-4.0 + 1.5 * torch.randn(512, 3, generator=g))
    valid = ref_q.mask_degraded_ellipsoids(scales[:, 0], scales[:, 1], scales[:, 2])
    # a6: cell centres for uniform and heavy-tailed valid ellipsoids
    uni = torch.rand(40, 3, generator=g) * 0.05 + 0.005
    hv = scales[valid][:24]
    abc = torch.cat((uni, hv), 0)
    pts, eid = ref_q.compute_quadricell_centers(abc[:, 0], abc[:, 1], abc[:, 2], target_points=50)
    npz("quadricell.npz", mask_scales=scales, mask_valid=valid, abc=abc, points=pts, ellipsoid_id=eid)


def gen_sym_eig():
    g = torch.Generator().manual_seed(5)
    X = torch.randn(192, 20, 3, generator=g) * torch.rand(192, 1, 3, generator=g)
    X = X - X.mean(1, keepdim=True)
    A = X.mT @ X
    diag = torch.diag_embed(torch.rand(8, 3, generator=g))
    near = A[:8] * 1e-3 + torch.eye(3) * 0.5
    flat = (X[:16] * torch.tensor([1.0, 1.0, 1e-3])).mT @ (X[:16] * torch.tensor([1.0, 1.0, 1e-3]))
    A = torch.cat((A, diag, near, flat), 0)
    vals, vecs = sym_eig_3x3(A, eigenvectors=True)
    npz("sym_eig.npz", A=A, vals=vals, vecs=vecs)


def gen_normals():
    g = torch.Generator().manual_seed(9)
    pts = torch.randn(600, 3, generator=g)
    n = ref_s.compute_normals(pts[:300], pts, k_neighbors=20)
    npz("normals.npz", cloud=pts, normals=n)


def gen_rays():
    sc = synthetic.synth_scene(192, seed=1)
    gm = ref_model(sc)
    torch.manual_seed(77)
    nvalid = int(ref_q.mask_degraded_ellipsoids(*torch.exp(sc["scaling"]).unbind(-1)).sum())
    perm = torch.randperm(nvalid, dtype=torch.long)[: min(1000, nvalid)]
    torch.manual_seed(77)
    ori, dirs, rgb = ref_s.generate_all_possible_rays(gm)
    npz("rays_small.npz", xyz=sc["xyz"], scaling=sc["scaling"], rotation=sc["rotation"],
        features_dc=sc["features_dc"], features_rest=sc["features_rest"], perm=perm,
        ori=ori, dirs=dirs, rgb=rgb)
    # capped case (N > 1000 valid): keep every 16th ray + the count
    sc2 = synthetic.synth_scene(1500, seed=2)
    gm2 = ref_model(sc2)
    torch.manual_seed(78)
    nvalid2 = int(ref_q.mask_degraded_ellipsoids(*torch.exp(sc2["scaling"]).unbind(-1)).sum())
    perm2 = torch.randperm(nvalid2, dtype=torch.long)[:1000]
    torch.manual_seed(78)
    o2, d2, c2 = ref_s.generate_all_possible_rays(gm2)
    npz("rays_capped.npz", scene_n=1500, scene_seed=2, perm=perm2, n_rays=o2.shape[0],
        ori_s=o2[::16], dirs_s=d2[::16], rgb_s=c2[::16],
        sums=torch.stack((o2.double().sum(0), d2.double().sum(0), c2.double().sum(0))))
    # heavy-tailed scales (log-normal, sigma 1.5): wide spread in cells per ellipsoid, some ellipsoids degraded
    sc3 = synthetic.synth_scene(160, seed=21, heavy_tail=True)
    sc3["scaling"] = -4.0 + 1.5 * torch.randn(160, 3, generator=torch.Generator().manual_seed(22))
    gm3 = ref_model(sc3)
    torch.manual_seed(79)
    nvalid3 = int(ref_q.mask_degraded_ellipsoids(*torch.exp(sc3["scaling"]).unbind(-1)).sum())
    perm3 = torch.randperm(nvalid3, dtype=torch.long)[: min(1000, nvalid3)]
    torch.manual_seed(79)
    o3, d3, c3 = ref_s.generate_all_possible_rays(gm3)
    npz("rays_heavy.npz", scene_n=160, scene_seed=21, scaling=sc3["scaling"], n_valid=nvalid3, perm=perm3, n_rays=o3.shape[0],
        ori_s=o3[::8], dirs_s=d3[::8], rgb_s=c3[::8],
        sums=torch.stack((o3.double().sum(0), d3.double().sum(0), c3.double().sum(0))))
    return sc, ori, dirs, rgb


def gen_id_module(ori, dirs, rgb):
    w = synthetic.synth_id_weights(seed=3)
    idm = IdentificationModule("dino").eval()
    missing = idm.load_state_dict(w, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    img = synthetic.synth_image(64, 64, seed=4)
    mask = torch.ones(64, 64, dtype=torch.bool)
    with torch.no_grad():
        tok_pe, tok, grid = idm.backbone_wrapper(img, mask)
        fea = idm.ray_preprocessor(ori, dirs, rgb)
        A = idm.attention(tok_pe, fea)
        idx, vals, scores, up, _ = idm.test_image(img, mask, ori, dirs, rgb, rays_to_output=100)
        # masked query: a disc-shaped mask removes some tokens (n_img < 256)
        yy, xx = torch.meshgrid(torch.arange(64), torch.arange(64), indexing="ij")
        mask2 = ((yy - 30) ** 2 + (xx - 34) ** 2) < 24 ** 2
        tok_pe2, _, _ = idm.backbone_wrapper(img, mask2)
        idx2, vals2, scores2, up2, _ = idm.test_image(img, mask2, ori, dirs, rgb, rays_to_output=100)
    sel = torch.arange(0, ori.shape[0], 23)
    q = torch.nn.functional.linear(tok_pe, w["attention.q_proj.weight"], w["attention.q_proj.bias"])
    k = torch.nn.functional.linear(fea, w["attention.k_proj.weight"], w["attention.k_proj.bias"])
    L = (q @ k.t()) / (384 ** 0.5)
    npz("id_module.npz", weight_seed=3, img=img, mask2=mask2, tok_pe=tok_pe, tok_pe2_n=tok_pe2.shape[0],
        tok_pe2_head=tok_pe2[:4], fea_sel=sel, fea=fea[sel], k_sel=k[sel], A_rows=A[[0, 100, 255]],
        row_max=L.max(-1).values, row_lse=torch.logsumexp(L, -1), scores=scores, topk_idx=idx,
        topk_vals=vals, up=up, scores2=scores2, topk_idx2=idx2, topk_vals2=vals2, up2=up2,
        weight_checksum=torch.stack([v.double().abs().sum() for _, v in sorted(w.items())]))
    return idm


def gen_id_module_peaked(ori, dirs, rgb, q_gain=20.0):
    """Same rays / image / weights as gen_id_module but with attention.q_proj scaled by q_gain: the random-init logits
    are nearly flat (std 0.33), which hides low-precision logit error; x20 gives std ~ 6.5 and per-token effective
    supports down to ~1 ray, like a trained softmax.  Scores, top-100 and the full pose from the UNMODIFIED reference."""
    w = synthetic.synth_id_weights(seed=3, q_gain=q_gain)
    idm = IdentificationModule("dino").eval()
    missing = idm.load_state_dict(w, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    img = synthetic.synth_image(64, 64, seed=4)
    mask = torch.ones(64, 64, dtype=torch.bool)
    with torch.no_grad():
        tok_pe, _, _ = idm.backbone_wrapper(img, mask)
        fea = idm.ray_preprocessor(ori, dirs, rgb)
        idx, vals, scores, up, _ = idm.test_image(img, mask, ori, dirs, rgb, rays_to_output=100)
        q = torch.nn.functional.linear(tok_pe, w["attention.q_proj.weight"], w["attention.q_proj.bias"])
        k = torch.nn.functional.linear(fea, w["attention.k_proj.weight"], w["attention.k_proj.bias"])
        L = (q @ k.t()) / (384 ** 0.5)
    R = np.eye(3, dtype=np.float32)
    T = np.array([0.1, -0.2, 2.5], dtype=np.float32)
    img_u8 = (img * 255).to(torch.uint8).numpy()
    cam = CameraInfo(uid=0, R=R, T=T, FovY=np.float32(0.9), FovX=np.float32(0.9), image=img_u8, image_path="",
                     image_name="0", width=64, height=64)
    results, _, _, _, _ = test_pose_estimation([cam], idm, ori, dirs, rgb, torch.tensor([0.0, 0.0, 1.0]))
    print(f"peaked: logit std {L.std():.2f}, range {L.max() - L.min():.1f}, score max/min {scores.max():.3e}/{scores.min():.3e}")
    npz("id_module_peaked.npz", weight_seed=3, q_gain=q_gain, img=img, img_u8=img_u8, scores=scores, topk_idx=idx,
        topk_vals=vals, up=up, row_max=L.max(-1).values, row_lse=torch.logsumexp(L, -1), logit_std=L.std(),
        R=R, T=T, pred_c2w=np.array(results[0]["pred_c2w"], dtype=np.float32))


def gen_ply():
    """point_cloud.ply written by the reference's OWN save_ply (scene/gaussian_model.py:298-332: attribute order from
    construct_list_of_attributes :284-296, SH coefficients transposed to channel-major) through the plyfile-format shim
    of oracle/ref_shims.py; the tensors it was written from are stored next to it."""
    sc = synthetic.synth_scene(96, seed=41)
    gm = ref_model(sc)
    gm._opacity = torch.randn(96, 1, generator=torch.Generator().manual_seed(42))
    path = os.path.join(OUT, "point_cloud_ref.ply")
    gm.save_ply(path)
    npz("ply_ref.npz", xyz=sc["xyz"], scaling=sc["scaling"], rotation=sc["rotation"], features_dc=sc["features_dc"],
        features_rest=sc["features_rest"], opacity=gm._opacity, attributes=np.array(gm.construct_list_of_attributes()))
    print(f"point_cloud_ref.ply: {os.path.getsize(path)} bytes")


def gen_line_intersection():
    g = torch.Generator().manual_seed(21)
    cases = {}
    centre = torch.tensor([0.3, -1.2, 2.0])
    o = torch.randn(100, 3, generator=g)
    d = torch.nn.functional.normalize(centre[None] - o + 0.02 * torch.randn(100, 3, generator=g), dim=-1)
    d[::7] = -d[::7]
    w = torch.rand(100, generator=g)
    cases["o"], cases["d"], cases["w"] = o, d, w
    cases["c_unweighted"] = compute_line_intersection_impl2(o, d)
    cases["c_weighted"] = compute_line_intersection_impl2(o, d, weights=w)
    cases["neg_mask"] = exclude_negatives(cases["c_unweighted"], o, d)
    par_d = torch.tensor([[0.0, 0.0, 1.0]]).repeat(5, 1)
    cases["o_par"], cases["d_par"] = o[:5], par_d
    cases["c_parallel"] = compute_line_intersection_impl2(o[:5], par_d)
    dirv = torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
    up = torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
    cases["rot_dir"], cases["rot_up"] = dirv, up
    cases["rot"] = make_rotation_mat(dirv, up)
    npz("line_intersection.npz", **cases)


def pose_cameras():
    """the three fixture cameras (random rigid poses, one RGBA image) of pose.npz / loss.npz"""
    g = torch.Generator().manual_seed(31)
    cams, imgs = [], []
    for i in range(3):
        qv = torch.randn(4, generator=g)
        qv = qv / qv.norm()
        w_, x, y, z = qv.tolist()
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w_ * z), 2 * (x * z + w_ * y)],
                      [2 * (x * y + w_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w_ * x)],
                      [2 * (x * z - w_ * y), 2 * (y * z + w_ * x), 1 - 2 * (x * x + y * y)]], dtype=np.float32)
        T = (torch.randn(3, generator=g) * 2).numpy().astype(np.float32)
        chans = 4 if i == 1 else 3
        img = (torch.rand(64, 64, chans, generator=g) * 255).to(torch.uint8).numpy()
        if chans == 4:
            img[..., 3] = np.where(np.add.outer(np.arange(64), np.arange(64)) > 30, 255, 40).astype(np.uint8)
        imgs.append(img)
        cams.append(CameraInfo(uid=i, R=R, T=T, FovY=np.float32(0.9), FovX=np.float32(0.9), image=img,
                               image_path="", image_name=str(i), width=64, height=64))
    return cams, imgs


def gen_pose(idm, ori, dirs, rgb):
    cams, imgs = pose_cameras()
    results, t_err, a_err, _, _ = test_pose_estimation(cams, idm, ori, dirs, rgb,
                                                       torch.tensor([0.0, 0.0, 1.0]))
    npz("pose.npz", R=np.stack([c.R for c in cams]), T=np.stack([c.T for c in cams]),
        img0=imgs[0], img1=imgs[1], img2=imgs[2],
        pred_c2w=np.array([r["pred_c2w"] for r in results], dtype=np.float32),
        gt_c2w=np.array([r["gt_c2w"] for r in results], dtype=np.float32),
        loss=np.array([r["loss"] for r in results], dtype=np.float32),
        avg_t_err=t_err, avg_ang_err=a_err)


def fixture_id_module():
    """the identification module of id_module.npz (weights seed 3), without rewriting that fixture"""
    idm = IdentificationModule("dino").eval()
    missing = idm.load_state_dict(synthetic.synth_id_weights(seed=3), strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    return idm


def gen_loss(idm, ori, dirs, rgb):
    """DistanceBasedScoreLoss / best_one_to_one_rays_selector of the UNMODIFIED reference on the fixture rays and the
    pose.npz cameras, and test_pose_estimation in its "oracle rays" mode (loss_fn given: the pose is solved from the
    top-100 TARGET scores, test.py:110-142)."""
    from pose_estimation.distance_based_loss import DistanceBasedScoreLoss, best_one_to_one_rays_selector
    from utils.graphics_utils import fov2focal
    cams, _ = pose_cameras()
    pred = torch.from_numpy(np.load(os.path.join(OUT, "id_module.npz"))["scores"])
    out = {}
    for i, cam in enumerate(cams):
        w2c = torch.eye(4)
        w2c[:3, :3] = torch.from_numpy(cam.R).T
        w2c[:3, 3] = torch.from_numpy(cam.T)
        pose = torch.inverse(w2c)
        f = fov2focal(cam.FovX, cam.width)
        K = torch.tensor([[f, 0.0, cam.width / 2], [0.0, fov2focal(cam.FovY, cam.height), cam.height / 2], [0.0, 0.0, 1.0]])
        for shape in ((800, 800), (480, 640)):
            _, inside, tgt, tgt_d = best_one_to_one_rays_selector(K, pose, shape, dirs, ori, backbone_wh=(16, 16))
            out[f"inside{i}_{shape[0]}"] = inside
        loss, target = DistanceBasedScoreLoss()(pred, pose, K, ori, dirs, 256, (16, 16), model_up=None)
        out.update({f"pose{i}": pose, f"K{i}": K, f"target_raw{i}": tgt, f"target_dist{i}": tgt_d, f"target{i}": target,
                    f"loss{i}": loss})
    results, t_err, a_err, avg_loss, avg_recall = test_pose_estimation(cams, idm, ori, dirs, rgb, torch.tensor([0.0, 0.0, 1.0]),
                                                                       loss_fn=DistanceBasedScoreLoss())
    npz("loss.npz", pred=pred, pred_c2w=np.array([r["pred_c2w"] for r in results], dtype=np.float32),
        scores_loss=np.array([r["scores_loss"] for r in results], dtype=np.float32),
        recall=np.array([r["recall"] for r in results], dtype=np.float32), avg_t_err=t_err, avg_ang_err=a_err,
        avg_loss=avg_loss, avg_recall=avg_recall, **out)


if __name__ == "__main__":
    if "--only-ply" in sys.argv:
        gen_ply()
        sys.exit(0)
    if "--only-peaked" in sys.argv:  # add the peaked-softmax fixture without touching the others (rays from the fixture)
        g = np.load(os.path.join(OUT, "rays_small.npz"))
        gen_id_module_peaked(*(torch.from_numpy(g[k]) for k in ("ori", "dirs", "rgb")))
        sys.exit(0)
    if "--only-loss" in sys.argv:  # add the loss / oracle-rays fixture without touching the others
        g = np.load(os.path.join(OUT, "rays_small.npz"))
        gen_loss(fixture_id_module(), *(torch.from_numpy(g[k]) for k in ("ori", "dirs", "rgb")))
        sys.exit(0)
    if "--only-heavy" in sys.argv:  # add the heavy-tail fixture without touching the others
        gen_rays()
        sys.exit(0)
    gen_quadricell()
    gen_sym_eig()
    gen_normals()
    sc, ori, dirs, rgb = gen_rays()
    idm = gen_id_module(ori, dirs, rgb)
    gen_id_module_peaked(ori, dirs, rgb)
    gen_line_intersection()
    gen_pose(idm, ori, dirs, rgb)
    gen_loss(idm, ori, dirs, rgb)
    gen_ply()
    os.system(f"du -sh {OUT}")
