"""Offline evaluation driver: the evaluation half of the reference's ``pretrain_eval_attention.py`` (:31-154,200-248)
without its network / ANTLR dependencies.

The reference driver discovers ``point_cloud/iteration_*/point_cloud.ply`` under an experiment directory
(``pose_estimation/file_utils.py:19-72``), re-parses ``cfg_args`` with an ANTLR-3 grammar that no longer loads,
re-reads the dataset, downloads DINOv2 through torch.hub, fits the identification module, explores the model and runs
``test_pose_estimation`` twice.  Here everything a 3DGS experiment directory already contains is enough:

    <exp>/point_cloud/iteration_<N>/point_cloud.ply     the scene (highest iteration wins, like the reference)
    <exp>/cameras.json                                  3DGS camera dump (id, img_name, width, height, position,
                                                        rotation, fx, fy -- scene/__init__.py / utils/camera_utils.py)
    <exp>/id_module.th                                  optional: {"model_state_dict": ...} (pose_estimation/train.py:309-317)
    <images>/<img_name>.{png,jpg,jpeg}                  the query images (RGBA alpha is composited like test.py:69-83)

    python tools/eval_pose.py --exp_path <exp> --images <dir> --out results.json [--dinov2 vits14.pth]

``--exp_path`` may also be a directory OF experiment directories, as the reference driver takes it
(``pretrain_eval_attention.py:200-248``): sub-directories ``<prefix><category>_<sequence>`` (``--data_type`` picks the
prefix) are evaluated one after the other (``parse_exp_dir``).  An experiment's ``cfg_args`` (the ``Namespace(...)`` repr
3DGS training writes, train.py:207-208) is read with a small parser of our own instead of the reference's ANTLR grammar
(``cfg_grammar/``, serialised for an ANTLR runtime that no longer loads): ``sh_degree`` is checked against the PLY.

Training is out of scope: without ``id_module.th`` the module is randomly initialised and the poses are
meaningless -- the driver says so and still runs (smoke / timing use).  Output: the reference's JSON schema
(list of per-frame dicts with pred_c2w / gt_c2w, test.py:290-302) plus the two averages.
"""
from __future__ import annotations

import argparse
import glob
import json
import math
import os
from collections import namedtuple
from typing import List, Optional

import numpy as np
import torch

CameraInfo = namedtuple("CameraInfo", "uid R T FovY FovX image image_path image_name width height")


def find_point_cloud(exp_path: str) -> str:
    """the ``point_cloud/iteration_<N>/point_cloud.ply`` with the largest N; directory names whose second ``_`` component
    is not an integer, and checkpoint directories without their PLY, are ignored (reference file_utils.py:19-43)"""
    best, best_path = None, ""
    # upstream walks the directory names in reverse lexicographic order and lets a later equal iteration number win
    # ("iteration_7000" before "iteration_07000"): same walk here, so ties resolve identically
    for path in sorted(glob.glob(os.path.join(exp_path, "point_cloud", "iteration_*", "point_cloud.ply")),
                       key=lambda p: os.path.basename(os.path.dirname(p)), reverse=True):
        parts = os.path.basename(os.path.dirname(path)).split("_")
        try:
            it = int(parts[1])
        except ValueError:
            continue
        if best is None or it >= best:
            best, best_path = it, path
    if not best_path:
        raise FileNotFoundError(f"no point_cloud/iteration_*/point_cloud.ply under {exp_path}")
    return best_path


def get_highest_valid_checkpoint(root_dir: str) -> str:
    """reference name and convention (file_utils.py:19-43): "" when the experiment holds no checkpoint"""
    try:
        return find_point_cloud(root_dir)
    except FileNotFoundError:
        return ""


class dotdict(dict):
    """attribute access, a missing key reads as None (file_utils.py:5-10)"""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


def parse_config(text: str) -> dict:
    """``Namespace(key=value, ...)`` -> dict.  Replaces cfg_grammar/parse_config.py (grammar Namespace.g4: INT, FLOAT, BOOL,
    STRING values): the text is a Python call expression, so Python's own parser reads it.  Beyond the grammar, ``None``,
    negative numbers and lists are accepted (argparse writes them; the reference's parser rejects the whole file), and
    ``False`` parses as False -- upstream evaluates ``bool("False")`` and gets True (parse_config.py:33-34)."""
    import ast

    try:
        call = ast.parse(text.strip(), mode="eval").body
    except SyntaxError as e:
        raise ValueError(f"cfg_args is not a Namespace(...) expression: {e}") from None
    if not (isinstance(call, ast.Call) and isinstance(call.func, ast.Name) and call.func.id == "Namespace" and not call.args):
        raise ValueError("cfg_args is not a Namespace(...) expression")
    out = {}
    for kw in call.keywords:
        if kw.arg is None:
            raise ValueError("cfg_args: **kwargs are not supported")
        if isinstance(kw.value, ast.Name) and kw.value.id in ("true", "false"):  # the grammar's lower-case booleans
            out[kw.arg] = kw.value.id == "true"
            continue
        try:
            out[kw.arg] = ast.literal_eval(kw.value)
        except ValueError:
            raise ValueError(f"cfg_args: value of {kw.arg!r} not recognised") from None  # upstream: 'type did not recognized'
    return out


def get_checkpoint_arguments(root_dir: str) -> dotdict:
    """<exp>/cfg_args -> dotdict (file_utils.py:13-16)"""
    with open(os.path.join(root_dir, "cfg_args")) as fh:
        return dotdict(parse_config(fh.read()))


DATA_TYPE_PREFIX = {"blender": "synthetic_", "mip360": "mip_360_", "tankstemple": "tt_", "cambridge_landmark": "cl_"}


def parse_exp_dir(exp_dir: str, prefix: str) -> dict:
    """sub-directories ``<prefix>...<category>_<sequence>`` with a checkpoint -> {sequence_id: {exp_dir_filepath,
    checkpoint_filepath, sequence_id, category_name}}, in sorted order; a later directory with the same sequence id
    replaces an earlier one (file_utils.py:46-77, pretrain_eval_attention.py:205-216)"""
    found = {}
    for name in sorted(os.listdir(exp_dir)):
        path = os.path.join(exp_dir, name)
        if not (os.path.isdir(path) and name.startswith(prefix)):
            continue
        parts = name.split("_")
        seq, cat = parts[-1], "_".join(parts[:-1])
        ckpt = get_highest_valid_checkpoint(path)
        if not ckpt:
            print(f"Object {seq} of category {cat} skipped because no valid checkpoint found")
            continue
        found[seq] = {"exp_dir_filepath": path, "checkpoint_filepath": ckpt, "sequence_id": seq, "category_name": cat}
    return found


def focal2fov(focal: float, pixels: float) -> float:
    return 2.0 * math.atan(pixels / (2.0 * focal))


def cameras_from_json(path: str, image_dir: Optional[str], every: int = 1, load_images: bool = True) -> List[CameraInfo]:
    """3DGS ``cameras.json`` -> CameraInfo list.  The dump stores the camera-to-world rotation and the camera position
    (utils/camera_utils.py:camera_to_JSON); CameraInfo keeps R = c2w rotation and T = w2c translation = -R^T pos
    (scene/dataset_readers.py), which is what test.py:47-67 turns back into a pose."""
    from PIL import Image

    cams = []
    for i, c in enumerate(json.load(open(path))):
        if i % every:
            continue
        R = np.asarray(c["rotation"], dtype=np.float32)
        pos = np.asarray(c["position"], dtype=np.float32)
        T = (-R.T @ pos).astype(np.float32)
        w, h = int(c["width"]), int(c["height"])
        img, img_path = None, ""
        if load_images:
            if image_dir is None:
                raise ValueError("--images is required to load the query images")
            for ext in ("", ".png", ".jpg", ".jpeg", ".JPG", ".PNG"):
                p = os.path.join(image_dir, c["img_name"] + ext)
                if os.path.isfile(p):
                    img_path = p
                    break
            if not img_path:
                raise FileNotFoundError(f"image {c['img_name']} not found in {image_dir}")
            img = np.array(Image.open(img_path))
            if img.ndim == 2:
                img = np.repeat(img[..., None], 3, axis=-1)
            h, w = img.shape[:2]
        cams.append(CameraInfo(int(c.get("id", i)), R, T, np.float32(focal2fov(c["fy"], c["height"])),
                               np.float32(focal2fov(c["fx"], c["width"])), img, img_path, c["img_name"], w, h))
    return cams


def model_up_from_cameras(cams: List[CameraInfo]) -> np.ndarray:
    """mean of the cameras' R[:3, 1] columns (pretrain_eval_attention.py:91-98)"""
    return np.mean(np.asarray([c.R[:3, 1] for c in cams], dtype=np.float32), axis=0)


def load_id_module(sx, exp_path: Optional[str], weights: Optional[str], score_impl: str, device, backbone=None):
    idm = sx.IdentificationModule("dino", score_impl=score_impl, backbone=backbone)
    path = weights or (os.path.join(exp_path, "id_module.th") if exp_path else None)
    trained = False
    if path and os.path.exists(path):
        ckpt = torch.load(path, map_location="cpu")
        sd = ckpt.get("model_state_dict", ckpt)
        missing = idm.load_state_dict(sd, strict=False)
        hot = [k for k in missing.missing_keys if k.startswith(("ray_preprocessor.", "attention.", "camera_direction"))]
        if hot:
            raise KeyError(f"{path} lacks hot-path parameters: {hot[:4]} ...")
        trained = True
    else:
        print("[eval_driver] no id_module.th: the identification module is RANDOMLY initialised -- poses are meaningless "
              "(training is out of scope here; fit it with the reference's train_id_module)")
    return idm.to(device).eval().requires_grad_(False), trained


def evaluate_experiment(sx, args, dev, exp_path: Optional[str], ply: str, image_dir: Optional[str], sequence_id: str,
                        category_id: str = "") -> dict:
    """one object: scene -> rays -> (optional oracle-rays pass) -> test_pose_estimation; the evaluation half of
    pretrain_single_object (pretrain_eval_attention.py:31-154)"""
    cams_json = args.cameras or os.path.join(exp_path, "cameras.json")
    scene = sx.GaussianScene.load_ply(ply, device=dev)
    cfg = None
    if exp_path and os.path.isfile(os.path.join(exp_path, "cfg_args")):
        cfg = get_checkpoint_arguments(exp_path)
        if cfg.sh_degree is not None and int(cfg.sh_degree) != int(scene.max_sh_degree):
            # the reference asserts the same on the f_rest_* count (gaussian_model.py:368)
            raise ValueError(f"cfg_args says sh_degree={cfg.sh_degree}, {ply} stores degree {scene.max_sh_degree}")
    all_cams = cameras_from_json(cams_json, image_dir, every=1, load_images=False)
    test_cams = cameras_from_json(cams_json, image_dir, every=max(1, args.every))
    model_up = torch.from_numpy(model_up_from_cameras(all_cams)).to(dev)
    backbone = sx.synthetic.SyntheticBackbone() if args.backbone == "synthetic" else None
    idm, trained = load_id_module(sx, exp_path, args.weights, args.score_impl, dev, backbone)
    rays = sx.generate_all_possible_rays(scene, sample_quadricell_targets=50,
                                         max_ellipsoids=None if args.max_ellipsoids == 0 else args.max_ellipsoids)
    oracle_rays = None
    if args.oracle_rays:
        _, o_t, o_a, o_loss, o_recall = sx.test_pose_estimation(test_cams, idm, *rays, model_up, sequence_id=sequence_id,
                                                                category_id=category_id, loss_fn=sx.DistanceBasedScoreLoss())
        oracle_rays = {"avg_translation_error": o_t, "avg_angular_error": o_a, "avg_score_loss": o_loss, "recall": o_recall}
    results, t_err, a_err, _, _ = sx.test_pose_estimation(test_cams, idm, *rays, model_up, sequence_id=sequence_id,
                                                          category_id=category_id)
    print(f"[eval_driver] {sequence_id or ply}: {len(results)} frames, {rays[0].shape[0]} rays: "
          f"translation {t_err:.4f}, angular {a_err:.3f} deg")
    return {"results": results, "avg_translation_error": t_err, "avg_angular_error": a_err, "n_rays": int(rays[0].shape[0]),
            "trained_weights": trained, "point_cloud": ply, "oracle_rays": oracle_rays,
            "source_path": cfg.source_path if cfg else None}


def main(argv=None, sx=None):
    import importlib
    import traceback

    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--exp_path", help="3DGS experiment directory (point_cloud/, cameras.json, optional id_module.th, cfg_args), "
                                       "or a directory of such directories (see --data_type)")
    ap.add_argument("--data_type", default="", choices=[""] + sorted(DATA_TYPE_PREFIX),
                    help="directory-of-experiments mode: only sub-directories with this dataset's prefix "
                         "(pretrain_eval_attention.py:211-220); empty = every sub-directory")
    ap.add_argument("--ply", help="explicit point_cloud.ply (overrides --exp_path discovery)")
    ap.add_argument("--cameras", help="explicit cameras.json")
    ap.add_argument("--images", help="directory with the query images named by cameras.json:img_name (directory-of-experiments "
                                     "mode: <images>/<experiment directory name>/ when it exists)")
    ap.add_argument("--weights", help="id_module.th (default <exp_path>/id_module.th)")
    ap.add_argument("--dinov2", help="dinov2_vits14 state dict (sets SIXDGS_DINOV2_WEIGHTS; torch.hub is not reachable offline)")
    ap.add_argument("--out", default="results.json")
    ap.add_argument("--every", type=int, default=8, help="test split = every N-th camera (3DGS llffhold convention)")
    ap.add_argument("--max_ellipsoids", type=int, default=1000, help="1000 = the reference's cap (sampling.py:146-148); 0 = all")
    ap.add_argument("--score_impl", default="tc_f16x2", choices=["tc_f16x2", "simt_fp32", "tc_bf16"])
    ap.add_argument("--backbone", default="dino", choices=["dino", "synthetic"],
                    help="synthetic = the deterministic stand-in used by the fixtures (no DINOv2 weights offline)")
    ap.add_argument("--seed", type=int, default=55176280)  # pretrain_eval_attention.py:183
    ap.add_argument("--oracle_rays", action="store_true",
                    help="also run the reference driver's first pass (pretrain_eval_attention.py:100-120): poses from the "
                         "top-100 distance-based TARGET scores, score loss and top-100 recall of the predictions")
    ap.add_argument("--device", default="cuda", help="the kernels need a CUDA device; anything else only suits a stand-in package")
    args = ap.parse_args(argv)
    if not (args.exp_path or args.ply):
        ap.error("--exp_path or --ply is required")
    if args.dinov2:
        os.environ["SIXDGS_DINOV2_WEIGHTS"] = args.dinov2
    if sx is None:
        sx = importlib.import_module(__package__ or "6dgs_b200")
    dev = torch.device(args.device)
    torch.manual_seed(args.seed)
    if args.ply or os.path.isdir(os.path.join(args.exp_path, "point_cloud")):  # one experiment
        ply = args.ply or find_point_cloud(args.exp_path)
        out = evaluate_experiment(sx, args, dev, args.exp_path, ply, args.images,
                                  os.path.basename(os.path.normpath(args.exp_path or ply)))
    else:  # a directory of experiments, like the reference driver's --exp_path
        objects = parse_exp_dir(args.exp_path, DATA_TYPE_PREFIX.get(args.data_type, ""))
        if not objects:
            raise FileNotFoundError(f"no experiment with a point_cloud/iteration_*/point_cloud.ply under {args.exp_path}")
        out = {"results": [], "objects": {}}
        for obj in objects.values():
            name = os.path.basename(obj["exp_dir_filepath"])
            img_dir = args.images
            if img_dir and os.path.isdir(os.path.join(img_dir, name)):
                img_dir = os.path.join(img_dir, name)
            try:
                torch.manual_seed(args.seed)  # every object starts from the same seed (pretrain_eval_attention.py:41)
                one = evaluate_experiment(sx, args, dev, obj["exp_dir_filepath"], obj["checkpoint_filepath"], img_dir,
                                          obj["sequence_id"], obj["category_name"])
            except RuntimeError:  # the reference logs the object's failure and goes on (pretrain_eval_attention.py:240-241)
                traceback.print_exc()
                continue
            out["results"].extend(one.pop("results"))
            out["objects"][obj["sequence_id"]] = one
    with open(args.out, "w") as fh:
        json.dump(out, fh)
    print(f"[eval_driver] {len(out['results'])} frames -> {args.out}")
    return out


if __name__ == "__main__":
    main()
