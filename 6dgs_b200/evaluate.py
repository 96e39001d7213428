"""Evaluation loop with the signature and result schema of the reference
``pose_estimation/test.py:test_pose_estimation`` (:23-323).  Host glue stays Python; per camera the
device work is: backbone -> q -> two key-cache passes -> top-k -> one fused pose-tail launch
(dedup, LS x2, exclude-negatives, watch direction, rotation, singular/NaN conventions)."""
from __future__ import annotations

import math
import time
from statistics import mean
from typing import List

import numpy as np
import torch

from . import ops


def compute_translation_error(t1, t2):
    return torch.linalg.norm(t1 - t2)


def compute_angular_error(rotation_gt, rotation_est):
    cos_angle = (torch.trace(rotation_gt @ torch.linalg.inv(rotation_est)) - 1) / 2
    return torch.rad2deg(torch.arccos(torch.clamp(cos_angle, min=-1, max=1)))


def fov2focal(fov, pixels):
    return pixels / (2 * math.tan(fov / 2))


def camera_to_query(camera_info, device):
    """CameraInfo -> (gt c2w[4,4], K[3,3], image [H,W,3] in [0,1], mask [H,W] bool) (test.py:47-83)."""
    w2c = torch.eye(4, dtype=torch.float32, device=device)
    w2c[:3, :3] = torch.from_numpy(np.asarray(camera_info.R)).T.to(device)
    w2c[:3, -1] = torch.from_numpy(np.asarray(camera_info.T)).to(device)
    pose = torch.inverse(w2c)
    fx, fy = fov2focal(camera_info.FovX, camera_info.width), fov2focal(camera_info.FovY, camera_info.height)
    K = torch.tensor([[fx, 0.0, camera_info.width / 2], [0.0, fy, camera_info.height / 2], [0.0, 0.0, 1.0]],
                     dtype=torch.float32, device=device)
    img = torch.from_numpy(np.array(camera_info.image)).to(device=device, dtype=torch.float32) / 255.0
    if img.shape[-1] == 4:
        mask = img[..., -1] > 0.3
        img = img[..., :3] * img[..., -1:] + (1 - img[..., -1:])
    else:
        mask = torch.ones_like(img[..., -1], dtype=torch.bool)
    return pose, K, img, mask


def test_pose_estimation(cameras_info: List, id_module, rays_ori, rays_dirs, rays_rgb, model_up, sequence_id="",
                         category_id="", loss_fn=None, save=False, save_all=False):
    """-> (results, avg_translation_error, avg_angular_error, avg_loss_score, avg_recall)."""
    id_module.eval()
    if save or save_all:
        raise NotImplementedError("the reference's debug dump writes to a hard-coded home path (test.py:209-212)")
    dev = rays_ori.device
    t_errs, a_errs, recalls, losses, results = [], [], [], [], []
    start = time.time()
    for img_idx, cam in enumerate(cameras_info):
        pose, K, img, mask = camera_to_query(cam, dev)
        idx, weights, pred_scores, up, amap = id_module.test_image(img, mask, rays_ori, rays_dirs, rays_rgb,
                                                                   rays_to_output=100)
        avg_score, recall = -1.0, -1.0
        if loss_fn is not None:  # "oracle rays" mode (test.py:110-142)
            n_img = int(id_module.backbone_wrapper(img, mask)[0].shape[0]) if amap is None else amap.shape[-2]
            avg_score_t, target_scores = loss_fn(pred_scores, pose, K, rays_ori, rays_dirs, n_img,
                                                 id_module.backbone_wrapper.backbone_wh, model_up=up)
            avg_score = avg_score_t.item()
            target_idx = torch.topk(weights, k=100).indices
            recall = torch.count_nonzero(torch.isin(target_idx, idx)).item() / target_idx.shape[0]
            weights, idx = torch.topk(target_scores, k=100, largest=True)
            weights, idx = weights.contiguous().float(), idx.contiguous()
        losses.append(avg_score)
        recalls.append(recall)
        c2w, aux = ops.pose_tail(rays_ori, rays_dirs, idx, weights, up)
        aux_h = aux.cpu()
        status = int(aux_h[7])
        if status & 1:
            print("camera_optical_center is nan")
        if status & 2:
            print("extracted rotation matrix is singular")
        if status & 4:
            print("wrong c2w")
        origin = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=pose.dtype, device=dev).reshape(1, 4)
        t_errs.append(compute_translation_error(origin @ pose[:3, :].T, origin @ c2w[:3, :].T).item())
        a_errs.append(compute_angular_error(pose[:3, :3], c2w[:3, :3]).item())
        # "loss" is the mean of the final (masked, renormalised) weights = 1 / n_kept (test.py:296)
        results.append({"sequence_id": sequence_id, "category_name": category_id, "frame_id": img_idx,
                        "loss": 1.0 / max(float(aux_h[6]), 1.0), "scores_loss": avg_score, "recall": recall,
                        "total_optimization_time_in_ms": 0.0, "pred_c2w": c2w.cpu().tolist(),
                        "gt_c2w": pose.cpu().tolist()})
    per = (time.time() - start) / max(len(cameras_info), 1)
    print("Average loss score: ", mean(losses))
    print("Average Recall: ", mean(recalls))
    print("Time per element: ", per)
    print("Translation Error: ", mean(t_errs))
    print("Angular Error: ", mean(a_errs))
    return results, mean(t_errs), mean(a_errs), mean(losses), mean(recalls)
