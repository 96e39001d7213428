"""Target scores and the score loss of the reference's ``pose_estimation/distance_based_loss.py``:
``best_one_to_one_rays_selector`` (:5-144) and ``DistanceBasedScoreLoss`` (:147-283), same names, arguments and returns.

They are what ``train_id_module`` regresses the ray scores against (``pose_estimation/train.py:150-159``) and what the
"oracle rays" pass of the evaluation uses (``test.py:110-142``, ``pretrain_eval_attention.py:100-120``).  Everything here
is elementwise work over the rays with no reduction other than one sum and one mean, evaluated once per image on the
caller's device with torch ops -- boundary glue around the scored path, not a kernel of it; the gradient
flows through ``pred_score`` only (the targets are built under ``no_grad`` like upstream), into the score backward
kernels of ``identification._RayScoreFunction``.

Not mirrored: ``least_squared_loss.LeastSquaredLoss`` -- it calls ``best_one_to_one_rays_selector`` without importing
it (NameError on first use, least_squared_loss.py:33-47), i.e. dead upstream; its all-ray weighted solve is
``ops.score_pass2_batch(..., ls_rays=...)`` + ``ops.ls_solve``.
"""
from __future__ import annotations

from typing import Tuple

import torch

BACKBONE_RESIZE = 256  # backbone.py:90-101: short side to 256, centre crop 224, 14-pixel patches
BACKBONE_CROP = 224
PATCH = 14.0


def best_one_to_one_rays_selector(camera_intrinsic, camera_pose, obs_img_shape, rays_dir, rays_ori, backbone_wh,
                                  tanh_denominator=1.0):
    """-> (None, is_inside[n] bool, target_score[n], target_score_with_distance[n]).

    target_score = 1 - tanh(distance between the camera centre and the ray) for rays that start in front of the camera
    (positive depth along the camera z axis), 0 behind it; a ray pointing away from the camera is measured from its
    origin (distance_based_loss.py:14-57).  ``..._with_distance`` also fades with the origin's distance to the camera
    (:59-61).  ``is_inside``: the ray origin projects into the backbone's 16 x 16 patch grid after the 256-resize /
    224-crop (:63-105; upstream computes it, and a per-patch table it never fills, and returns None for the indices)."""
    centre = camera_pose[:3, 3][None]  # [0,0,0,1] @ pose[:3,:].T
    to_cam = centre - rays_ori
    along = (to_cam * rays_dir).sum(-1, keepdim=True)
    closest = torch.where(along < 0, rays_ori, rays_ori + along * rays_dir)
    target = 1 - torch.tanh(torch.linalg.norm(closest - centre, dim=-1) / tanh_denominator)
    depth = ((rays_ori - centre) * camera_pose[:3, 2][None]).sum(-1)  # along [0,0,1] @ pose[:3,:3].T
    target = target * ((depth / torch.abs(depth) + 1.0) / 2.0)  # NaN for a depth of exactly 0, as upstream
    target_with_distance = target * (1 - torch.tanh(torch.linalg.norm(to_cam, dim=-1) / tanh_denominator))

    proj = camera_intrinsic @ torch.linalg.inv(camera_pose)[:3, :]
    pix = (proj @ torch.cat((rays_ori.mT, torch.ones_like(rays_ori[:, :1]).mT), 0)).mT  # homogeneous, like upstream
    pix = pix[:, :2] / pix[:, 2:]
    s = BACKBONE_RESIZE / (obs_img_shape[0] if obs_img_shape[0] < obs_img_shape[1] else obs_img_shape[1])
    off = torch.tensor([((s * obs_img_shape[0]) - BACKBONE_CROP) // 2, ((s * obs_img_shape[1]) - BACKBONE_CROP) // 2],
                       dtype=pix.dtype, device=pix.device)
    pix = (pix * s - off) / PATCH
    is_inside = (pix[:, 1] >= 0.0) & (pix[:, 1] <= backbone_wh[1]) & (pix[:, 0] >= 0.0) & (pix[:, 0] <= backbone_wh[0])
    return None, is_inside, target, target_with_distance


class DistanceBasedScoreLoss(torch.nn.Module):
    """MSE between the predicted ray scores and the distance-based targets rescaled to the same total mass
    (``total_number_of_features`` = n_img: the predicted scores of one image sum to its token count).
    The re-weighting / LDS options are accepted and, as upstream, unused by ``forward``."""

    def __init__(self, reweight_method="none", lds=False, lds_kernel="gaussian", lds_ks=5, lds_sigma=2,
                 total_number_of_elements: float = 256.0):
        super().__init__()
        assert reweight_method in {"none", "inverse", "sqrt_inv"}
        assert reweight_method != "none" if lds else True, "Set reweight to 'sqrt_inv' (default) or 'inverse' when using LDS"
        self.reweight_method, self.lds, self.lds_kernel, self.lds_ks, self.lds_sigma = (reweight_method, lds, lds_kernel,
                                                                                        lds_ks, lds_sigma)

    def forward(self, pred_score: torch.Tensor, camera_pose: torch.Tensor, camera_intrinsic: torch.Tensor,
                rays_ori: torch.Tensor, rays_dir: torch.Tensor, total_number_of_features: int,
                backbone_wh: Tuple[int, int], model_up=None, obs_img_shape=(800, 800)):
        """-> (mean squared score error, the target scores it was measured against [n])"""
        with torch.no_grad():
            _, _, target, _ = best_one_to_one_rays_selector(camera_intrinsic, camera_pose, obs_img_shape, rays_dir,
                                                            rays_ori, backbone_wh=backbone_wh, tanh_denominator=1.0)
            target = target * (total_number_of_features / target.sum())
        return torch.square(pred_score - target).mean(), target
