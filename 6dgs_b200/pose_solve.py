"""Least-squares ray intersection and the rotation helpers, with the reference names and
conventions (pose_estimation/line_intersection.py:5-34,75-154)."""
from typing import Optional

import torch

from . import ops


def compute_line_intersection_impl2(points: torch.Tensor, directions: torch.Tensor,
                                    weights: Optional[torch.Tensor] = None, return_residuals: bool = False):
    """LS intersection of n lines; NaN vector when det(R) < 1e-7 (line_intersection.py:139-142)."""
    if return_residuals:
        # the reference reads `.residuals` off a torch.linalg.solve result, which does not exist
        # (line_intersection.py:151-152): the branch can only raise there too.
        raise AttributeError("'Tensor' object has no attribute 'residuals'")
    centre, _ = ops.line_intersect(points, directions, weights)
    return centre


def exclude_negatives(camera_optical_center: torch.Tensor, sample_points: torch.Tensor, dirs: torch.Tensor):
    """rays whose direction points towards the centre (line_intersection.py:29-34).  <=100 elements
    on the live path (test.py:174): plain tensor algebra on the caller's device, no kernel."""
    return ((camera_optical_center[None] - sample_points) * dirs).sum(-1) > 0


def make_rotation_mat(direction: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """rows [x; y; direction]; returned on the CPU like the reference (line_intersection.py:12-26)."""
    xa = torch.linalg.cross(up, direction)
    xa = xa / torch.linalg.norm(xa, dim=-1, keepdim=True)
    ya = torch.linalg.cross(direction, xa)
    ya = ya / torch.linalg.norm(ya, dim=-1, keepdim=True)
    return torch.stack((xa, ya, direction), 0).cpu()


def pose_from_topk(rays_ori, rays_dir, idx, vals, camera_up):
    """Fused pose tail (test.py:157-198) -> (c2w[4,4], aux[8]); one launch instead of ~40 tiny ops."""
    return ops.pose_tail(rays_ori, rays_dir, idx, vals, camera_up)
