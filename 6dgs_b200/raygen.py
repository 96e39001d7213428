"""Scene -> candidate rays.  Same entry point as the reference
``pose_estimation/sampling.py:generate_all_possible_rays`` (:127-267).

Differences that matter to callers:
  * the reference silently caps the scene at 1000 randomly chosen ellipsoids (sampling.py:146-148)
    because its cell search materialises a [cells, 1000] table; the kernels here have no such
    temporary, so ``max_ellipsoids`` is a keyword (default 1000 = reference behaviour, ``None`` =
    every valid ellipsoid);
  * the random subset can be injected (``ellipsoid_idx``, indices into the VALID subset exactly
    like the reference's randperm) or seeded (``generator``) for reproducible parity runs.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


@torch.no_grad()
def generate_all_possible_rays(model, num_viewdirs_per_chunk: int = 10240, sample_quadricell_targets: int = 50, *,
                               max_ellipsoids: Optional[int] = 1000, ellipsoid_idx: Optional[torch.Tensor] = None,
                               generator: Optional[torch.Generator] = None, k_neighbors: int = 20,
                               resolution: int = 1000, return_ids: bool = False, normals_cloud: Optional[torch.Tensor] = None,
                               shard: Optional[tuple] = None):
    """-> (rays_ori[Nr,3], rays_dir[Nr,3], rays_rgb[Nr,3]) fp32 on the model's device.

    ``num_viewdirs_per_chunk`` is accepted for signature compatibility (the reference chunks its SH
    evaluation, sampling.py:225-251; the fill kernel evaluates SH per ray in-register).
    ``shard=(rank, world)`` generates only this rank's contiguous block of the selected ellipsoids
    (normals still use the full selected cloud).  ``return_ids`` appends the global
    Gaussian id of every ray."""
    dev = model._xyz.device
    xyz, scaling, rotation, feats = model._xyz, model._scaling, model._rotation, model.get_features
    valid, _ = ops.degrade_mask(scaling, sample_quadricell_targets)
    valid_ids = torch.nonzero(valid).squeeze(1)  # device->host sync on the count, as sampling.py:145
    nvalid = valid_ids.shape[0]
    if ellipsoid_idx is None:
        cap = nvalid if max_ellipsoids is None else min(max_ellipsoids, nvalid)
        if max_ellipsoids is None and generator is None:
            ellipsoid_idx = torch.arange(nvalid, dtype=torch.long, device=dev)  # all of them: order is irrelevant
        else:
            gdev = generator.device if generator is not None else dev
            ellipsoid_idx = torch.randperm(nvalid, dtype=torch.long, device=gdev, generator=generator)[:cap].to(dev)
    else:
        ellipsoid_idx = ellipsoid_idx.to(device=dev, dtype=torch.long)
    sel = valid_ids[ellipsoid_idx].contiguous()
    centers = xyz[sel].contiguous()
    m = sel.shape[0]
    if m == 0:  # every ellipsoid degraded (or an empty scene): no rays
        e = torch.empty(0, 3, dtype=torch.float32, device=dev)
        return (e, e.clone(), e.clone(), sel) if return_ids else (e, e.clone(), e.clone())
    lo, hi = 0, m
    if shard is not None:
        rank, world = shard
        per = (m + world - 1) // world
        lo, hi = min(rank * per, m), min((rank + 1) * per, m)
    cloud = centers if normals_cloud is None else normals_cloud
    normals = ops.knn_normals(cloud, k_neighbors, lo, hi - lo)
    sel_local = sel[lo:hi].contiguous()
    ori, dirs, rgb, ell, _ = ops.raygen(xyz, scaling, rotation, feats, model.active_sh_degree, sel_local, normals,
                                        sample_quadricell_targets, resolution, mode=0)
    if return_ids:
        return ori, dirs, rgb, sel_local[ell]
    return ori, dirs, rgb


def quadricell_cells(a_b_c_log: torch.Tensor, target_points: int = 50, resolution: int = 1000):
    """a6 in isolation (reference quadricell.py:191-319): log semi-axes [M,3] -> (points[C,3], ellipsoid_id[C])."""
    s = a_b_c_log.contiguous()
    sel = torch.arange(s.shape[0], dtype=torch.long, device=s.device)
    pts, _, _, ell, _ = ops.raygen(None, s, None, None, 0, sel, None, target_points, resolution, mode=1)
    return pts, ell
