"""``sym_eig_3x3`` with the reference signature (pose_estimation/sym_eig_3x3.py:246-248)."""
from typing import Optional

import torch

from . import ops


def sym_eig_3x3(inputs: torch.Tensor, eigenvectors: bool = True, eps: Optional[float] = None):
    if inputs.shape[-2:] != (3, 3):
        raise ValueError("Only inputs of shape (..., 3, 3) are supported.")
    batch = inputs.shape[:-2]
    vals, vecs = ops.sym_eig3x3(inputs, eigenvectors, eps)
    vals = vals.reshape(*batch, 3)
    return vals, (vecs.reshape(*batch, 3, 3) if eigenvectors else None)
