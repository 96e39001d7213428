"""Seeded synthetic inputs: N-Gaussian scenes, identification-module weights,
query images and a deterministic stand-in for the DINOv2 backbone (the real weights are a
torch.hub download, reference pose_estimation/backbone.py:15, unavailable offline).

Everything is generated on CPU from an explicit ``torch.Generator`` so the same seed gives the
same tensors on the build container and on the GPU box.
"""
from __future__ import annotations

from typing import Dict

import torch

# parameter shapes of the reference IdentificationModule that live on the hot path
# (reference ray_preprocessor.py:15-19, our_multihead_attention.py:58-59, camera_direction_network.py:21-43)
ID_MODULE_SHAPES = {
    "ray_preprocessor.mlp.0": (512, 141),
    "ray_preprocessor.mlp.2": (512, 512),
    "ray_preprocessor.mlp2.0": (512, 653),
    "ray_preprocessor.mlp2.2": (384, 512),
    "attention.q_proj": (384, 398),
    "attention.k_proj": (384, 384),
    "camera_direction_prediction_network.dim_reducer1.0": (384, 384, 5, 5),
    "camera_direction_prediction_network.dim_reducer1.2": (384, 384, 5, 5),
    "camera_direction_prediction_network.dim_reducer1.4": (384, 384, 5, 5),
    "camera_direction_prediction_network.dim_reducer2.0": (384, 384, 4, 4),
    "camera_direction_prediction_network.mlp.0": (256, 384),
    "camera_direction_prediction_network.mlp.2": (3, 256),
}


def synth_scene(n: int, seed: int = 0, extent: float = 1.0, heavy_tail: bool = False,
                sh_degree: int = 3) -> Dict[str, torch.Tensor]:
    """xyz ~ N(0,I)*extent; log-scale = log U(0.005,0.055) (or exp N(-4,1) heavy tail);
    quaternion ~ N(0,I); SH dc ~ 0.3 N, rest ~ 0.05 N."""
    g = torch.Generator().manual_seed(seed)
    xyz = torch.randn(n, 3, generator=g) * extent
    if heavy_tail:
        scaling = -4.0 + torch.randn(n, 3, generator=g)
    else:
        scaling = torch.log(torch.rand(n, 3, generator=g) * 0.05 + 0.005)
    rot = torch.randn(n, 4, generator=g)
    f_dc = 0.3 * torch.randn(n, 1, 3, generator=g)
    f_rest = 0.05 * torch.randn(n, 15, 3, generator=g)
    return {"xyz": xyz, "scaling": scaling, "rotation": rot, "features_dc": f_dc,
            "features_rest": f_rest, "sh_degree": sh_degree}


def synth_id_weights(seed: int = 0, gain: float = 1.0, q_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Identification-module state dict with U(-b, b), b = gain*sqrt(3/fan_in) weights (unit-ish
    variance propagation) and small biases; keys match the reference ``id_module.th`` names.
    ``q_gain`` scales ``attention.q_proj`` (weight and bias) afterwards: the random-init logits are nearly flat
    (std 0.3); q_gain = 20 gives std ~ 6.5, i.e. a peaked, trained-looking softmax over the rays."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in ID_MODULE_SHAPES.items():
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = gain * (3.0 / fan_in) ** 0.5
        sd[name + ".weight"] = (torch.rand(*shape, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(shape[0], generator=g) * 2 - 1) * 0.05
    if q_gain != 1.0:
        sd["attention.q_proj.weight"] = sd["attention.q_proj.weight"] * q_gain
        sd["attention.q_proj.bias"] = sd["attention.q_proj.bias"] * q_gain
    return sd


def synth_image(h: int, w: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(h, w, 3, generator=g)


class SyntheticBackbone(torch.nn.Module):
    """Deterministic stand-in with DINOv2's ``forward_features`` contract: 14x14 patches of the
    224x224 crop -> fixed random projection -> per-token standardisation -> [1, 256, 384]."""

    def __init__(self, dim: int = 384, patch: int = 14, seed: int = 1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        w = torch.randn(dim, 3 * patch * patch, generator=g) / (3 * patch * patch) ** 0.5
        self.register_buffer("proj", w)
        self.patch = patch
        self.dim = dim

    def forward_features(self, x: torch.Tensor):
        b, c, h, w = x.shape
        p = self.patch
        patches = x.unfold(2, p, p).unfold(3, p, p)
        gh, gw = patches.shape[2], patches.shape[3]
        patches = patches.permute(0, 2, 3, 1, 4, 5).reshape(b, gh * gw, c * p * p)
        tok = patches @ self.proj.t()
        tok = (tok - tok.mean(-1, keepdim=True)) / (tok.std(-1, keepdim=True) + 1e-6)
        return {"x_norm_patchtokens": tok}
