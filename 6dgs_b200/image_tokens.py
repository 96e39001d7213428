"""Image side of the query (boundary only -- stays PyTorch): preprocessing, backbone
patch tokens, 14-channel positional encoding and mask handling, with the contract of the reference
``BackboneWrapper`` (pose_estimation/backbone.py:29-139).

The reference downloads DINOv2 ViT-S/14 through torch.hub (backbone.py:15).  There is no network
here, so ``create_backbone`` builds the same architecture locally (``DinoV2ViTS14``; weights from
``$SIXDGS_DINOV2_WEIGHTS`` if set, random otherwise) or takes any injected module that exposes
``forward_features(x)["x_norm_patchtokens"]``.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Optional

import torch
import torch.nn.functional as F
from torchvision import transforms

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class _Block(torch.nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4):
        super().__init__()
        self.norm1 = torch.nn.LayerNorm(dim, eps=1e-6)
        self.attn = torch.nn.Module()
        self.attn.qkv = torch.nn.Linear(dim, dim * 3)
        self.attn.proj = torch.nn.Linear(dim, dim)
        self.ls1 = torch.nn.Module()
        self.ls1.gamma = torch.nn.Parameter(torch.ones(dim))
        self.norm2 = torch.nn.LayerNorm(dim, eps=1e-6)
        self.mlp = torch.nn.Module()
        self.mlp.fc1 = torch.nn.Linear(dim, dim * mlp_ratio)
        self.mlp.fc2 = torch.nn.Linear(dim * mlp_ratio, dim)
        self.ls2 = torch.nn.Module()
        self.ls2.gamma = torch.nn.Parameter(torch.ones(dim))
        self.heads = heads

    def forward(self, x):
        b, n, c = x.shape
        qkv = self.attn.qkv(self.norm1(x)).reshape(b, n, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(b, n, c)
        x = x + self.ls1.gamma * self.attn.proj(a)
        return x + self.ls2.gamma * self.mlp.fc2(F.gelu(self.mlp.fc1(self.norm2(x))))


class DinoV2ViTS14(torch.nn.Module):
    """ViT-S/14 with DINOv2's parameter names (cls_token, pos_embed[1,1+37*37,384], patch_embed.proj,
    blocks.N.{norm1,attn.qkv,attn.proj,ls1.gamma,norm2,mlp.fc1,mlp.fc2,ls2.gamma}, norm)."""

    def __init__(self, dim=384, depth=12, heads=6, patch=14, base_grid=37, interpolate_offset=0.1):
        super().__init__()
        self.interpolate_offset = interpolate_offset  # hub value for dinov2_vits14; 0 / None = resample to an exact size
        self.patch_embed = torch.nn.Module()
        self.patch_embed.proj = torch.nn.Conv2d(3, dim, patch, patch)
        self.cls_token = torch.nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = torch.nn.Parameter(torch.randn(1, 1 + base_grid * base_grid, dim) * 0.02)
        self.mask_token = torch.nn.Parameter(torch.zeros(1, dim))
        self.blocks = torch.nn.ModuleList([_Block(dim, heads) for _ in range(depth)])
        self.norm = torch.nn.LayerNorm(dim, eps=1e-6)
        self.base_grid = base_grid
        self._pos_cache = None

    def _pos(self, gh, gw):
        if gh == self.base_grid and gw == self.base_grid:
            return self.pos_embed
        # the 37x37 -> 16x16 bicubic resample of the position table only depends on the weights: cache it
        # (torch runs this interpolate as a single 1024-thread block, ~1 ms per call on B200)
        key = (gh, gw, self.pos_embed.data_ptr(), self.pos_embed._version, self.pos_embed.device)
        if self._pos_cache is None or self._pos_cache[0] != key or torch.is_grad_enabled() and self.pos_embed.requires_grad:
            cls, grid = self.pos_embed[:, :1], self.pos_embed[:, 1:]
            g = grid.reshape(1, self.base_grid, self.base_grid, -1).permute(0, 3, 1, 2)
            if self.interpolate_offset:
                # the torch.hub dinov2_vits14 the reference loads (backbone.py:15) is built with interpolate_offset=0.1,
                # interpolate_antialias=False: the table is resampled by SCALE FACTOR (g + 0.1) / 37, not to a size --
                # bicubic with a scale factor samples at slightly different source coordinates than size=(g, g)
                sf = ((gh + self.interpolate_offset) / self.base_grid, (gw + self.interpolate_offset) / self.base_grid)
                g = F.interpolate(g, scale_factor=sf, mode="bicubic", align_corners=False)
                assert g.shape[-2:] == (gh, gw)
            else:
                g = F.interpolate(g, size=(gh, gw), mode="bicubic", align_corners=False)
            pos = torch.cat((cls, g.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)), 1)
            if torch.is_grad_enabled() and self.pos_embed.requires_grad:
                return pos
            self._pos_cache = (key, pos.detach())
        return self._pos_cache[1]

    def forward_features(self, x):
        t = self.patch_embed.proj(x)
        gh, gw = t.shape[-2:]
        t = t.flatten(2).transpose(1, 2)
        t = torch.cat((self.cls_token.expand(t.shape[0], -1, -1), t), 1) + self._pos(gh, gw)
        for blk in self.blocks:
            t = blk(t)
        t = self.norm(t)
        return {"x_norm_clstoken": t[:, 0], "x_norm_patchtokens": t[:, 1:]}


def create_backbone(type: str = "dino", backbone: Optional[torch.nn.Module] = None, **kwargs):
    """-> (model, (grid_h, grid_w), num_features), the tuple of reference backbone.py:6-22."""
    if backbone is not None:
        return backbone, (16, 16), 384
    if type != "dino":
        raise NotImplementedError("only the DINOv2 ViT-S/14 backbone is on the supported path "
                                  "(SuperPoint, reference backbone.py:18-21, is unreachable from the entry points)")
    model = DinoV2ViTS14()
    path = os.environ.get("SIXDGS_DINOV2_WEIGHTS")
    if path:
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
    else:
        warnings.warn("DINOv2 weights unavailable offline: ViT-S/14 is randomly initialised "
                      "(set SIXDGS_DINOV2_WEIGHTS to a dinov2_vits14 state dict)")
    return model, (16, 16), 384


class BackboneWrapper(torch.nn.Module):
    def __init__(self, backbone_type: str = "dino", backbone: Optional[torch.nn.Module] = None) -> None:
        super().__init__()
        self.image_preprocessing_net, self.backbone_wh, self.img_num_features = create_backbone(backbone_type, backbone)
        self.norm_mean = torch.nn.Parameter(torch.tensor(IMAGENET_MEAN, dtype=torch.float32), requires_grad=False)
        self.norm_std = torch.nn.Parameter(torch.tensor(IMAGENET_STD, dtype=torch.float32), requires_grad=False)
        bicubic, bilinear = transforms.InterpolationMode.BICUBIC, transforms.InterpolationMode.BILINEAR
        # Normalize is applied with the registered norm_mean/norm_std parameters: torchvision's Normalize
        # builds its mean/std tensors from Python tuples on every call (a pageable H2D copy, which also
        # cannot be captured in a CUDA graph)
        self.transformations = transforms.Compose([
            transforms.Resize(256, interpolation=bicubic, antialias=True), transforms.CenterCrop(224)])
        self.mask_transformations = transforms.Compose([
            transforms.Resize(256, interpolation=bilinear, antialias=True), transforms.CenterCrop(224),
            transforms.Resize(self.backbone_wh[0], interpolation=bilinear, antialias=True)])

    @staticmethod
    def get_img_position_encoding(img_features_shape, freqs, dtype=torch.float32, device="cpu"):
        """[x, y, sin(x*2^f) sin(y*2^f) (coordinate-major), cos(...)] on linspace(-1,1) (backbone.py:116-139)."""
        shape = tuple(img_features_shape)
        axes = [torch.linspace(-1.0, 1.0, steps=s, dtype=dtype, device=device) for s in shape]
        pos = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, len(shape))
        bands = 2.0 ** torch.arange(freqs, dtype=dtype, device=pos.device)
        ang = (pos[..., None] * bands).reshape(pos.shape[0], -1)
        return torch.cat([pos, torch.sin(ang), torch.cos(ang)], dim=-1).reshape(*shape, -1)

    def tokens_dense_batch(self, imgs: torch.Tensor, masks: torch.Tensor):
        """Batched, sync-free front end: imgs [B,H,W,3] in [0,1], masks [B,H,W] bool ->
        (tokens+pe [B,16,16,398], tokens [B,16,16,384], keep [B,16,16] bool).  One backbone pass for the batch."""
        x = self.transformations(imgs.permute(0, 3, 1, 2))
        x = (x - self.norm_mean.view(1, 3, 1, 1)) / self.norm_std.view(1, 3, 1, 1)
        keep = self.mask_transformations(masks[:, None] * 1.0)[:, 0] > 0.1
        gh, gw = self.backbone_wh
        b = imgs.shape[0]
        tok = self.image_preprocessing_net.forward_features(x)["x_norm_patchtokens"].reshape(b, gh, gw, self.img_num_features)
        pe = self.get_img_position_encoding((gh, gw), 3, dtype=imgs.dtype, device=imgs.device)
        return torch.cat((tok, pe[None].expand(b, -1, -1, -1)), dim=-1), tok, keep

    def tokens_dense(self, img: torch.Tensor, mask: torch.Tensor):
        """All 256 grid tokens plus their validity, no host synchronisation:
        -> (tokens+pe [16,16,398], tokens [16,16,384], keep [16,16] bool)."""
        tok_pe, tok, keep = self.tokens_dense_batch(img[None], mask[None])
        return tok_pe[0], tok[0], keep[0]

    def forward(self, img: torch.Tensor, mask: torch.Tensor):
        """img [H,W,3] in [0,1], mask [H,W] bool -> (tokens+pe [n_img,398], tokens [n_img,384], grid [384,16,16]).
        Boolean-mask compaction like the reference (backbone.py:110-114): one host sync on n_img."""
        tok_pe, tok, keep = self.tokens_dense(img, mask)
        return tok_pe[keep].view(-1, tok_pe.shape[-1]), tok[keep].view(-1, tok.shape[-1]), tok.permute(2, 0, 1)
