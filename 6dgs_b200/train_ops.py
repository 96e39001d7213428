"""Training-mode building blocks on the libsixdgs GEMM.

The reference trains the identification module through torch autograd (pose_estimation/train.py:146-176:
``id_module(...)`` -> ``loss.backward()``): the ray MLP (ray_preprocessor.py:36-46), the two projections
(our_multihead_attention.py:74-75) and the softmax-over-rays score.  The score already runs forward and backward
on the kernels (``identification._RayScoreFunction``); this file does the same for the linear layers:

    y  = act(x W^T + b)              forward   -> one sixdgs_linear launch
    dx = (dy . act') W               backward  -> one sixdgs_linear launch on W^T
    dW = (dy . act')^T x             backward  -> sixdgs_linear over ray chunks (the reduction dim is the ray axis)
    db = sum_rows (dy . act')        backward  -> a column sum

Every GEMM is the fp32 FMA kernel behind ``sixdgs_linear`` (csrc/features.cu: linear_kernel), called only with
shapes it is exercised with on the query path: reduction dims padded to a multiple of 16, output widths padded to a
multiple of 128 (zero rows of the weight operand), row counts arbitrary.  torch supplies memory, transposes, the
ReLU mask and the bias column sum.  There is no CPU path: ``ops.linear`` rejects non-CUDA tensors.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

K_ALIGN = 16    # sixdgs_linear: the reduction dim must be a multiple of 16
N_ALIGN = 128   # one CTA tile of output columns; full tiles only (see the module docstring)
RAY_CHUNK = 1 << 17  # rays per dW GEMM (two [<=768, 131072] fp32 staging buffers)


def _round_up(v: int, a: int) -> int:
    return -(-v // a) * a


def _padded(t: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """contiguous fp32 [rows, cols] copy of the 2-D tensor t, zero padded (no copy when nothing changes)."""
    t = t.detach().to(torch.float32)
    if t.shape == (rows, cols):
        return t.contiguous()
    out = torch.zeros(rows, cols, dtype=torch.float32, device=t.device)
    out[: t.shape[0], : t.shape[1]] = t
    return out


def gemm_nt(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """act(x w^T + b) for x [m,k], w [n,k] of any k, n (padded to the kernel's alignment here) -> contiguous [m,n]."""
    m, k = x.shape
    n = w.shape[0]
    if w.shape[1] != k:
        raise ValueError(f"gemm_nt: x is [{m},{k}] but w is {list(w.shape)}")
    if m == 0:
        return torch.zeros(0, n, dtype=torch.float32, device=x.device)
    kp, np_ = _round_up(k, K_ALIGN), _round_up(n, N_ALIGN)
    bp = None
    if b is not None:
        bp = _padded(b.reshape(1, -1), 1, np_).reshape(-1)
    y = ops.linear(_padded(x, m, kp), _padded(w, np_, kp), bp, relu)
    return y if np_ == n else y[:, :n].contiguous()


def gemm_tn(a: torch.Tensor, b: torch.Tensor, chunk: int = RAY_CHUNK) -> torch.Tensor:
    """a^T b for a [m,na], b [m,nb] with a long row axis m (rays) -> [na,nb]: per chunk of rows both operands are
    transposed into zero-padded [*, chunk] buffers and multiplied by the same NT kernel (the reduction runs over
    the ray axis), partial products added in chunk order (deterministic)."""
    m, na = a.shape
    nb = b.shape[1]
    if b.shape[0] != m:
        raise ValueError("gemm_tn: row counts differ")
    out = torch.zeros(na, nb, dtype=torch.float32, device=a.device)
    nbp = _round_up(nb, N_ALIGN)
    for lo in range(0, m, chunk):
        hi = min(m, lo + chunk)
        cp = _round_up(hi - lo, K_ALIGN)
        at = _padded(a[lo:hi].t(), na, cp)
        bt = _padded(b[lo:hi].t(), nbp, cp)
        out += ops.linear(at, bt, None)[:, :nb]
    return out


class _LinearFunction(torch.autograd.Function):
    """torch.nn.functional.linear (+ optional ReLU) with forward and backward on sixdgs_linear."""

    @staticmethod
    def forward(ctx, x, w, b, relu: bool):
        y = gemm_nt(x, w, b, relu)
        ctx.relu = bool(relu)
        ctx.dtypes = (x.dtype, w.dtype, b.dtype if b is not None else None)  # gradients go back in the inputs' dtypes
        ctx.save_for_backward(x.detach(), w.detach(), y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        gy = gy.to(torch.float32)
        if ctx.relu:
            gy = gy * (y > 0)  # the ReLU's mask from its own output (the reference rectifies in place)
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm_nt(gy, w.t()).to(ctx.dtypes[0])          # [m, n] x [k, n]^T
        if ctx.needs_input_grad[1]:
            gw = gemm_tn(gy, x).to(ctx.dtypes[1])              # [n, k]
        if ctx.needs_input_grad[2]:
            gb = gy.sum(0).to(ctx.dtypes[2])
        return gx, gw, gb, None


def linear(x: torch.Tensor, layer: torch.nn.Linear, relu: bool = False) -> torch.Tensor:
    return _LinearFunction.apply(x, layer.weight, layer.bias, relu)


def positional_encoding(p: torch.Tensor, freqs: int) -> torch.Tensor:
    """[sin(p 2^f), cos(p 2^f)], coordinate-major, no pi (ray_preprocessor.py:3-9)."""
    ang = (p[..., None] * (2.0 ** torch.arange(freqs, device=p.device, dtype=p.dtype))).reshape(p.shape[0], -1)
    return torch.cat((torch.sin(ang), torch.cos(ang)), -1)


def ray_mlp_input(rp, rays_ori, rays_dir, rays_rgb) -> torch.Tensor:
    """the 141-wide MLP input (ray_preprocessor.py:36-44); rays carry no gradient"""
    return torch.cat((rays_ori, rays_dir, rays_rgb, positional_encoding(rays_ori, rp.pospe),
                      positional_encoding(rays_dir, rp.viewpe), positional_encoding(rays_rgb, rp.rgbpe)), -1)


def ray_keys(rp, attention, rays_ori, rays_dir, rays_rgb) -> torch.Tensor:
    """rays -> PE -> mlp -> [h, x] -> mlp2 -> k_proj (ray_preprocessor.py:36-46, our_multihead_attention.py:75),
    differentiable w.r.t. the ten weight / bias tensors, every GEMM of both directions on the kernels."""
    x = ray_mlp_input(rp, rays_ori, rays_dir, rays_rgb)
    h = linear(linear(x, rp.mlp[0], True), rp.mlp[2], True)
    fea = linear(linear(torch.cat((h, x), -1), rp.mlp2[0], True), rp.mlp2[2], False)
    return linear(fea, attention.k_proj, False)


def image_queries(attention, tokens_pe: torch.Tensor) -> torch.Tensor:
    """q = Wq [f_img | pe] + bq (our_multihead_attention.py:74)"""
    return linear(tokens_pe, attention.q_proj, False)
