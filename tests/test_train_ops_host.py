"""Host-side check of the training-mode linear layers (6dgs_b200/train_ops.py; reference: the autograd of
pose_estimation/ray_preprocessor.py:36-46 + our_multihead_attention.py:74-75 as driven by train.py:146-176).

No GPU here, so ``ops.linear`` -- the ctypes wrapper of ``sixdgs_linear`` -- is replaced by a stand-in that
(i) implements the kernel's contract  y[m,n] = act(x[m,k] w[n,k]^T + b)  and (ii) REJECTS every call outside the
envelope the kernel has been run with on a B200 (contiguous fp32, k a multiple of 16, full 128-column output tiles,
bias of the padded width).  What is verified is therefore the host logic around the kernel: padding, transposes, the
chunked ray-axis reduction of dW, the ReLU masks, which gradients are formed.  The GPU twin of this test is
tests/test_zz_gpu_training_mlp.py."""
import importlib

import pytest
import torch


@pytest.fixture()
def train_ops(monkeypatch):
    ops = importlib.import_module("6dgs_b200.ops")
    calls = []

    def linear_contract(x, w, b, relu=False):
        assert x.dtype == w.dtype == torch.float32 and x.is_contiguous() and w.is_contiguous()
        assert x.dim() == w.dim() == 2 and x.shape[1] == w.shape[1]
        m, k = x.shape
        n = w.shape[0]
        assert m > 0 and k % 16 == 0 and n % 128 == 0, (m, k, n)
        if b is not None:
            assert b.dtype == torch.float32 and b.is_contiguous() and b.shape == (n,)
        calls.append((m, k, n))
        y = x.double() @ w.double().t()
        if b is not None:
            y = y + b.double()
        if relu:
            y = y.clamp_min(0)
        return y.float()

    monkeypatch.setattr(ops, "linear", linear_contract)
    mod = importlib.import_module("6dgs_b200.train_ops")
    mod._calls = calls
    return mod


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("m,k,n,relu,bias", [(1, 141, 512, True, True), (300, 653, 512, True, True),
                                             (257, 512, 384, False, True), (40, 398, 384, False, False),
                                             (5, 7, 3, True, True)])
def test_linear_function_vs_fp64_autograd(train_ops, m, k, n, relu, bias):
    gen = torch.Generator().manual_seed(m * 1000 + k)
    x = torch.randn(m, k, generator=gen).requires_grad_(True)
    w = (torch.randn(n, k, generator=gen) / k ** 0.5).requires_grad_(True)
    b = torch.randn(n, generator=gen).requires_grad_(True) if bias else None
    g = torch.randn(m, n, generator=gen)
    y = train_ops._LinearFunction.apply(x, w, b, relu)
    assert y.shape == (m, n) and y.is_contiguous()
    (y * g).sum().backward()
    x64, w64 = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    b64 = b.detach().double().requires_grad_(True) if bias else None
    y64 = torch.nn.functional.linear(x64, w64, b64)
    if relu:
        y64 = torch.relu(y64)
    (y64 * g.double()).sum().backward()
    assert _rel(y, y64.detach()) < 1e-6
    assert _rel(x.grad, x64.grad) < 1e-5 and _rel(w.grad, w64.grad) < 1e-5
    if bias:
        assert _rel(b.grad, b64.grad) < 1e-5


def test_only_the_needed_gradients_are_formed(train_ops):
    """the 141-wide MLP input carries no gradient: layer 1 must not launch the dx GEMM"""
    x = torch.randn(64, 141)
    w = torch.randn(512, 141, requires_grad=True)
    b = torch.zeros(512, requires_grad=True)
    y = train_ops._LinearFunction.apply(x, w, b, True)
    n_fwd = len(train_ops._calls)
    y.sum().backward()
    assert n_fwd == 1 and len(train_ops._calls) == 2  # forward + dW only
    assert train_ops._calls[1] == (512, 64, 256)      # dy^T [512, 64 rays] x x^T [141 -> 256 rows, 64]


def test_dw_reduction_over_ray_chunks(train_ops):
    """a^T b over a ray axis longer than one chunk, ragged last chunk (not a multiple of 16)"""
    gen = torch.Generator().manual_seed(3)
    a, b = torch.randn(1003, 96, generator=gen), torch.randn(1003, 141, generator=gen)
    out = train_ops.gemm_tn(a, b, chunk=256)
    assert _rel(out, a.double().t() @ b.double()) < 1e-6
    assert [c[1] for c in train_ops._calls] == [256, 256, 256, 240]  # 235 rays -> 240
    assert train_ops.gemm_tn(a[:0], b[:0]).abs().sum() == 0


def test_ray_keys_and_queries_gradients_vs_torch_modules(train_ops, sx, synthetic):
    """the composed training forward (PE -> mlp -> [h, x] -> mlp2 -> k_proj; q_proj) and all twelve parameter
    gradients equal those of the module's torch-op route (identification.py: the SIXDGS_TRAIN_MLP=torch branch)"""
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone(), score_impl="simt_fp32")
    idm.load_state_dict(synthetic.synth_id_weights(seed=3, q_gain=8.0), strict=False)
    idm = idm.double()
    gen = torch.Generator().manual_seed(0)
    n = 333
    ori, dirs, rgb = (torch.randn(n, 3, generator=gen).double(),
                      torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1).double(),
                      torch.rand(n, 3, generator=gen).double())
    tok = torch.randn(201, 398, generator=gen).double()
    rp, att = idm.ray_preprocessor, idm.attention

    def loss_of(q, k):
        a = torch.softmax(q.double() @ k.double().t() / 384 ** 0.5, -1)
        return (a.sum(0) ** 2).sum()

    x = train_ops.ray_mlp_input(rp, ori, dirs, rgb)
    assert x.shape == (n, 141)
    k_ref = att.k_proj(rp.mlp2(torch.cat((rp.mlp(x), x), -1)))
    q_ref = att.q_proj(tok)
    loss_of(q_ref, k_ref).backward()
    ref = {name: p.grad.clone() for name, p in idm.named_parameters() if p.grad is not None}
    idm.zero_grad()
    idm = idm.float()
    k = train_ops.ray_keys(rp, att, ori.float(), dirs.float(), rgb.float())
    q = train_ops.image_queries(att, tok.float())
    assert _rel(k, k_ref.detach()) < 1e-5 and _rel(q, q_ref.detach()) < 1e-5
    loss_of(q, k).backward()
    got = {name: p.grad for name, p in idm.named_parameters() if p.grad is not None}
    assert set(got) == set(ref) and len(ref) == 12
    floor = 1e-6 * max(v.abs().max().item() for v in ref.values())
    for name, g64 in ref.items():
        assert (got[name].double() - g64).abs().max().item() <= 2e-4 * g64.abs().max().item() + floor, name
    # every launch stayed inside the kernel's exercised envelope (the stand-in asserts it) and none had zero rows
    assert all(m > 0 for m, _, _ in train_ops._calls)


def test_no_rays(train_ops):
    w = torch.randn(384, 384, requires_grad=True)
    y = train_ops._LinearFunction.apply(torch.zeros(0, 384), w, None, False)
    assert y.shape == (0, 384)
    y.sum().backward()
    assert w.grad.abs().sum() == 0 and not train_ops._calls


def test_gradients_come_back_in_the_inputs_dtype(train_ops):
    """the kernel computes in fp32; a double-precision parameter (e.g. a module under .double() in a test) still gets a
    gradient of its own dtype, as autograd requires"""
    x = torch.randn(9, 32, dtype=torch.float64, requires_grad=True)
    w = torch.randn(16, 32, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(16, dtype=torch.float64, requires_grad=True)
    y = train_ops._LinearFunction.apply(x, w, b, True)
    assert y.dtype == torch.float32
    y.sum().backward()
    assert x.grad.dtype == w.grad.dtype == b.grad.dtype == torch.float64
