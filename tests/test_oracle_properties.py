"""Property tests of the CPU oracle (hypothesis): independent of the fixtures, they check that the restated
algorithm has the mathematical properties it must have whatever the inputs."""
import math

import torch
from hypothesis import given, settings, strategies as st


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 10_000))
def test_sym_eig_reconstructs_matrix(oracle, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(16, 12, 3, generator=g) * (0.1 + torch.rand(16, 1, 3, generator=g) * 3)
    A = X.mT @ X
    vals, vecs = oracle.sym_eig_3x3(A)
    assert (vals[:, 1:] >= vals[:, :-1] - 1e-4 * vals.abs().max()).all()
    rec = vecs @ torch.diag_embed(vals) @ vecs.mT
    scale = A.abs().amax(dim=(1, 2), keepdim=True)
    assert ((rec - A).abs() / scale).max() < 2e-3
    assert ((vecs.mT @ vecs - torch.eye(3)).abs()).max() < 2e-3


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 10_000), st.integers(3, 200))
def test_line_intersection_recovers_common_point(oracle, seed, n):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(3, generator=g) * 3
    o = torch.randn(n, 3, generator=g) * 2
    d = torch.nn.functional.normalize(c[None] - o, dim=-1)
    w = torch.rand(n, generator=g) + 0.1
    for weights in (None, w):
        est = oracle.line_intersection(o, d, weights)
        assert (est - c).abs().max() < 2e-3 * (1 + c.abs().max())
    assert oracle.exclude_negatives(c, o, d).all()


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 10_000))
def test_quadricell_points_lie_on_the_ellipsoid_and_counts_are_consistent(oracle, seed):
    g = torch.Generator().manual_seed(seed)
    abc = torch.rand(12, 3, generator=g) * 0.05 + 0.005
    pts, eid = oracle.quadricell_centers(abc[:, 0], abc[:, 1], abc[:, 2], 50)
    a, b, c = abc[eid, 0], abc[eid, 1], abc[eid, 2]
    r = (pts[:, 0] / b) ** 2 + (pts[:, 1] / c) ** 2 + (pts[:, 2] / a) ** 2  # a-axis lives in z
    assert (r - 1).abs().max() < 1e-3
    assert (eid[1:] >= eid[:-1]).all()
    counts = torch.bincount(eid, minlength=12)
    assert (counts >= 20).all() and (counts <= 400).all()  # ~50 targeted cells


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 10_000), st.integers(1, 256))
def test_scores_sum_to_token_count_and_chunking_is_exact(oracle, synthetic, seed, n_img):
    g = torch.Generator().manual_seed(seed)
    w = synthetic.synth_id_weights(seed=seed % 7)
    tok = torch.randn(n_img, 398, generator=g)
    fea = torch.randn(700, 384, generator=g)
    score, A = oracle.attention_scores(tok, fea, w)
    assert abs(float(score.sum()) - n_img) < 1e-3 * n_img + 1e-3
    assert (A.sum(1) - 1).abs().max() < 1e-4
    s2, m, z = oracle.attention_scores_chunked(tok, lambda lo, hi: fea[lo:hi], 700, w, chunk=97)
    assert (s2 - score).abs().max() <= 1e-5 * score.abs().max() + 1e-9


def test_make_rotation_is_orthonormal_and_aligned(oracle):
    g = torch.Generator().manual_seed(3)
    for _ in range(20):
        d = torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
        up = torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
        R = oracle.make_rotation_mat(d, up)
        assert (R @ R.T - torch.eye(3)).abs().max() < 1e-5
        assert torch.allclose(R[2], d) and abs(float(torch.linalg.det(R)) - 1.0) < 1e-5
        assert abs(float(R[0] @ up)) < 1e-5  # x axis is orthogonal to up
