"""Opt-in checks for code that is built but not on any default path yet (run on a B200 with
SIXDGS_EXPERIMENTAL=1 python -m pytest tests/test_experimental.py).  They are skipped in the driver's CPU and GPU
test runs on purpose: an experimental kernel must never be able to break the validated suite."""
import os

import pytest
import torch

run = os.environ.get("SIXDGS_EXPERIMENTAL") == "1" and torch.cuda.is_available()
pytestmark = pytest.mark.skipif(not run, reason="set SIXDGS_EXPERIMENTAL=1 on a GPU box")


def test_cta_pair_tf32_gemm_matches_the_one_cta_kernel(sx, synthetic):
    """features_tc2.cu (CTA pairs, full-width tiles) vs features_tc.cu (validated) vs the fp32 build"""
    dev = "cuda"
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone())
    idm.load_state_dict(synthetic.synth_id_weights(seed=3), strict=False)
    pw = idm.to(dev).packed_weights()
    gen = torch.Generator().manual_seed(3)
    for n in (1, 255, 256, 257, 200_000):
        ori = (torch.randn(n, 3, generator=gen) * 3).to(dev)
        dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1).to(dev)
        rgb = torch.rand(n, 3, generator=gen).to(dev)
        k1, f1 = sx.ops.ray_features(ori, dirs, rgb, pw, k_dtype=sx._lib.F32, want_features=True, impl=sx.ops.FEATURES_TC)
        k2, f2 = sx.ops.ray_features(ori, dirs, rgb, pw, k_dtype=sx._lib.F32, want_features=True, impl=sx.ops.FEATURES_TC2)
        k0, _ = sx.ops.ray_features(ori, dirs, rgb, pw, k_dtype=sx._lib.F32, impl=sx.ops.FEATURES_SIMT)
        torch.cuda.synchronize()
        scale = k0.abs().max().item()
        assert (k2 - k1).abs().max().item() <= 2e-4 * scale, n   # same TF32 products, different accumulation order
        assert (k2 - k0).abs().max().item() <= 1.5e-3 * scale, n
        kb, _ = sx.ops.ray_features(ori, dirs, rgb, pw, k_dtype=sx._lib.BF16, impl=sx.ops.FEATURES_TC2)
        assert (kb.float() - k0).abs().max().item() <= 4e-3 * scale


def test_staged_tma_store_epilogue_matches_the_direct_store_kernel(sx, synthetic):
    """features_tc.cu: linear_tc_staged_kernel (smem-staged TMA tensor stores) must be bit-identical to
    linear_tc_kernel (same MMAs, same epilogue math; only the way the tile reaches HBM differs)"""
    dev = "cuda"
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone())
    idm.load_state_dict(synthetic.synth_id_weights(seed=3), strict=False)
    pw = idm.to(dev).packed_weights()
    gen = torch.Generator().manual_seed(4)
    for n in (1, 127, 128, 129, 200_000):
        ori = (torch.randn(n, 3, generator=gen) * 3).to(dev)
        dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1).to(dev)
        rgb = torch.rand(n, 3, generator=gen).to(dev)
        for kd in (sx._lib.F32, sx._lib.BF16):
            k1, f1 = sx.ops.ray_features(ori, dirs, rgb, pw, k_dtype=kd, want_features=True, impl=sx.ops.FEATURES_TC)
            k3, f3 = sx.ops.ray_features(ori, dirs, rgb, pw, k_dtype=kd, want_features=True, impl=sx.ops.FEATURES_TC_STAGED)
            torch.cuda.synchronize()
            assert torch.equal(k1, k3) and torch.equal(f1, f3), (n, kd)
