"""Host logic of the implicit key cache (6dgs_b200/identification.py: _cache_for).  The reference recomputes the ray MLP
for every query (identification_module.py:79-80); here K is cached per (ray set, weights) and must be rebuilt exactly when
either changes -- the training loop steps the optimiser between its evaluations (pose_estimation/train.py:188,222-282)
and regenerates the rays every 10 iterations (train.py:70-71).  The build itself is replaced by a counter (no GPU)."""
import torch


def _module(sx, synthetic, monkeypatch):
    idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone(), score_impl="tc_f16x2")
    built = []

    def fake_build(ori, dirs, rgb):
        built.append((ori, dirs, rgb))
        return sx.RayKeyCache(torch.zeros(ori.shape[0], 768, dtype=torch.float16), ori.shape[0], ())

    monkeypatch.setattr(idm, "build_key_cache", fake_build)
    return idm, built


def test_cache_is_reused_for_the_same_rays_and_weights(sx, synthetic, monkeypatch):
    idm, built = _module(sx, synthetic, monkeypatch)
    ori, dirs, rgb = torch.randn(10, 3), torch.randn(10, 3), torch.rand(10, 3)
    a = idm._cache_for(ori, dirs, rgb)
    b = idm._cache_for(ori, dirs, rgb)
    assert a is b and len(built) == 1
    assert a.rays[0] is ori  # the cache keeps the tensors alive: their address cannot be recycled under it


def test_optimizer_step_invalidates_the_cache(sx, synthetic, monkeypatch):
    idm, built = _module(sx, synthetic, monkeypatch)
    ori, dirs, rgb = torch.randn(10, 3), torch.randn(10, 3), torch.rand(10, 3)
    idm._cache_for(ori, dirs, rgb)
    opt = torch.optim.SGD(list(idm.ray_preprocessor.parameters()) + list(idm.attention.parameters()), lr=0.1)
    for p in idm.ray_preprocessor.parameters():
        p.grad = torch.ones_like(p)
    opt.step()  # in-place update: data_ptr unchanged, version counter bumped
    idm._cache_for(ori, dirs, rgb)
    assert len(built) == 2
    # a change outside the hot path (camera-up head) does not touch the keys
    with torch.no_grad():
        for p in idm.camera_direction_prediction_network.parameters():
            p.add_(1.0)
    idm._cache_for(ori, dirs, rgb)
    assert len(built) == 2
    # loading a checkpoint rewrites the parameters in place as well
    idm.load_state_dict(synthetic.synth_id_weights(seed=4), strict=False)
    idm._cache_for(ori, dirs, rgb)
    assert len(built) == 3


def test_new_or_modified_rays_invalidate_the_cache(sx, synthetic, monkeypatch):
    idm, built = _module(sx, synthetic, monkeypatch)
    ori, dirs, rgb = torch.randn(10, 3), torch.randn(10, 3), torch.rand(10, 3)
    idm._cache_for(ori, dirs, rgb)
    rgb.mul_(0.5)                                   # edited in place through torch: version counter
    idm._cache_for(ori, dirs, rgb)
    assert len(built) == 2
    idm._cache_for(ori.clone(), dirs, rgb)          # regenerated rays (train.py:70-71): new tensor objects, same count
    assert len(built) == 3
    idm.invalidate_key_cache()                      # raw-pointer writes are invisible to torch: explicit invalidation
    idm._cache_for(ori, dirs, rgb)
    assert len(built) == 5 - 1
    idm.score_impl = "simt_fp32"                    # another key format
    idm._cache_for(ori, dirs, rgb)
    assert len(built) == 5
