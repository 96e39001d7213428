"""Distance-based target scores and score loss (6dgs_b200/losses.py) against the UNMODIFIED reference
(tests/golden/loss.npz, written by oracle/gen_golden.py --only-loss from pose_estimation/distance_based_loss.py), and the
"oracle rays" branch of test_pose_estimation (test.py:110-142) against the reference's own run of it.  CPU only: the
losses are torch ops; the evaluation loop is driven with an oracle-backed stand-in for the kernels (tests may use the
oracle, the package never does)."""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_golden


@pytest.fixture(scope="module")
def losses():
    return importlib.import_module("6dgs_b200.losses")


def test_target_scores_match_reference(losses):
    g, r = load_golden("loss.npz"), load_golden("rays_small.npz")
    for i in range(3):
        pose, K = g[f"pose{i}"], g[f"K{i}"]
        for shape in ((800, 800), (480, 640)):
            idx, inside, tgt, tgt_d = losses.best_one_to_one_rays_selector(K, pose, shape, r["dirs"], r["ori"], backbone_wh=(16, 16))
            assert idx is None and inside.dtype == torch.bool
            # a projection within 1 ulp of a patch-grid edge may land on the other side
            assert (inside != g[f"inside{i}_{shape[0]}"]).sum() <= 1
        torch.testing.assert_close(tgt, g[f"target_raw{i}"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(tgt_d, g[f"target_dist{i}"], rtol=1e-5, atol=1e-6)
        assert torch.equal(tgt == 0, g[f"target_raw{i}"] == 0)  # rays behind the camera: exactly zero
        loss, target = losses.DistanceBasedScoreLoss()(g["pred"], pose, K, r["ori"], r["dirs"], 256, (16, 16), model_up=None)
        torch.testing.assert_close(target, g[f"target{i}"], rtol=1e-5, atol=1e-6)
        assert abs(float(loss) - float(g[f"loss{i}"])) < 1e-5 * float(g[f"loss{i}"])
        assert abs(float(target.sum()) - 256.0) < 1e-2  # same total mass as the predicted scores of one image


def test_loss_gradient_reaches_the_scores_only(losses):
    g, r = load_golden("loss.npz"), load_golden("rays_small.npz")
    pred = g["pred"].clone().requires_grad_(True)
    ori = r["ori"].clone().requires_grad_(True)
    loss, target = losses.DistanceBasedScoreLoss()(pred, g["pose0"], g["K0"], ori, r["dirs"], 256, (16, 16))
    loss.backward()
    assert not target.requires_grad and ori.grad is None  # the targets are constants, as upstream (no_grad)
    torch.testing.assert_close(pred.grad, 2.0 * (pred.detach() - target) / pred.numel(), rtol=1e-6, atol=1e-9)


def test_constructor_contract(losses):
    with pytest.raises(AssertionError):
        losses.DistanceBasedScoreLoss(reweight_method="bogus")
    with pytest.raises(AssertionError):
        losses.DistanceBasedScoreLoss(lds=True)  # LDS needs a re-weighting method
    losses.DistanceBasedScoreLoss(reweight_method="sqrt_inv", lds=True)


class _OracleIdModule:
    """stand-in with the package module's front end (torch, CPU) and the oracle's scores: what evaluate.py needs"""

    def __init__(self, sx, oracle, synthetic, w):
        self.idm = sx.IdentificationModule("dino", backbone=synthetic.SyntheticBackbone())
        self.idm.load_state_dict(w, strict=False)
        self.backbone_wrapper = self.idm.backbone_wrapper
        self.oracle, self.w, self._fea = oracle, w, None

    def eval(self):
        return self

    @torch.no_grad()
    def test_image(self, img, mask, ori, dirs, rgb, rays_to_output=100):
        tok_pe, _, grid = self.idm.backbone_wrapper(img, mask)
        up = self.idm._camera_up(grid)
        if self._fea is None:
            self._fea = self.oracle.ray_features(ori, dirs, rgb, self.w)
        scores, _ = self.oracle.attention_scores(tok_pe, self._fea, self.w, return_map=False)
        top = torch.topk(scores, rays_to_output)
        return top.indices, top.values, scores, up, None  # no attention map: the loop re-derives n_img


def test_oracle_rays_evaluation_matches_reference(sx, oracle, synthetic, losses, monkeypatch):
    """test_pose_estimation(..., loss_fn=DistanceBasedScoreLoss()): per frame the score loss, the recall of the
    predicted top-100 and the pose solved from the top-100 TARGET scores, as the unmodified reference reports them"""
    from collections import namedtuple
    evaluate = importlib.import_module("6dgs_b200.evaluate")
    g, r, p = load_golden("loss.npz"), load_golden("rays_small.npz"), load_golden("pose.npz")

    def pose_tail(ori, dirs, idx, weights, up):
        c2w, info = oracle.pose_tail(idx, weights, ori, dirs, up)
        aux = torch.zeros(8)
        aux[6] = float(info["weights"].numel())  # rays kept by the dedup, as the kernel reports it
        return c2w, aux

    monkeypatch.setattr(evaluate.ops, "pose_tail", pose_tail)
    Cam = namedtuple("Cam", "uid R T FovY FovX image image_path image_name width height")
    cams = [Cam(i, p["R"][i].numpy(), p["T"][i].numpy(), np.float32(0.9), np.float32(0.9), p[f"img{i}"].numpy(), "", str(i), 64, 64)
            for i in range(3)]
    idm = _OracleIdModule(sx, oracle, synthetic, synthetic.synth_id_weights(seed=3))
    res, t_err, a_err, avg_loss, avg_recall = evaluate.test_pose_estimation(cams, idm, r["ori"], r["dirs"], r["rgb"],
                                                                            torch.tensor([0.0, 0.0, 1.0]),
                                                                            loss_fn=losses.DistanceBasedScoreLoss())
    for i in range(3):
        torch.testing.assert_close(torch.tensor(res[i]["pred_c2w"]), g["pred_c2w"][i], rtol=1e-4, atol=1e-4)
        assert abs(res[i]["scores_loss"] - float(g["scores_loss"][i])) < 1e-5 * float(g["scores_loss"][i]) + 1e-8
        assert abs(res[i]["recall"] - float(g["recall"][i])) < 1e-6
    assert abs(t_err - float(g["avg_t_err"])) < 1e-4 and abs(a_err - float(g["avg_ang_err"])) < 1e-2
    assert abs(avg_loss - float(g["avg_loss"])) < 1e-6 and abs(avg_recall - float(g["avg_recall"])) < 1e-6
