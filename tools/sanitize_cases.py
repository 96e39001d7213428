#!/usr/bin/env python
"""Small-size exercise of the round-2 kernels for compute-sanitizer (tools/sanitize.sh): fast enough under memcheck.
  python tools/sanitize_cases.py [simt]     simt = only the kernels without tcgen05 / TMA (racecheck models neither)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sx = importlib.import_module("6dgs_b200")
simt_only = len(sys.argv) > 1 and sys.argv[1] == "simt"
dev = "cuda"
g = torch.Generator().manual_seed(0)

# ray generation (cells -> fill -> compact) and kNN normals
scene = sx.GaussianScene.from_dict(sx.synthetic.synth_scene(60, seed=3), device=dev)
ori, dirs, rgb = sx.generate_all_possible_rays(scene, max_ellipsoids=None)
print("raygen", tuple(ori.shape))
# top-k with more ties than the tie table holds
x = torch.zeros(20000, device=dev)
x[-3:] = 1.0
v, i = sx.ops.topk(x, 100)
assert i[:3].tolist() == [19997, 19998, 19999] and i[3:].tolist() == list(range(97))
# score backward (SIMT)
n = 500
k = (torch.randn(n, 384, generator=g) * 0.7).to(dev)
q = (torch.randn(201, 384, generator=g) * 2).to(dev)
pm, pz = sx.ops.score_pass1(k, q, sx.ops.SCORE_SIMT)
m, z = sx.ops.score_merge(pm, pz, 201)
dq, dk = sx.ops.score_backward(k, q, m, z, torch.randn(n, generator=g).to(dev))
print("score_backward", tuple(dq.shape), tuple(dk.shape))
# weighted-LS solve from a hand-made system
sysm = torch.tensor([[2.0, 0, 0, 2.0, 0, 2.0, 1.0, 2.0, 3.0, 0, 0, 1.0, 3.0]], dtype=torch.float64, device=dev)
c2w, aux = sx.ops.ls_solve(sysm, 1.0, torch.tensor([[0.0, 1.0, 0.0]], device=dev))
assert torch.allclose(c2w[0, :3, 3].cpu(), torch.tensor([0.5, 1.0, 1.5]))
if not simt_only:
    idm = sx.IdentificationModule("dino", backbone=sx.synthetic.SyntheticBackbone(), score_impl="tc_f16x2")
    idm.load_state_dict(sx.synthetic.synth_id_weights(seed=3), strict=False)
    idm = idm.to(dev).eval().requires_grad_(False)
    cache = idm.build_key_cache(ori, dirs, rgb)                       # features_x2.cu
    qb = (torch.randn(3, 256, 384, generator=g) * 3).to(dev)
    pmb, pzb = sx.ops.score_pass1_batch(cache.keys, qb)               # score_tc_mq.cu, f16x2
    parts = pmb.shape[0] // 3
    mz = [sx.ops.score_merge(pmb, pzb, 256, rows=parts, first_row=b * parts) for b in range(3)]
    sc, ls = sx.ops.score_pass2_batch(cache.keys, qb, torch.stack([a for a, _ in mz]), torch.stack([b for _, b in mz]),
                                      ls_rays=(ori, dirs))
    kb = (cache.keys[:, :384].float() / 16).to(torch.bfloat16).contiguous()
    pm1, pz1 = sx.ops.score_pass1_batch(kb, qb)                        # score_tc_mq.cu, bf16
    torch.cuda.synchronize()
    print("tensor-core kernels", tuple(sc.shape), tuple(ls.shape), float(sc.sum()))
torch.cuda.synchronize()
print("sanitize cases done")
