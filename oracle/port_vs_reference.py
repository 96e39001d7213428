"""Test infrastructure (runs in the build container only: it imports the UNMODIFIED reference from /root/reference).

Times the reference's own ``IdentificationModule.test_image`` beside the oracle port that bench.py's ``cpu_baseline`` /
``--impl reference`` legs execute, on the same 29,000 rays (the reference's largest CPU-runnable case, BASELINE configs[0])
with the same weights and threads, so that the "kind": "port" baseline can be read against the real thing.
Result of the last run: profiles/cpu_port_vs_reference_r2.md.   python oracle/port_vs_reference.py
"""
import sys, time, importlib, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shims; ref_shims.install()
import sixdgs_oracle as oracle
syn = importlib.import_module("6dgs_b200.synthetic")
from pose_estimation.identification_module import IdentificationModule
torch.set_num_threads(8)
g = np.load(os.path.join(ROOT, "tests/golden/rays_small.npz"))
ori, dirs, rgb = (torch.from_numpy(g[k]) for k in ("ori", "dirs", "rgb"))
rep = 29000 // ori.shape[0] + 1
ori, dirs, rgb = ori.repeat(rep, 1)[:29000], dirs.repeat(rep, 1)[:29000], rgb.repeat(rep, 1)[:29000]
w = syn.synth_id_weights(seed=3)
idm = IdentificationModule("dino").eval(); idm.load_state_dict(w, strict=False)
img = syn.synth_image(64, 64, seed=4); mask = torch.ones(64, 64, dtype=torch.bool)
with torch.no_grad():
    tok_pe, _, _ = idm.backbone_wrapper(img, mask)
def ref_query():
    with torch.no_grad():
        idx, vals, scores, up, _ = idm.test_image(img, mask, ori, dirs, rgb, rays_to_output=100)
    return idx
def port_query():
    fea = oracle.ray_features(ori, dirs, rgb, w)
    s, _ = oracle.attention_scores(tok_pe, fea, w, return_map=False)
    top = torch.topk(s, 100)
    return oracle.pose_tail(top.indices, top.values, ori, dirs, torch.tensor([0., 0., 1.]))
for name, fn in (("reference test_image (unmodified, incl. its backbone wrapper + up head)", ref_query), ("oracle port (MLP + attention + top-100 + pose tail)", port_query)):
    fn(); ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(f"{name}: {min(ts)*1e3:.1f} ms per query on 29000 rays = {min(ts)/29000*1e6:.2f} us/ray (best of 5, 8 threads)")
def t(fn, n=5):
    fn(); ts=[]
    for _ in range(n):
        t0=time.perf_counter(); fn(); ts.append(time.perf_counter()-t0)
    return min(ts)*1e3
with torch.no_grad():
    print("ref ray_preprocessor", t(lambda: idm.ray_preprocessor(ori, dirs, rgb)))
    fea = idm.ray_preprocessor(ori, dirs, rgb)
    print("ref attention", t(lambda: idm.attention(tok_pe, fea).sum(0)))
    print("ref backbone wrapper", t(lambda: idm.backbone_wrapper(img, mask)))
print("port ray_features", t(lambda: oracle.ray_features(ori, dirs, rgb, w)))
print("port attention_scores", t(lambda: oracle.attention_scores(tok_pe, fea, w, return_map=False)))
print("port attention_scores(map)", t(lambda: oracle.attention_scores(tok_pe, fea, w, return_map=True)))
