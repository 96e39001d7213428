#!/usr/bin/env python
"""Time the three key-cache builds on the same rays (second call of each, CUDA events):
fp32 FMA + split (exact, round-2 start), split-fp16 tensor cores (exact, default), TF32 tensor cores (bf16 cache)."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sx = importlib.import_module("6dgs_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
ori = (torch.randn(n, 3, generator=g) * 3).to(dev)
dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
rgb = torch.rand(n, 3, generator=g).to(dev)
for impl, feat in (("tc_f16x2", "auto"), ("tc_f16x2", "simt"), ("tc_bf16", "auto")):
    idm = sx.IdentificationModule("dino", backbone=sx.synthetic.SyntheticBackbone(), score_impl=impl)
    idm.load_state_dict(sx.synthetic.synth_id_weights(seed=3), strict=False)
    idm = idm.to(dev).eval().requires_grad_(False)
    idm.features_impl = feat
    for rep in range(2):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        c = idm.build_key_cache(ori, dirs, rgb)
        b.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        del c
    ms = a.elapsed_time(b)
    terms = 3 if (impl == "tc_f16x2" and feat == "auto") else 1
    print(f"{impl:9s} features={feat:5s} {n} rays: {ms:8.1f} ms device ({wall * 1e3:8.1f} ms wall) = {n / ms / 1e3:7.1f} M rays/s, "
          f"{2.03e6 * terms * n / ms / 1e9:7.1f} TFLOP/s (MMA terms counted)", flush=True)
