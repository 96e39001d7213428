#!/usr/bin/env python
"""Sustained-state probe for the ray-score kernel: run pass 1 / pass 2 back to back for ~2 s each while
sampling nvidia-smi (SM clock, memory clock, power, throttle reasons) every 20 ms, and print per-launch
times next to the clocks.  Explains the burst-vs-sustained gap reported by bench.py."""
import importlib
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sx = importlib.import_module("6dgs_b200")

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
dev = torch.device("cuda:0")
K = (torch.randn(n_rays, 384, device=dev) * 0.5).to(torch.bfloat16)
q = torch.randn(256, 384, device=dev)
scores = torch.empty(n_rays, device=dev)
pm, pz = sx.ops.score_pass1(K, q, sx.ops.SCORE_TC)
m, z = sx.ops.score_merge(pm, pz, 256)
torch.cuda.synchronize()
time.sleep(2.0)  # cool down

for name, fn in (("pass1", lambda: sx.ops.score_pass1(K, q, sx.ops.SCORE_TC)),
                 ("pass2", lambda: sx.ops.score_pass2(K, q, m, z, sx.ops.SCORE_TC, out=scores))):
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,"
                          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu",
                          "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=f)
    time.sleep(0.1)
    evs = []
    n_it = 400
    for i in range(n_it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    time.sleep(0.05)
    p.terminate()
    p.wait()
    ts = [a.elapsed_time(b) for a, b in evs]
    rows = [r.strip() for r in open(f.name) if r.strip()]
    os.unlink(f.name)
    gb = n_rays * 768 / 1e9
    print(f"== {name}: {n_rays} rays, {gb:.2f} GB per launch, {n_it} launches")
    for lo in (0, 5, 20, 50, 100, 200, 300, 390):
        seg = ts[lo:lo + 10]
        t = sum(seg) / len(seg)
        print(f"  launches {lo:3d}-{lo + 9:3d}: {t:6.3f} ms  {gb / t * 1e3:7.1f} GB/s  {2 * 256 * 384 * n_rays / t / 1e9:7.1f} TFLOP/s")
    print("  nvidia-smi samples (sm MHz, mem MHz, W, power_cap, hw_slowdown, sw_thermal, temp):")
    step = max(1, len(rows) // 16)
    for r in rows[::step]:
        print("   ", r)
    time.sleep(2.0)
