#!/usr/bin/env python
"""Sustained comparison of the two ways to score a batch of queries against one bf16 key cache:
  per-query   8 x (pass 1 + pass 2) of score_tc.cu      -- the default path, 16 sweeps over the keys
  multi-query 1 x (pass 1 + pass 2) of score_tc_mq.cu   -- EXPERIMENTAL, 2 sweeps
Each variant runs back to back for `--seconds` so the numbers are taken in the power-capped steady state
(DESIGN.md §6.1); prints ms per query, the effective key bandwidth and tensor rate, and the SM clock.
Run it under a shell `timeout` on the GPU box: a tcgen05 kernel with a barrier mistake hangs rather than fails."""
import argparse
import importlib
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sx = importlib.import_module("6dgs_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=12_000_000)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seconds", type=float, default=4.0)
args = ap.parse_args()

dev = torch.device("cuda:0")
B, n = args.batch, args.rays
K = (torch.randn(n, 384, device=dev) * 0.5).to(torch.bfloat16)
q = torch.randn(B, 256, 384, device=dev)
scores1 = torch.empty(n, device=dev)
scores_b = torch.empty(B, n, device=dev)


def per_query():
    for i in range(B):
        pm, pz = sx.ops.score_pass1(K, q[i], sx.ops.SCORE_TC)
        m, z = sx.ops.score_merge(pm, pz, 256)
        sx.ops.score_pass2(K, q[i], m, z, sx.ops.SCORE_TC, out=scores1)


def multi_query():
    pm, pz = sx.ops.score_pass1_batch(K, q)
    parts = pm.shape[0] // B
    mz = [sx.ops.score_merge(pm, pz, 256, rows=parts, first_row=i * parts) for i in range(B)]
    sx.ops.score_pass2_batch(K, q, torch.stack([x[0] for x in mz]), torch.stack([x[1] for x in mz]), out=scores_b)


def sm_clock():
    try:
        out = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                       "--format=csv,noheader,nounits", "-i", "0"], text=True)
        return out.strip()
    except OSError:
        return "n/a"


for name, fn in (("per-query", per_query), ("multi-query", multi_query), ("per-query", per_query), ("multi-query", multi_query)):
    fn()
    torch.cuda.synchronize()
    t_end = time.perf_counter() + args.seconds
    times, clk = [], ""
    while time.perf_counter() < t_end:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b))
        if len(times) % 8 == 0:
            clk = sm_clock()
    tail = times[len(times) // 2:]  # steady state: second half of the run
    ms = sum(tail) / len(tail) / B
    print(f"{name:12s} {ms:7.3f} ms/query  ({len(times)} batches; keys {2 * n * 768 / ms / 1e6:7.1f} GB/s per query-equivalent, "
          f"{2 * 2 * 256 * 384 * n / ms / 1e9:6.1f} TFLOP/s)  sm MHz, W, power cap: {clk}")
    time.sleep(3.0)
