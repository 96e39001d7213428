#!/usr/bin/env python
"""Sustained rates of the ray-score kernels on one key cache, each variant run back to back for `--seconds` so the
numbers are taken in the power-capped steady state:
  bf16  per-query   B x (pass 1 + pass 2) of score_tc.cu            -- 2B sweeps over 768-B keys, 1 MMA term
  bf16  multi-query 1 x (pass 1 + pass 2) of score_tc_mq.cu<bf16>   -- 2 sweeps
  f16x2 multi-query 1 x (pass 1 + pass 2) of score_tc_mq.cu<f16x2>  -- 2 sweeps over 1536-B keys, 3 MMA terms (exact mode)
Prints ms per query, the MMA rate (terms counted), the HBM rate of the keys actually read, and clocks / power.
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
ap.add_argument("--batches", default="8,4,1")
ap.add_argument("--seconds", type=float, default=3.0)
ap.add_argument("--variants", default="bf16_pq,bf16_mq,f16x2_mq,f16f8_mq")
args = ap.parse_args()

dev = torch.device("cuda:0")
n = args.rays
kf = torch.randn(n, 384, device=dev) * 0.5
K16 = kf.to(torch.bfloat16)
KX = sx.ops.split_keys(kf)
K8 = sx.ops.keys_to_f16f8(sx.ops.split_keys(kf))
del kf
scores1 = torch.empty(n, device=dev)


def per_query(K, q, sb):
    for i in range(q.shape[0]):
        pm, pz = sx.ops.score_pass1(K, q[i], sx.ops.SCORE_TC)
        m, z = sx.ops.score_merge(pm, pz, 256)
        sx.ops.score_pass2(K, q[i], m, z, sx.ops.SCORE_TC, out=scores1)


def multi_query(K, q, sb):
    B = q.shape[0]
    pm, pz = sx.ops.score_pass1_batch(K, q)
    parts = pm.shape[0] // B
    mz = [sx.ops.score_merge(pm, pz, 256, rows=parts, first_row=i * parts) for i in range(B)]
    sx.ops.score_pass2_batch(K, q, torch.stack([x[0] for x in mz]), torch.stack([x[1] for x in mz]), out=sb)


def sm_clock():
    try:
        out = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                       "--format=csv,noheader,nounits", "-i", "0"], text=True)
        return out.strip()
    except OSError:
        return "n/a"


VARIANTS = {"bf16_pq": (per_query, K16, 768, 1, False), "bf16_mq": (multi_query, K16, 768, 1, True),
            "f16x2_mq": (multi_query, KX, 1536, 3, True),
            # e4m3 cross terms: 3 MMA terms' worth of FLOPs issued as 2 term-units of tensor time
            "f16f8_mq": (multi_query, K8, 1536, 3, True)}
for B in [int(b) for b in args.batches.split(",")]:
    q = torch.randn(B, 256, 384, device=dev)
    sb = torch.empty(B, n, device=dev)
    for name in args.variants.split(","):
        fn, K, row, terms, shared = VARIANTS[name]
        fn(K, q, sb)
        torch.cuda.synchronize()
        t_end = time.perf_counter() + args.seconds
        times, clk = [], ""
        while time.perf_counter() < t_end:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(K, q, sb)
            b.record()
            b.synchronize()
            times.append(a.elapsed_time(b))
            if len(times) % 4 == 0:
                clk = sm_clock()
        tail = times[len(times) // 2:]  # steady state: second half of the run
        ms_batch = sum(tail) / len(tail)
        sweeps = 2 if shared else 2 * B
        print(f"B={B} {name:9s} {ms_batch / B:7.3f} ms/query ({len(times)} batches)  MMA {2 * 2 * 256 * 384 * terms * n * B / ms_batch / 1e9:7.1f} "
              f"TFLOP/s  keys from HBM {sweeps * n * row / ms_batch / 1e6:7.1f} GB/s  sm MHz, W, power cap: {clk}", flush=True)
        time.sleep(2.0)
