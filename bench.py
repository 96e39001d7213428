#!/usr/bin/env python
"""bench.py -- pose queries/sec of the 6DGS hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one rank per GPU
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

Workload (config.workload): BASELINE.json configs[2] -- a synthetic 1M-Gaussian scene (every valid
ellipsoid casts rays, ~29 rays each), 1080x1920 query images, fp32 LS solve.  The default score mode is the one
that is PARITY-GREEN (tests/test_gpu_exact_tc.py: scores 1e-3 rel, pose 1e-4 against the reference, also on a
peaked softmax): `tc_f16x2`, the exact tensor-core mode -- keys and queries as fp16 hi+lo pairs, three MMA terms
per logit in one fp32 TMEM accumulator (csrc/score_tc_mq.cu).  `--score-impl tc_bf16` is the throughput mode
(one term, 768 B/ray, its own 3e-2 tolerance); its rate is reported next to the headline as `throughput_mode`.
A "step" is one batch of `--batch` (default 8) pose queries, each with its own image: images -> backbone tokens
-> q (once per batch) -> ONE sweep over the key cache per softmax pass for the whole batch (multi-query kernel)
-> per query merge, top-100, fused LS pose tail -> c2w.  `value` counts QUERIES per second; `latency_b1` in the
same line is the one-query-per-step figure.  Scene preparation (ray generation + key cache) is per scene, not per
query, and is reported separately.

  value     queries/s, image already resident in HBM, whole query captured in one CUDA graph
  e2e       queries/s through ShardedPoseEstimator.query_batch() from pinned uint8 HOST images
            (H2D + /255 + mask inside the timed region) to the 4x4 poses back on the host
  roofline  the ray-score kernel (pass 1 / pass 2 of the batched kernel, CUDA events around each launch) against BOTH
            roofs -- HBM (algorithmic bytes / measured copy bandwidth) and tensor pipe (MMA FLOPs / measured cuBLAS
            bf16 rate) -- with the binding one (the larger time bound) named in `bound`
  cpu_baseline  the oracle port (torch CPU, host threads stated) on a bounded sample of the same workload
            (>= 64 reference-sized 1000-ellipsoid chunks), extrapolated linearly in the ray count (`sample` says so)

N > 1 (torchrun): the selected ellipsoids -- hence rays and key cache -- are sharded in contiguous
blocks over the ranks (strong scaling: the scene is fixed); per query two tiny NCCL all-gathers
(softmax statistics [2,256]; top-100 candidates) couple the shards.
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gaussians", type=int, default=None)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--score-impl", default="tc_f16x2", choices=["tc_f16x2", "tc_f16f8", "tc_bf16", "simt_bf16", "simt_fp32"],
                    help="tc_f16x2 = exact tensor-core mode (parity-green, default); tc_bf16 = throughput mode")
    ap.add_argument("--backbone", default="vits14", choices=["vits14", "synthetic"])
    ap.add_argument("--backbone-matmul", default="tf32", choices=["fp32", "tf32"],
                    help="precision of the torch matmuls inside the ViT backbone / camera-up head (boundary "
                         "components, PyTorch): tf32 = torch.set_float32_matmul_precision('high')")
    ap.add_argument("--cpu-sample-ellipsoids", type=int, default=64000,
                    help="CPU arm: ellipsoids whose rays (x29) are generated and timed per step: 64 reference-sized "
                         "(1000-ellipsoid, sampling.py:146-148) chunks of ~29k rays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--front-end", choices=("replicated", "sharded"), default="sharded",
                    help="N>1: each rank runs the image front end only for its B/N images (one more all-gather per batch; "
                         "the e2e leg then uploads each image once), or every rank for the whole batch")
    ap.add_argument("--per-query-sweeps", action="store_true",
                    help="score each query with its own two sweeps over the key cache instead of one sweep per pass "
                         "for the whole batch (tensor-core modes default to the batched kernel)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary figures (bf16 throughput mode, weighted-LS solve)")
    ap.add_argument("--solve", choices=("topk", "weighted_ls"), default="topk",
                    help="topk = the reference's evaluation path (top-100 rays -> LS, test.py:85-198); weighted_ls = all-ray "
                         "weighted least squares fused into the pass-2 epilogue + one all-reduce (least_squared_loss.py:47-64)")
    ap.add_argument("--config", choices=("c3", "c5"), default="c3",
                    help="c3 = BASELINE configs[2]/[3] (1M Gaussians, 8 queries per step); c5 = configs[4] (5M Gaussians, "
                         "32 queries per step, 8 GPUs); explicit --gaussians / --batch override")
    ap.add_argument("--emulate-shard", type=int, default=0, metavar="W",
                    help="single process, but generate and score only rank 0's share of a W-rank run (ray shard 0 of W, one "
                         "image of the batch per W): the per-rank kernel list of an N = W run for ncu, which cannot wrap torchrun")
    ap.add_argument("--no-latency", action="store_true", help="skip the one-query-per-step figure (profiling runs)")
    ap.add_argument("--no-breakdown", action="store_true", help="skip the eager per-stage CUDA-event breakdown")
    ap.add_argument("--batch", type=int, default=None,
                    help="queries per step: the image front end (resize, backbone, q projection, up head) runs once per "
                         "batch, the key cache is streamed per query; 1 = one query per step")
    args = ap.parse_args()
    if args.gaussians is None:
        args.gaussians = 5_000_000 if args.config == "c5" else 1_000_000
    if args.batch is None:
        args.batch = 32 if args.config == "c5" else 8
    return args


def load_peaks():
    """-> dict(hbm_gbs, bf16_tflops (burst), bf16_tflops_sustained, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1650.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.gpu = gpu_index
        self.p = None

    def start(self):
        if self.gpu is None:
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            self.f.close()
            os.unlink(self.f.name)
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        for stop in (self.p.terminate, self.p.kill):  # never let a stuck nvidia-smi hold the bench
            stop()
            try:
                self.p.wait(timeout=5)
                break
            except subprocess.TimeoutExpired:
                continue
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if "Active" in v and "Not" not in v:
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
REF_CHUNK = 1000  # the reference's ellipsoid cap per ray-generation call (sampling.py:146-148)


def cpu_query_rate(args, n_rays_total, sample_ellipsoids, steps=2, warmup=1, budget_s=240.0, threads=None):
    """Reference algorithm (oracle port) on the host: per query the reference recomputes the ray MLP, the attention
    over all rays, top-100 and the pose tail (identification_module.py:77-133, test.py:157-198).  One step = one query
    over the rays of `sample_ellipsoids` ellipsoids, generated for real in reference-sized chunks of 1000 ellipsoids
    (softmax statistics merged across chunks); the rate is extrapolated linearly in the ray count to n_rays_total.
    `steps` is reduced (never below 1) if warmup + steps would not fit `budget_s`; the line reports what was run."""
    sx = importlib.import_module("6dgs_b200")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    oracle = importlib.import_module("sixdgs_oracle")
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(threads or ncpu)
    sc = sx.synthetic.synth_scene(sample_ellipsoids, seed=0, extent=5.0)
    feats = torch.cat((sc["features_dc"], sc["features_rest"]), 1)
    t0 = time.perf_counter()
    parts = []
    valid = oracle.mask_degraded_ellipsoids(*torch.exp(sc["scaling"]).unbind(-1))
    nvalid = int(valid.sum())
    for lo in range(0, nvalid, REF_CHUNK):
        idx = torch.arange(lo, min(lo + REF_CHUNK, nvalid))
        parts.append(oracle.generate_rays(sc["xyz"], sc["scaling"], sc["rotation"], feats, ellipsoid_idx=idx))
    t_gen = time.perf_counter() - t0
    ori = torch.cat([p[0] for p in parts])
    dirs = torch.cat([p[1] for p in parts])
    rgb = torch.cat([p[2] for p in parts])
    n_chunks = len(parts)
    chunk_rays = max(1, -(-ori.shape[0] // max(n_chunks, 1)))
    w = sx.synthetic.synth_id_weights(seed=3)
    tok = torch.randn(256, 398, generator=torch.Generator().manual_seed(2))
    up = torch.tensor([0.0, 0.0, 1.0])

    def one_query(o, d, c):
        cache = {}

        def fea(lo, hi):
            if (lo, hi) not in cache:  # each chunk's features are computed once per query, as the reference does
                cache[(lo, hi)] = oracle.ray_features(o[lo:hi], d[lo:hi], c[lo:hi], w)
            return cache[(lo, hi)]

        n = o.shape[0]
        if n <= chunk_rays:
            # one reference-sized chunk = exactly what the reference as shipped executes: ONE softmax over the
            # materialised [n_img, n] map and its column sum (our_multihead_attention.py:4-12)
            scores, _ = oracle.attention_scores(tok, fea(0, n), w, return_map=False)
        else:  # uncapped equivalent: two sweeps with merged softmax statistics, no [n_img, n] map
            scores, _, _ = oracle.attention_scores_chunked(tok, fea, n, w, chunk=chunk_rays)
        top = torch.topk(scores, min(100, scores.shape[0]))
        return oracle.pose_tail(top.indices, top.values, o, d, up)[0]

    # torch's CPU GEMM/elementwise kernels do not always scale to every hardware thread of a big host:
    # give the reference arm its best thread count (all, half, 32, 16), probed on one chunk
    probe = (ori[:chunk_rays], dirs[:chunk_rays], rgb[:chunk_rays])
    if threads is None:
        best = (float("inf"), ncpu)
        for th in sorted({ncpu, max(1, ncpu // 2), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
            torch.set_num_threads(th)
            one_query(*probe)
            t0 = time.perf_counter()
            one_query(*probe)
            dt = time.perf_counter() - t0
            if dt < best[0]:
                best = (dt, th)
        threads = best[1]
        torch.set_num_threads(threads)
    t0 = time.perf_counter()
    one_query(*probe)
    t_shipped = time.perf_counter() - t0  # one query of the reference as shipped (1000-ellipsoid cap)
    est_step = t_shipped * n_chunks
    warmup = max(1, warmup)
    if est_step * (warmup + steps) > budget_s:
        warmup = 1
        steps = max(1, min(steps, int(budget_s / max(est_step, 1e-9)) - warmup))
    for _ in range(warmup):
        one_query(ori, dirs, rgb)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_query(ori, dirs, rgb)
    t_total = time.perf_counter() - t0
    t_step = t_total / steps
    per_ray = t_step / ori.shape[0]
    t_full = per_ray * n_rays_total
    return {"value": 1.0 / t_full, "unit": "queries/s", "cores": threads, "host_cpus": ncpu, "kind": "port",
            "extrapolated": True, "steps": steps, "warmup": warmup, "ms_per_step_sample": t_step * 1e3,
            "sample_rays": int(ori.shape[0]), "sample_chunks": n_chunks,
            "sample": (f"oracle port (torch CPU fp32, {threads} of {ncpu} host threads): per step one query = ray MLP + attention "
                       f"+ top-100 + pose tail over {ori.shape[0]} rays ({n_chunks} reference-sized chunks of {REF_CHUNK} "
                       f"ellipsoids, rays generated by the port, not tiled) in {t_step:.2f} s ({per_ray * 1e6:.2f} us/ray, "
                       f"mean of {steps} steps after {warmup} warm-up); value extrapolated linearly to {n_rays_total} rays; "
                       f"CPU ray generation {t_gen / max(ori.shape[0], 1) * 1e6:.1f} us/ray (per scene, not in value)"),
            "s_per_query_sample": t_step, "us_per_ray": per_ray * 1e6,
            # the reference as shipped caps the scene at 1000 ellipsoids whatever its size (sampling.py:146-148):
            # one query over the first chunk, and one ray-generation call, timed on their own
            "as_shipped": {"rays": int(probe[0].shape[0]), "ms_per_query": t_shipped * 1e3,
                           "raygen_s_per_call": t_gen / max(n_chunks, 1)}}


def time_score_kernels(sx, idm, cache, dev, warm, iters, nq, batched):
    """CUDA-event time of each score launch (pass 1, pass 2) on the current stream; `batched`: the multi-query kernel
    with nq queries per launch, else the single-query entry points (nq = 1)."""
    impl = idm._impl
    tok = torch.randn(nq, 256, 398, device=dev, generator=torch.Generator(device=dev).manual_seed(11))
    pw = idm.packed_weights()
    q = torch.stack([sx.ops.project_queries(tok[b], pw) for b in range(nq)])
    sb = torch.empty(nq, cache.n_rays, dtype=torch.float32, device=dev) if batched else None
    p1, p2 = [], []
    for i in range(warm + iters):
        a, b, b2, c = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        if batched:
            a.record()
            pm, pz = sx.ops.score_pass1_batch(cache.keys, q)
            b.record()
            parts = pm.shape[0] // nq
            mz = [sx.ops.score_merge(pm, pz, 256, rows=parts, first_row=j * parts) for j in range(nq)]
            m_, z_ = torch.stack([x[0] for x in mz]), torch.stack([x[1] for x in mz])
            b2.record()
            sx.ops.score_pass2_batch(cache.keys, q, m_, z_, out=sb)
            c.record()
        else:
            a.record()
            pm, pz = sx.ops.score_pass1(cache.keys, q[0], impl)
            b.record()
            m_, z_ = sx.ops.score_merge(pm, pz, 256)
            b2.record()
            sx.ops.score_pass2(cache.keys, q[0], m_, z_, impl, out=cache.scores)
            c.record()
        torch.cuda.synchronize()
        if i >= warm:
            p1.append(a.elapsed_time(b))
            p2.append(b2.elapsed_time(c))
    return p1, p2


def stage_breakdown(est, imgs, masks, local, world, reps=3):
    """ms per batch of each pipeline stage, eager launches bracketed by CUDA events (mean of `reps` after one warm-up).
    front = image front end (+ its all-gather when sharded); pass1 = statistics sweep; gather = the statistics
    all-gather; stage2 = merge + score sweep + per-query top-k / LS system; exchange = candidate all-gather or
    system all-reduce; tail = global top-k + pose tails.  The sum is larger than a graph-replayed step (launch gaps)."""
    names = ("front", "pass1", "gather_stats", "stage2", "exchange", "tail")
    acc = {k: 0.0 for k in names}
    for it in range(reps + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        ev[0].record()
        if est._shards_front(imgs.shape[0], local):
            q, up, valid = est._front(*est._local_chunk(imgs, masks, local))
            rec = est._all_gather(est._pack_front(q, up, valid))
            front = est._unpack_front(rec, q.shape[1], q.shape[2])
        else:
            front = est._front(imgs, masks)
        ev[1].record()
        st = est._pass1_all(*front)
        ev[2].record()
        pmz = est._all_gather(st["pmz"]) if world > 1 else st["pmz"]
        ev[3].record()
        if est.solve == "weighted_ls":
            ls_sys = est._stage2_weighted(pmz, st)
            ev[4].record()
            if world > 1:
                ls_sys = est._all_reduce_sum(ls_sys)
            ev[5].record()
            est._stage3_weighted(ls_sys, st)
        else:
            vals, idxs, cand = est._stage2(pmz, st, 100)
            ev[4].record()
            allc = est._all_gather(cand) if world > 1 else None
            ev[5].record()
            if world > 1:
                est._stage3(allc, st["up"], 100, st["nb"])
            else:
                [est.backend.pose_tail(est.ori, est.dirs, idxs[i], vals[i], st["up"][i]) for i in range(st["nb"])]
        ev[6].record()
        torch.cuda.synchronize()
        if it:
            for i, k in enumerate(names):
                acc[k] += ev[i].elapsed_time(ev[i + 1]) / reps
    acc["total"] = sum(acc[k] for k in names)
    acc["queries_per_batch"] = int(st["nb"])
    return acc


# ray counts of the seed-0 synthetic scenes (synth_scene(n, seed=0, extent=5.0)) as read back from the GPU arm's ray
# generation (config.n_rays of profiles/bench_r2_n1_c2_exact.json, bench_r2_n1_final.json, bench_r2_n8_c5_exact.json):
# the CPU arm extrapolates to the SAME number of rays the GPU arm scores; other sizes use the mean 29.05 rays / ellipsoid
MEASURED_RAYS = {100_000: 2_887_090, 1_000_000: 28_879_457, 5_000_000: 144_401_385}


def expected_rays(n_gaussians):
    return MEASURED_RAYS.get(int(n_gaussians), int(n_gaussians * 29.05))


def run_reference_arm(args):
    """`--impl reference`: the reference algorithm on this box's host cores (oracle port; the reference itself is a
    script tree that cannot travel to the GPU box).  steps / ms_per_step are what was actually run and measured (one
    step = one query over a bounded sample of the workload); value = the sample's per-ray cost extrapolated to the
    configured scene (`extrapolated: true`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rays = expected_rays(args.gaussians)
    base = cpu_query_rate(args, n_rays, args.cpu_sample_ellipsoids, steps=max(1, args.steps), warmup=max(1, args.warmup))
    cfg = workload_config(args, n_rays, None)
    cfg.update({"workload": f"{args.gaussians} synthetic Gaussians (all valid ellipsoids, uncapped), {args.height}x{args.width} "
                            "image, fp32 LS solve (BASELINE.json configs[2]): the reference algorithm on the host cores",
                "solve": "topk",
                "score_impl": "reference algorithm, fp32 torch CPU (ray MLP recomputed per query, no key cache)",
                "queries_per_step": 1, "backbone": "none (random 256x398 tokens; the ViT is outside the timed CPU path)",
                "backbone_matmul": None, "front_end": None, "score_sweeps": None, "parallelism": f"{base['cores']} host threads",
                "l2": None, "sample_rays_per_step": base["sample_rays"], "sample_chunks": base["sample_chunks"]})
    line = {"impl": "reference", "metric": "pose queries/sec, 1M-Gaussian scene", "value": base["value"], "unit": "queries/s",
            "n_gpus": args.gpus, "steps": base["steps"], "warmup": base["warmup"], "ms_per_step": base["ms_per_step_sample"],
            "extrapolated": True,
            "extrapolation": f"ms_per_step is measured on {base['sample_rays']} rays; value = 1 / (us_per_ray x {n_rays} rays)",
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, n_rays, n_rays_local):
    batched = args.score_impl.startswith("tc_") and not args.per_query_sweeps
    return {"workload": f"{args.gaussians} synthetic Gaussians (all valid ellipsoids, uncapped), {args.height}x{args.width} "
                        f"image, {args.score_impl} key cache, fp32 LS solve (BASELINE.json configs[{4 if args.config == 'c5' else (2 if args.gpus == 1 else 3)}])",
            "gaussians": args.gaussians, "n_rays": n_rays, "n_rays_per_rank": n_rays_local, "image": [args.height, args.width],
            "n_img_tokens": 256, "queries_per_step": args.batch, "score_impl": args.score_impl, "backbone": args.backbone, "backbone_matmul": args.backbone_matmul,
            "front_end": args.front_end if args.gpus > 1 else "single",
            "score_sweeps": "per batch (multi-query kernel)" if batched else "per query", "solve": args.solve,
            "parallelism": f"ray-shard x{args.gpus}", "l2": "inputs larger than L2 (key cache >> 126 MB), no flush needed"}


KEY_FORMATS = {  # score_impl -> (bytes per key row, MMA terms per logit, dtype label)
    "tc_f16x2": (1536, 3, "f16x2"), "tc_f16f8": (1536, 2, "f16f8"), "tc_bf16": (768, 1, "bf16"), "simt_bf16": (768, 1, "bf16"), "simt_fp32": (1536, 1, "f32")}


def build_roofline(args, peaks, n_local, nq, burst, sustained, batched, sm_mhz, sm_max_mhz):
    """Both roofs for each pass of the ray-score kernel.  Algorithmic bytes per launch: the key rows once (the batched
    kernel serves nq queries from one sweep; in f16x2 mode the lo halves are re-read per query from L2, not HBM) plus
    the outputs; MMA FLOPs per launch: 2 * 256 tokens * 384 * terms per ray per query."""
    row_bytes, terms, _ = KEY_FORMATS[args.score_impl]
    n_launch_q = nq if batched else 1          # queries covered by one timed pass
    sweeps = -(-n_launch_q // 8)               # the batched kernel takes up to 8 queries per launch (one key sweep each)
    flops = 2.0 * 256 * 384 * terms * n_local * n_launch_q
    bytes_p = {"pass1": sweeps * n_local * row_bytes + n_launch_q * 74 * 2 * 256 * 4,
               "pass2": sweeps * n_local * row_bytes + n_launch_q * n_local * 4}
    name = "score_tc_mq_kernel" if batched else ("score_tc_kernel" if args.score_impl.startswith("tc_") else "score_simt_kernel")

    def view(ms, which, tensor_peak):
        gbs = bytes_p[which] / (ms * 1e-3) / 1e9
        tf = flops / (ms * 1e-3) / 1e12
        return {"ms": ms, "GBps": gbs, "hbm_frac": gbs / peaks["hbm_gbs"], "TFLOPs": tf, "tensor_frac": tf / tensor_peak}

    det = {"sustained": {k: view(sum(v) / len(v), k, peaks["bf16_tflops_sustained"]) for k, v in sustained.items()},
           "burst": {k: view(min(v), k, peaks["bf16_tflops"]) for k, v in burst.items()},
           "note": "burst = the same launches timed alone on a cool GPU before the timed region (tensor roof = cuBLAS burst "
                   "figure); sustained = right after it in the timed region's thermal / power state (tensor roof = cuBLAS "
                   "sustained figure); `achieved` / `frac` are the sustained figures of the slower pass",
           "algorithmic_bytes_per_launch": bytes_p, "mma_flops_per_launch": flops, "queries_per_launch": n_launch_q,
           "kernel_launches_per_pass": sweeps,
           "key_row_bytes": row_bytes, "mma_terms_per_logit": terms,
           "tensor_peak_at_sampled_clock_tflops": (peaks["bf16_tflops"] * sm_mhz / sm_max_mhz) if sm_mhz and sm_max_mhz else None}
    dom = max(det["sustained"], key=lambda k: det["sustained"][k]["ms"])
    d = det["sustained"][dom]
    t_hbm = bytes_p[dom] / (peaks["hbm_gbs"] * 1e9)
    t_tensor = flops / (peaks["bf16_tflops_sustained"] * 1e12)
    bound = "tensor" if t_tensor >= t_hbm else "hbm"
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        per_ray = tj.get("bytes_per_ray_" + args.score_impl + ("_mq" if batched else ""), {}).get("score_" + dom)
        if per_ray:
            traffic = per_ray * n_local
            traffic_src = tj.get("source", "profiles/ncu_traffic.json") + " (bytes per ray from that capture x this run's rays)"
    roof = {"bound": bound, "kernel": f"{name}<{dom[-1]}> (score_{dom}, {args.score_impl})",
            "achieved": d["TFLOPs"] if bound == "tensor" else d["GBps"],
            "peak": peaks["bf16_tflops_sustained"] if bound == "tensor" else peaks["hbm_gbs"],
            "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
            "frac": d["tensor_frac"] if bound == "tensor" else d["hbm_frac"],
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"],
            "other_roof": {"bound": "hbm" if bound == "tensor" else "tensor",
                           "frac": d["hbm_frac"] if bound == "tensor" else d["tensor_frac"],
                           "time_bound_ms": {"hbm": t_hbm * 1e3, "tensor": t_tensor * 1e3}},
            "ms_per_launch": d["ms"], "detail": det}
    return roof


_T0 = time.perf_counter()


def trace(msg):
    """progress lines on stderr when SIXDGS_BENCH_TRACE=1 (every rank): where a multi-rank run spends its time"""
    if os.environ.get("SIXDGS_BENCH_TRACE") == "1":
        print(f"[bench r{os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def arm_watchdog():
    """A bench that cannot finish (a wedged collective, a kernel that never retires) must end on its own with a
    diagnostic instead of holding the GPU box until somebody else's limit kills it: after SIXDGS_BENCH_WATCHDOG_S
    seconds (default 1500; the default run takes about a minute; 0 disables) the process prints where it was and
    exits 124.  A daemon timer: it costs nothing on the normal path and dies with the process."""
    import threading

    limit = float(os.environ.get("SIXDGS_BENCH_WATCHDOG_S", "1500"))
    if limit <= 0:
        return None

    def fire():
        import faulthandler

        sys.stderr.write(f"[bench] watchdog: no result after {limit:.0f} s on rank {os.environ.get('RANK', '0')}; "
                         "stacks follow, exiting 124\n")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        sys.stderr.flush()
        os._exit(124)

    t = threading.Timer(limit, fire)
    t.daemon = True
    t.start()
    return t


def main():
    args = parse()
    arm_watchdog()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    sx = importlib.import_module("6dgs_b200")
    from importlib import import_module
    sharding = import_module("6dgs_b200.sharding")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.backbone_matmul == "tf32":
        torch.set_float32_matmul_precision("high")
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a collective that does not complete in 5 min aborts the job (NCCL watchdog) instead of hanging the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    trace("process group up")
    batched = args.score_impl.startswith("tc_") and not args.per_query_sweeps

    # ---------------- scene preparation (per scene; untimed for the metric, reported) ----------------
    t0 = time.perf_counter()
    sc = sx.synthetic.synth_scene(args.gaussians, seed=0, extent=5.0)
    scene = sx.GaussianScene.from_dict(sc, device=dev)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    shard = (rank, world) if world > 1 else ((0, args.emulate_shard) if args.emulate_shard > 1 else None)
    ori, dirs, rgb = sx.generate_all_possible_rays(scene, max_ellipsoids=None, shard=shard)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    trace(f"rays generated ({ori.shape[0]})")
    import warnings
    torch.manual_seed(0)  # identical (random-init) backbone / head weights on every rank and every run
    backbone = sx.synthetic.SyntheticBackbone() if args.backbone == "synthetic" else sx.DinoV2ViTS14()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        idm = sx.IdentificationModule("dino", backbone=backbone, score_impl=args.score_impl)
    idm.load_state_dict(sx.synthetic.synth_id_weights(seed=3), strict=False)
    idm = idm.to(dev).eval().requires_grad_(False)
    cache = idm.build_key_cache(ori, dirs, rgb)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    trace("key cache built")
    n_local = ori.shape[0]
    n_total = n_local
    if world > 1:
        t = torch.tensor([n_local], device=dev, dtype=torch.long)
        dist.all_reduce(t)
        n_total = int(t.item())
    est = sharding.ShardedPoseEstimator(idm, ori, dirs, cache, rank, world, front_end=args.front_end, multi_query=batched,
                                        solve=args.solve)

    B = args.batch
    img_u8 = torch.stack([(sx.synthetic.synth_image(args.height, args.width, seed=7 + i) * 255).to(torch.uint8) for i in range(B)])
    img_host = img_u8.pin_memory()
    img_dev = (img_u8.to(dev).float() / 255.0).contiguous()
    mask_dev = torch.ones(B, args.height, args.width, dtype=torch.bool, device=dev)

    # front_end="sharded": every rank holds (and, in the e2e leg, uploads) only its own B/world images
    own_only = args.front_end == "sharded" and world > 1 and B % world == 0
    lo, hi = (rank * (B // world), (rank + 1) * (B // world)) if own_only else (0, B)
    img_q, mask_q = img_dev[lo:hi], mask_dev[lo:hi]

    def query():
        return est.query_batch(img_q, mask_q, local=own_only)

    # one-query-per-step latency figure (eager + its own graphs), measured before the batched graphs are captured
    lat_b1 = None
    if B > 1 and not args.no_latency:
        for _ in range(3):
            est.query_batch(img_dev[:1], mask_dev[:1])
        if not args.no_graph:
            est.enable_cuda_graphs(img_dev[:1].clone(), mask_dev[:1].clone())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(10):
            est.query_batch(img_dev[:1], mask_dev[:1])
        l1.record()
        torch.cuda.synchronize()
        lat_b1 = l0.elapsed_time(l1) / 10
        est._g = None

    trace("latency figure done")
    # ---------------- warm-up, optional CUDA graph ----------------
    for _ in range(max(args.warmup, 3)):
        c2w, aux = query()
    torch.cuda.synchronize()
    trace("eager warm-up done")
    time.sleep(1.0)  # let the GPU cool before the burst figure
    b1, b2 = time_score_kernels(sx, idm, cache, dev, 1, 3, B, batched)
    burst = {"pass1": b1, "pass2": b2}
    graph = False
    if not args.no_graph:
        graph = est.enable_cuda_graphs(img_q, mask_q, local=own_only)
        if graph:
            g_out = query()
            torch.cuda.synchronize()
            if not torch.allclose(g_out[0], c2w, atol=1e-5, equal_nan=True):
                print("[bench] CUDA graph replay does not reproduce the eager pose; timing eager launches", file=sys.stderr)
                est._g = None
                graph = False

    def step():
        query()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    trace(f"graphs: {graph}")
    # ---------------- timed region: `value` ----------------
    sampler = ClockSampler(local if rank == 0 else None)  # the line is rank 0's; N polling nvidia-smi loops only contend
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.start()
    torch.cuda.profiler.start()  # no-op unless a profiler is attached (ncu --profile-from-start off)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    torch.cuda.profiler.stop()
    # ---------------- roofline timing: the score launches right after the timed region (same power / thermal state)
    s1, s2 = time_score_kernels(sx, idm, cache, dev, 1, 4, B, batched)
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = B * args.steps / (ms / 1e3)
    trace("timed region done")

    # ---------------- eager per-stage breakdown (CUDA events, outside the timed region; what does not shrink with N) -----
    breakdown = None
    if not args.no_breakdown:
        breakdown = stage_breakdown(est, img_q, mask_q, own_only, world)

    # ---------------- e2e: host image -> pose on host, through the public API ----------------
    pose_host = torch.empty(B, 4, 4).pin_memory()

    img_e2e = torch.empty_like(img_q)
    img_host_q = img_host[lo:hi]

    def e2e_step():
        d = img_host_q.to(dev, non_blocking=True)
        torch.div(d, 255.0, out=img_e2e)  # uint8 -> [0,1] float (test.py:69-73)
        m = torch.ones_like(img_e2e[..., 0], dtype=torch.bool)
        c, _ = est.query_batch(img_e2e, m, local=own_only)
        pose_host.copy_(c, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(3, args.warmup)):
        e2e_step()
    barrier()
    t0e = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(args.steps):
        e2e_step()
    ee1.record()
    wall_ms = (time.perf_counter() - t0e) * 1e3  # every step ends with a stream synchronize, so the host clock is valid too
    barrier()
    e2e_ms = max(ee0.elapsed_time(ee1), wall_ms)  # what the caller waits for: the slower of device and host view
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": B * args.steps / (e2e_ms / 1e3), "unit": "queries/s",
           "h2d_bytes_per_step": int(img_host_q.numel()) * (world if own_only else 1),
           "d2h_bytes_per_step": 64 * B, "ms_per_step": e2e_ms / args.steps}

    trace("e2e done")
    launches = (est.launches_per_query, est.launches_per_batch)
    peaks = load_peaks()
    roofline = build_roofline(args, peaks, n_local, B, burst, {"pass1": s1, "pass2": s2}, batched,
                              clocks.get("sm_mhz"), clocks.get("sm_max_mhz"))

    # ---------------- secondary figures on the same scene (N = 1 only; not the headline) ----------
    # (1) the fast variant of the exact mode: e4m3 cross terms, the key cache converted in place
    fast = None
    if world == 1 and args.score_impl == "tc_f16x2" and not args.no_secondary:
        try:
            est._g = None
            cache.keys = sx.ops.keys_to_f16f8(cache.keys)
            for _ in range(3):
                est.query_batch(img_dev, mask_dev)
            if not args.no_graph:
                est.enable_cuda_graphs(img_dev, mask_dev)
            torch.cuda.synchronize()
            n2 = max(5, min(args.steps, 20))
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(n2):
                est.query_batch(img_dev, mask_dev)
            f1.record()
            torch.cuda.synchronize()
            fast = {"score_impl": "tc_f16f8", "value": B * n2 / (f0.elapsed_time(f1) / 1e3), "unit": "queries/s", "steps": n2,
                    "note": "main MMA term in fp16, the two 2^-11 cross terms as e4m3 MMAs at twice the rate (2 term-units instead "
                            "of 3).  Parity-green in tests/test_gpu_exact_tc.py (scores <= 1e-3, identical top-100, pose <= 1e-4) "
                            "but its error grows with the logit spread: 2.8e-4 at logit std 6.5, 8.4e-4 on this scene at 9.7 -- no "
                            "margin beyond ~12, which is why the headline stays tc_f16x2 (1.4e-4 at 9.7)"}
        except Exception as e:  # noqa: BLE001
            fast = {"score_impl": "tc_f16f8", "error": f"{type(e).__name__}: {e}"}
    # (2) the bf16 throughput mode
    secondary = None
    if world == 1 and args.score_impl == "tc_f16x2" and not args.no_secondary:
        try:
            est._g = None
            del est
            cache.keys = None
            torch.cuda.empty_cache()
            idm2 = sx.IdentificationModule("dino", backbone=backbone, score_impl="tc_bf16")
            idm2.load_state_dict(sx.synthetic.synth_id_weights(seed=3), strict=False)
            idm2 = idm2.to(dev).eval().requires_grad_(False)
            cache2 = idm2.build_key_cache(ori, dirs, rgb)
            est2 = sharding.ShardedPoseEstimator(idm2, ori, dirs, cache2, 0, 1, multi_query=True)
            for _ in range(3):
                est2.query_batch(img_dev, mask_dev)
            if not args.no_graph:
                est2.enable_cuda_graphs(img_dev, mask_dev)
            torch.cuda.synchronize()
            n2 = max(5, min(args.steps, 20))
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(n2):
                est2.query_batch(img_dev, mask_dev)
            f1.record()
            torch.cuda.synchronize()
            # the HBM-bound formulation (one query per sweep over 768-B keys, score_tc.cu): the "ray-score HBM GB/s" figure
            h1, h2 = time_score_kernels(sx, idm2, cache2, dev, 1, 4, 1, False)
            hb = {}
            for nm, ts, by in (("pass1", h1, n_local * 768 + 74 * 2 * 256 * 4), ("pass2", h2, n_local * 772)):
                ms_ = sum(ts) / len(ts)
                hb[nm] = {"ms": ms_, "GBps": by / ms_ / 1e6, "hbm_frac": by / ms_ / 1e6 / peaks["hbm_gbs"],
                          "TFLOPs": 2.0 * 256 * 384 * n_local / ms_ / 1e9}
            secondary = {"score_impl": "tc_bf16", "value": B * n2 / (f0.elapsed_time(f1) / 1e3), "unit": "queries/s", "steps": n2,
                         "hbm_bound_kernel": {"kernel": "score_tc_kernel<1|2> (bf16 keys, one query per sweep)", "sustained": hb,
                                              "note": "256 MMA-FLOP per key byte = on the ridge: this is the formulation whose roof is "
                                                      "HBM; measured right after the throughput-mode steps (power-capped state)"},
                         "note": "throughput mode (one bf16 MMA term, 768 B/ray, TF32 key build): 3e-2 score tolerance on flat "
                                 "logits, NOT parity-green on a peaked softmax (tests/test_gpu_exact_tc.py) -- reported for "
                                 "reference only, the headline is the exact mode"}
        except Exception as e:  # noqa: BLE001
            secondary = {"score_impl": "tc_bf16", "error": f"{type(e).__name__}: {e}"}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_query_rate(args, n_total, args.cpu_sample_ellipsoids, steps=2, warmup=1, budget_s=60.0)

    if rank == 0:
        lpq, lpb = launches
        line = {"metric": "pose queries/sec, 1M-Gaussian scene", "value": value, "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": KEY_FORMATS[args.score_impl][2],
                "data": "synthetic", "config": workload_config(args, n_total, n_local), "clocks": clocks, "e2e": e2e,
                "gpu_launches": (lpq * B + lpb) * args.steps, "cuda_graph": bool(graph),
                "queries_per_step": B, "latency_b1": {"ms_per_query": lat_b1, "queries_per_s": (1e3 / lat_b1) if lat_b1 else None},
                "roofline": roofline, "cpu_baseline": cpu, "fast_mode": fast, "throughput_mode": secondary, "breakdown_ms_per_batch": breakdown,
                "parity": "tests/test_gpu_exact_tc.py (scores <= 1e-3 rel, pose <= 1e-4 vs the reference fixtures incl. a peaked "
                          "softmax, and every score of this 1M-Gaussian scene vs fp64)" if args.score_impl == "tc_f16x2" else
                          "throughput / alternative mode; see tests/test_gpu_parity.py for its tolerance",
                "prepare": {"scene_to_gpu_s": t1 - t0, "raygen_s": t2 - t1, "key_cache_s": t3 - t2,
                            "rays_per_s_raygen": n_local / max(t2 - t1, 1e-9),
                            # cold = scene preparation (rays + key cache, once per scene / weight update) + one query
                            "cold_first_query_s": (t3 - t1) + ms / 1e3 / (args.steps * B)},
                "pose": {"centre": [float(x) for x in c2w[0, :3, 3].tolist()], "status": int(aux[0, 7].item())}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
