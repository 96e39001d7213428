#!/usr/bin/env python
"""bench.py -- pose queries/sec of the 6DGS hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one rank per GPU
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

Workload (config.workload): BASELINE.json configs[2] -- a synthetic 1M-Gaussian scene (every valid
ellipsoid casts rays, ~29 rays each), one 1080x1920 query image, bf16 key cache scored on the
tcgen05 path, fp32 LS solve.  A "step" is one batch of `--batch` (default 8) pose queries, each with its own
image: images -> backbone tokens -> q (once per batch; latency-bound, so 8 images cost about one) and then per
query two streaming passes over the key cache -> top-100 -> fused LS pose tail -> c2w.  `value` counts
QUERIES per second; `latency_b1` in the same line is the one-query-per-step figure.  Scene preparation
(ray generation + key cache) is per scene, not per query, and is reported separately.

  value     queries/s, image already resident in HBM, whole query captured in one CUDA graph
  e2e       queries/s through ShardedPoseEstimator.query_batch() from pinned uint8 HOST images
            (H2D + /255 + mask inside the timed region) to the 4x4 poses back on the host
  roofline  ray-score kernels (score_tc pass 1 / pass 2), algorithmic bytes / CUDA-event time / measured HBM peak
  cpu_baseline  the oracle port (torch CPU, all host threads) on a bounded sample of the same rays,
            extrapolated linearly in the ray count (per-ray work is independent; stated in `sample`)

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
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--score-impl", default="tc_bf16", choices=["tc_bf16", "simt_bf16", "simt_fp32"])
    ap.add_argument("--backbone", default="vits14", choices=["vits14", "synthetic"])
    ap.add_argument("--backbone-matmul", default="tf32", choices=["fp32", "tf32"],
                    help="precision of the torch matmuls inside the ViT backbone / camera-up head (boundary "
                         "components, PyTorch): tf32 = torch.set_float32_matmul_precision('high')")
    ap.add_argument("--cpu-sample-ellipsoids", type=int, default=16000,
                    help="CPU arm: ellipsoids whose rays (x29) are timed, 16 reference-sized chunks of 29k rays; ~10-20 s")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--front-end", choices=("replicated", "sharded"), default="replicated",
                    help="N>1: every rank runs the image front end for the whole batch, or only for its B/N images "
                         "(one more all-gather per batch; the e2e leg then uploads each image once)")
    ap.add_argument("--fused-topk", action="store_true", help="EXPERIMENTAL: sixdgs_topk_fused (7 launches instead of 11)")
    ap.add_argument("--multi-query", action="store_true",
                    help="EXPERIMENTAL: score the whole batch in one sweep over the key cache per pass (score_tc_mq.cu)")
    ap.add_argument("--batch", type=int, default=8,
                    help="queries per step: the image front end (resize, backbone, q projection, up head) runs once per "
                         "batch, the key cache is streamed per query; 1 = one query per step")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
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
def cpu_query_rate(args, n_rays_total, sample_ellipsoids, threads=None):
    """Reference algorithm (oracle port) on the host: per query the reference recomputes the ray MLP,
    the attention over all rays, top-100 and the pose tail (identification_module.py:77-133, test.py:157-198).
    Timed on `sample_ellipsoids` ellipsoids' rays, extrapolated linearly to n_rays_total."""
    sx = importlib.import_module("6dgs_b200")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    oracle = importlib.import_module("sixdgs_oracle")
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(threads or ncpu)
    sc = sx.synthetic.synth_scene(sample_ellipsoids, seed=0, extent=5.0)
    feats = torch.cat((sc["features_dc"], sc["features_rest"]), 1)
    t0 = time.perf_counter()
    # ray generation in 1000-ellipsoid chunks like the reference's cap (per scene, reported separately)
    parts = []
    valid = oracle.mask_degraded_ellipsoids(*torch.exp(sc["scaling"]).unbind(-1))
    nvalid = int(valid.sum())
    for lo in range(0, min(nvalid, 2000), 1000):
        idx = torch.arange(lo, min(lo + 1000, nvalid))
        parts.append(oracle.generate_rays(sc["xyz"], sc["scaling"], sc["rotation"], feats, ellipsoid_idx=idx))
    t_gen = time.perf_counter() - t0
    ori = torch.cat([p[0] for p in parts])
    dirs = torch.cat([p[1] for p in parts])
    rgb = torch.cat([p[2] for p in parts])
    gen_rays = ori.shape[0]
    # tile the generated rays up to the sample size (the per-ray cost does not depend on the values)
    want = int(sample_ellipsoids * 29)
    rep = max(1, math.ceil(want / gen_rays))
    ori, dirs, rgb = ori.repeat(rep, 1)[:want], dirs.repeat(rep, 1)[:want], rgb.repeat(rep, 1)[:want]
    w = sx.synthetic.synth_id_weights(seed=3)
    tok = torch.randn(256, 398, generator=torch.Generator().manual_seed(2))
    up = torch.tensor([0.0, 0.0, 1.0])

    def one_query():
        cache = {}

        def fea(lo, hi):
            if (lo, hi) not in cache:  # each chunk's features are computed once per query, as the reference does
                cache[(lo, hi)] = oracle.ray_features(ori[lo:hi], dirs[lo:hi], rgb[lo:hi], w)
            return cache[(lo, hi)]

        scores, _, _ = oracle.attention_scores_chunked(tok, fea, ori.shape[0], w, chunk=29000)
        top = torch.topk(scores, 100)
        return oracle.pose_tail(top.indices, top.values, ori, dirs, up)[0]

    # torch's CPU GEMM/elementwise kernels do not always scale to every hardware thread of a big host:
    # give the reference arm its best thread count (all, half, 32, 16), probed on a 29k-ray slice
    if threads is None:
        full = (ori, dirs, rgb)
        ori, dirs, rgb = ori[:29000], dirs[:29000], rgb[:29000]
        best = (float("inf"), ncpu)
        for th in sorted({ncpu, max(1, ncpu // 2), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
            torch.set_num_threads(th)
            one_query()
            t0 = time.perf_counter()
            one_query()
            dt = time.perf_counter() - t0
            if dt < best[0]:
                best = (dt, th)
        threads = best[1]
        torch.set_num_threads(threads)
        ori, dirs, rgb = full
    one_query()
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        one_query()
        ts.append(time.perf_counter() - t0)
    t_sample = min(ts)
    per_ray = t_sample / ori.shape[0]
    t_full = per_ray * n_rays_total
    return {"value": 1.0 / t_full, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": (f"oracle port (torch CPU fp32, {threads} threads): ray MLP + attention + top-100 + pose tail on "
                       f"{ori.shape[0]} rays in {t_sample:.2f} s/query ({per_ray * 1e6:.2f} us/ray), extrapolated linearly "
                       f"to {n_rays_total} rays; CPU ray generation {t_gen / max(gen_rays, 1) * 1e6:.1f} us/ray (per scene)"),
            "s_per_query_sample": t_sample, "us_per_ray": per_ray * 1e6}


def time_score_kernels(sx, idm, cache, dev, warm, iters):
    """CUDA-event time of each score launch (pass 1, pass 2) on the current stream."""
    impl = idm._impl
    tok = torch.randn(256, 398, device=dev, generator=torch.Generator(device=dev).manual_seed(11))
    q = sx.ops.project_queries(tok, idm.packed_weights())
    p1, p2 = [], []
    for i in range(warm + iters):
        a, b, b2, c = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        a.record()
        pm, pz = sx.ops.score_pass1(cache.keys, q, impl)
        b.record()
        m_, z_ = sx.ops.score_merge(pm, pz, 256)
        b2.record()
        sx.ops.score_pass2(cache.keys, q, m_, z_, impl, out=cache.scores)
        c.record()
        torch.cuda.synchronize()
        if i >= warm:
            p1.append(a.elapsed_time(b))
            p2.append(b2.elapsed_time(c))
    return p1, p2


def expected_rays(n_gaussians):
    return int(n_gaussians * 29.05)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rays = expected_rays(args.gaussians)
    steps = max(1, min(args.steps, 3))
    base = cpu_query_rate(args, n_rays, args.cpu_sample_ellipsoids)
    line = {"impl": "reference", "metric": "pose queries/sec, 1M-Gaussian scene", "value": base["value"], "unit": "queries/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 / base["value"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n_rays, None), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, n_rays, n_rays_local):
    return {"workload": f"{args.gaussians} synthetic Gaussians (all valid ellipsoids, uncapped), {args.height}x{args.width} "
                        f"image, {args.score_impl} key cache, fp32 LS solve (BASELINE.json configs[2])",
            "gaussians": args.gaussians, "n_rays": n_rays, "n_rays_per_rank": n_rays_local, "image": [args.height, args.width],
            "n_img_tokens": 256, "queries_per_step": args.batch, "score_impl": args.score_impl, "backbone": args.backbone, "backbone_matmul": args.backbone_matmul,
            "front_end": args.front_end if args.gpus > 1 else "single",
            "score_sweeps": "per batch (multi-query kernel)" if args.multi_query else "per query",
            "parallelism": f"ray-shard x{args.gpus}", "l2": "inputs larger than L2 (key cache >> 126 MB), no flush needed"}


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
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
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    # ---------------- scene preparation (per scene; untimed for the metric, reported) ----------------
    t0 = time.perf_counter()
    sc = sx.synthetic.synth_scene(args.gaussians, seed=0, extent=5.0)
    scene = sx.GaussianScene.from_dict(sc, device=dev)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ori, dirs, rgb = sx.generate_all_possible_rays(scene, max_ellipsoids=None, shard=(rank, world) if world > 1 else None)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
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
    n_local = ori.shape[0]
    n_total = n_local
    if world > 1:
        t = torch.tensor([n_local], device=dev, dtype=torch.long)
        dist.all_reduce(t)
        n_total = int(t.item())
    est = sharding.ShardedPoseEstimator(idm, ori, dirs, cache, rank, world, front_end=args.front_end,
                                        multi_query=args.multi_query,
                                        backend=sharding.CudaBackend(idm, fused_topk=args.fused_topk))

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
    if B > 1:
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

    # ---------------- warm-up, optional CUDA graph ----------------
    for _ in range(max(args.warmup, 3)):
        c2w, aux = query()
    torch.cuda.synchronize()
    burst1, burst2 = time_score_kernels(sx, idm, cache, dev, 2, 4)
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

    # ---------------- timed region: `value` ----------------
    sampler = ClockSampler(local)
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
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = B * args.steps / (ms / 1e3)

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
    e2e = {"value": B * args.steps / (e2e_ms / 1e3), "unit": "queries/s", "h2d_bytes_per_step": int(img_host.numel()),
           "d2h_bytes_per_step": 64 * B, "ms_per_step": e2e_ms / args.steps}

    # ---------------- roofline of the ray-score kernels (CUDA events around each launch) ----------------
    # sustained = right after the timed regions (same thermal / power-cap state as the timed steps);
    # burst = the same loop run before them, on a cool GPU (reported in detail.burst)
    peak, peak_src = load_peaks()
    p1, p2 = time_score_kernels(sx, idm, cache, dev, 3, 10)
    kbytes = cache.keys.element_size() * 384
    t1_ms, t2_ms = sum(p1) / len(p1), sum(p2) / len(p2)
    bytes1 = n_local * kbytes + est.parts * 2 * 256 * 4
    bytes2 = n_local * kbytes + n_local * 4
    ach1, ach2 = bytes1 / (t1_ms * 1e-3) / 1e9, bytes2 / (t2_ms * 1e-3) / 1e9
    dom = ("score_pass1", ach1, t1_ms, bytes1) if t1_ms >= t2_ms else ("score_pass2", ach2, t2_ms, bytes2)
    # DRAM traffic per launch from the committed `ncu --set full` capture (bytes per ray x this run's rays)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        per_ray = json.load(open(tp)).get("bytes_per_ray", {}).get(dom[0])
        traffic = per_ray * n_local if per_ray else None
    roofline = {"bound": "hbm", "kernel": f"score_tc_kernel<{1 if dom[0] == 'score_pass1' else 2}> ({dom[0]})",
                "achieved": dom[1], "peak": peak, "unit": "GB/s", "frac": dom[1] / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom[3], "ms_per_launch": dom[2],
                "detail": {"burst": {"pass1_ms": min(burst1), "pass2_ms": min(burst2),
                                     "pass1_frac": bytes1 / (min(burst1) * 1e-3) / 1e9 / peak,
                                     "pass2_frac": bytes2 / (min(burst2) * 1e-3) / 1e9 / peak,
                                     "note": "same launches timed on a cool GPU before the timed region; `achieved` is the "
                                             "sustained figure measured right after it under sw_power_cap"},
                           "pass1": {"ms": t1_ms, "GBps": ach1, "frac": ach1 / peak},
                           "pass2": {"ms": t2_ms, "GBps": ach2, "frac": ach2 / peak},
                           "flops_per_launch": 2.0 * 256 * 384 * n_local,
                           "tflops_pass2": 2.0 * 256 * 384 * n_local / (t2_ms * 1e-3) / 1e12}}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_query_rate(args, n_total, args.cpu_sample_ellipsoids)

    if rank == 0:
        line = {"metric": "pose queries/sec, 1M-Gaussian scene", "value": value, "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if "bf16" in args.score_impl else "f32",
                "data": "synthetic", "config": workload_config(args, n_total, n_local), "clocks": clocks, "e2e": e2e,
                "gpu_launches": (est.launches_per_query * B + est.launches_per_batch) * args.steps, "cuda_graph": bool(graph),
                "queries_per_step": B, "latency_b1": {"ms_per_query": lat_b1, "queries_per_s": (1e3 / lat_b1) if lat_b1 else None},
                "roofline": roofline, "cpu_baseline": cpu,
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
