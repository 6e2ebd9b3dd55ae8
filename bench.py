#!/usr/bin/env python
"""bench.py -- benchmark of the B200 path-tracing core (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

HEADLINE (top-level keys) = BASELINE.json configs[2]: Cornell box 1024x1024, max depth 8, a fixed
1024 spp per frame, sample-sharded over the N ranks with ONE NCCL all-reduce of the fp32
accumulation buffers per frame.  One STEP = one frame (all ranks together), so the job is the
same at every N ("scaling": "strong") and `value` = spp/s of the whole job.  The frame runs
through pyrenderer_b200.core.tracing.render_distributed -> prt_render_sharded (C ABI); bounce 0
runs in exact mode (bit-exact primary-hit ids).  `e2e` = the same frame with a pinned HOST
accumulation buffer, upload and download inside the timed region.

`closest_hit` = BASELINE.json configs[3] + north_star ">= 2 Grays/s closest-hit on a 1M-triangle
scene per B200": 2^24 incoherent rays per step against a 1 000 000-triangle soup, in EXACT mode
(ids bit-exact against the reference's intersection code -- the same persistent kernel, checked
against the CPU oracle inside this run) and in plain FP32 mode, with the traversal roofline; every
rank traces its own batches (no collective).  `soup10m` = the 10M-triangle soup (the HBM-bound
case).  `c5` = BASELINE.json configs[4] (subdivided Cornell, 4.7M triangles, dielectric +
conductor, 3840x2160), a bounded number of spp per frame, sharded the same way.  `c1` =
BASELINE.json configs[0] on the GPU next to the CPU port.

`--impl reference` times the reference's CPU algorithm (oracle port of main.py's pixel x sample
loop over core/tracing.py's estimator, brute-force intersection as core/scene.py:66-73) on the
headline config with every host thread, a bounded number of spp per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SOUP_TRIS = 1_000_000
SOUP10_TRIS = 10_000_000
RAYS_PER_BATCH = 1 << 24
N_BATCHES = 4
HBM_FALLBACK_GBS = 6650.0
C3_RES, C3_DEPTH = 1024, 8


def soup(n, seed=7):
    rng = np.random.default_rng(seed)
    h = 0.75 * n ** (-1.0 / 3.0)
    c = rng.uniform(0, 1, (n, 1, 3))
    e = rng.uniform(-h, h, (n, 2, 3))
    return np.concatenate([c, c + e[:, :1], c + e[:, 1:]], 1).astype(np.float32)


def host_rays(n, seed=11):
    rng = np.random.default_rng(seed)
    r = np.empty((n, 8), np.float32)
    r[:, 0:3] = rng.uniform(0, 1, (n, 3))
    d = rng.normal(size=(n, 3))
    r[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    r[:, 3] = 1e-5
    r[:, 7] = 3.4e38
    return r


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def host_threads():
    """Host threads this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def ncu_profile(name):
    """Counters of one ncu capture (profiles/ncu_<name>.json, written by profiles/ncu_extract.py).
    They cannot be measured inside a timed run; they are quoted only while the CUDA sources still
    have the fingerprint the capture was taken from."""
    from pyrenderer_b200.kernel_fingerprint import fingerprint
    p = os.path.join(ROOT, "profiles", f"ncu_{name}.json")
    if not os.path.exists(p):
        return None
    rec = json.load(open(p))
    if rec.get("source_fingerprint") != fingerprint():
        sys.stderr.write(f"[bench] WARNING: profiles/ncu_{name}.json is STALE (captured from sources "
                         f"{rec.get('source_fingerprint')}, now {fingerprint()}): its counters are not quoted\n")
        return {"stale": True, "source": f"profiles/ncu_{name}.json", "captured_fingerprint": rec.get("source_fingerprint")}
    rec["source"] = f"profiles/ncu_{name}.json"
    return rec


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 100 ms from `start()` on; `summary()`
    keeps the samples taken between `mark_begin()` and `mark_end()` (the timed region); if the
    region was shorter than one sampling period it falls back to the samples since start()."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for line in self.proc.stdout:
                    self.rows.append((time.time(), line))
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
            time.sleep(0.5)  # let nvidia-smi come up before the GPU work starts
        except Exception:
            self.proc = None
        return self

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, line in rows:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e18) + 0.1]
        scope = "timed region"
        sm, mx, reasons = parse(inside)
        if not sm:
            sm, mx, reasons = parse(self.rows)
            scope = "warm-up + timed region"
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "scope": scope}


def c3_workload(total_spp):
    return (f"cornell-box {C3_RES}x{C3_RES}, max depth {C3_DEPTH}, {total_spp} spp per frame "
            "(BASELINE configs[2]), sample-sharded over the ranks, one all-reduce of the fp32 accumulation buffer per frame")


def load_cornell():
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    from pyrenderer_b200.main import DEFAULT_SCENE
    scene, cam = read_file(DEFAULT_SCENE)
    assert cam.get_resolution() == [C3_RES, C3_RES], cam.get_resolution()
    return scene, cam


def run_reference(args):
    """CPU arm: oracle port of the reference's render loop on the headline config, all host threads.
    Under torchrun only rank 0 works; the other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    scene, cam = load_cornell()
    a = scene.arrays()
    iview, sw, sh, focal, W, H = cam.device_record()
    ocam = oracle.make_camera(iview, sw, sh, focal, W, H)
    cores = host_threads()
    spp = 2  # samples per pixel per step: ~2e6 paths of depth <= 8 against 36 triangles (about a second on 16 threads)
    times = []
    for s in range(args.warmup + args.steps):
        P = oracle.make_params(seed=1, spp_begin=s * spp, spp_end=(s + 1) * spp, max_depth=C3_DEPTH)
        t0 = time.perf_counter()
        oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"], ocam, P, nthreads=cores)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    v = spp / (ms * 1e-3)
    sample = f"{spp} spp of the {args.total_spp}-spp frame per step ({W * H * spp} paths), {cores} host threads"
    line = {"impl": "reference", "metric": "cornell-box render spp/s", "value": v, "unit": "spp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": c3_workload(args.total_spp), "sample": sample,
                       "note": "reference algorithm: main.py:28-55 loop over core/tracing.py:116-155, every triangle tested per ray "
                               "(core/scene.py:66-73); C + OpenMP port (oracle/pt_oracle.c) -- the reference's Python does not run at HEAD"},
            "cpu_baseline": {"value": v, "unit": "spp/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "spp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--total-spp", type=int, default=1024, help="samples per pixel of one Cornell frame (all ranks together)")
    ap.add_argument("--c5-spp", type=int, default=16, help="samples per pixel of one C5 frame (all ranks together)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-soup", action="store_true")
    ap.add_argument("--skip-soup10m", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from pyrenderer_b200 import _abi
    from pyrenderer_b200.core import tracing
    from pyrenderer_b200.core.scene import Scene
    from pyrenderer_b200.core.camera import Camera

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ================================================================== headline: C3
    scene, cam = load_cornell()
    W = H = C3_RES
    rctx = scene.commit(local)
    if world > 1:
        tracing.init_distributed(scene, device=local)
    total = args.total_spp
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)

    def frame(k, acc):
        # frame k = samples [k*total, (k+1)*total) of every pixel: all K frames accumulate into one image
        tracing.render_distributed(scene, cam, total, max_depth=C3_DEPTH, seed=1, spp_begin=k * total, device=local, accum=acc)

    clk = ClockSampler(local).start()
    for k in range(args.warmup):
        frame(k, accum)
    rctx.reset_counters()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rctx.profile_begin()  # per-kernel-class cudaEvent pairs on the launching stream, over the timed region
    clk.mark_begin()
    e0.record()
    for k in range(args.steps):
        frame(args.warmup + k, accum)
    e1.record()
    barrier()
    clk.mark_end()
    clk.stop()
    frame_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    prof = rctx.profile_end()
    rc = rctx.counters()
    clocks = clk.summary()
    value = total / (frame_ms * 1e-3)
    rays_frame = sum_over_ranks((rc["rays_closest"] + rc["rays_shadow"]) / args.steps)
    launches_frame = sum(v[1] for v in prof.values()) / args.steps + 2  # + memset of the shard buffer, add_into_kernel
    kernel_ms = {k: v[0] / args.steps for k, v in prof.items()}
    n_acc = float(accum[..., 3].mean().item())
    assert abs(n_acc - (args.warmup + args.steps) * total) < 0.5, f"accumulated {n_acc} samples per pixel"

    # counted twin of one wave (16 spp) outside the timed region: node visits / triangle tests per ray of
    # the render's own ray population
    acc_c = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    rctx.reset_counters()
    rctx.render(rctx.render_params(seed=1, spp_begin=0, spp_end=16, max_depth=C3_DEPTH,
                                   flags=_abi.RENDER_EXACT_PRIMARY | _abi.RENDER_COUNT), acc_c)
    cc = rctx.counters()
    n_rays_c = cc["rays_closest"] + cc["rays_shadow"]
    c3_n_node, c3_n_tri = cc["node_visits"] / n_rays_c, cc["tri_tests"] / n_rays_c
    # cost of the exact bounce 0: the same 16 spp with and without PRT_RENDER_EXACT_PRIMARY
    def timed_wave(flags):
        best = 1e9
        for rep in range(3):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            rctx.render(rctx.render_params(seed=1, spp_begin=16 * rep, spp_end=16 * rep + 16, max_depth=C3_DEPTH, flags=flags), acc_c)
            g1.record()
            torch.cuda.synchronize()
            best = min(best, g0.elapsed_time(g1))
        return best
    wave_exact_ms, wave_plain_ms = timed_wave(_abi.RENDER_EXACT_PRIMARY), timed_wave(0)
    del acc_c

    # e2e: the same frame with a pinned HOST accumulation buffer (upload + frame + download per step)
    e2e_steps = max(2, min(args.steps, 3))
    host_acc = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
    dev_acc = torch.empty((H, W, 4), dtype=torch.float32, device=dev)

    def e2e_frame(k):
        dev_acc.copy_(host_acc, non_blocking=True)
        frame(1000 + k, dev_acc)
        host_acc.copy_(dev_acc, non_blocking=True)
        torch.cuda.synchronize()

    e2e_frame(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_frame(1 + k)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    assert abs(float(host_acc[..., 3].mean()) - (1 + e2e_steps) * total) < 0.5
    acc_bytes = W * H * 16
    del host_acc, dev_acc

    # roofline of the frame's dominant kernel class (closest-hit traversal: exact persistent kernel on
    # bounce 0, closest_kernel on the others)
    prof_c = ncu_profile("cornell_closest")
    c3_bytes_per_ray = 52.0 + 64.0 * c3_n_node + 40.0 * c3_n_tri  # 32 B ray + 16 B hit + 4 B queue entry; 64 B / record, 40 B / triangle
    closest_rays_frame = rc["rays_closest"] / args.steps  # this rank
    c3_achieved = c3_bytes_per_ray * closest_rays_frame / (kernel_ms["closest"] * 1e-3) / 1e9 if kernel_ms["closest"] > 0 else 0.0
    c3_roof = {
        "bound": "issue (instruction issue + SIMT divergence): the 7-record BVH of the Cornell box lives in L1; only the 52 B/ray "
                 "of ray, hit and queue records move through HBM",
        "kernel": "prt::closest_kernel / prt::trace_persistent_kernel<CLOSEST,EXACT> (persist.cuh), all launches of the class in the timed region",
        "achieved": c3_achieved, "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": c3_achieved / hbm_peak,
        "frac_meaning": "SURVEY 8(d) nominal: algorithmic bytes (52 + 64 N_node + 40 N_tri per ray) / kernel time / HBM peak -- "
                        "NOT a DRAM fraction; see dram_frac",
        "bytes_per_ray": c3_bytes_per_ray, "n_node": c3_n_node, "n_tri": c3_n_tri,
        "kernel_ms_per_frame": kernel_ms["closest"], "share_of_frame": kernel_ms["closest"] / frame_ms,
        "kernel_ms_by_class_per_frame": kernel_ms,
        "streamed_GBps": 52.0 * closest_rays_frame / (kernel_ms["closest"] * 1e-3) / 1e9 if kernel_ms["closest"] > 0 else 0.0,
        "traffic": None, "ncu": prof_c,
    }
    if prof_c and not prof_c.get("stale"):
        # ncu captured ONE launch (a 16-spp wave, bounce 1); scale its bytes per ray to this rank's frame
        c3_roof["traffic"] = prof_c["dram_bytes"] / prof_c["rays_in_launch"] * closest_rays_frame if prof_c.get("rays_in_launch") else None
        c3_roof["dram_frac"] = prof_c.get("dram_throughput_pct", 0.0) / 100.0
        c3_roof["issue_active_frac"] = prof_c.get("issue_active_pct", 0.0) / 100.0
        c3_roof["simt_lanes_per_warp"] = prof_c.get("lanes_per_instruction")
    prof_s = ncu_profile("cornell_shade")
    shade_paths_frame = rc["rays_closest"] / args.steps
    shade_roof = {
        "kernel": "prt::shade_kernel<false,false> (wavefront.cu)", "bound": "hbm / latency",
        "bytes_per_path_bounce": 170.0,
        "achieved": 170.0 * shade_paths_frame / (kernel_ms["shade"] * 1e-3) / 1e9 if kernel_ms["shade"] > 0 else 0.0,
        "peak": hbm_peak, "unit": "GB/s", "kernel_ms_per_frame": kernel_ms["shade"], "share_of_frame": kernel_ms["shade"] / frame_ms,
        "ncu": prof_s,
    }
    shade_roof["frac"] = shade_roof["achieved"] / hbm_peak

    # CPU baseline of the headline (rank 0, N == 1): the oracle port on a bounded sample; same seed and
    # samples as the GPU => the two images are also a parity check of the timed configuration
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        import oracle
        a = scene.arrays()
        iview, sw, sh, focal, _, _ = cam.device_record()
        ocam = oracle.make_camera(iview, sw, sh, focal, W, H)
        cores = host_threads()
        cpu_spp = 32
        oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"], ocam,
                      oracle.make_params(seed=1, spp_begin=0, spp_end=1, max_depth=2), rows=(0, 8), nthreads=cores)
        t0 = time.perf_counter()
        acc_o, ids_o, _ = oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"], ocam,
                                        oracle.make_params(seed=1, spp_begin=0, spp_end=cpu_spp, max_depth=C3_DEPTH),
                                        want_ids=True, nthreads=cores)
        dt = time.perf_counter() - t0
        acc_g = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        ids_g = torch.empty((H, W, cpu_spp), dtype=torch.int32, device=dev)
        rctx.render(rctx.render_params(seed=1, spp_begin=0, spp_end=cpu_spp, max_depth=C3_DEPTH, flags=_abi.RENDER_EXACT_PRIMARY), acc_g, ids_g)
        torch.cuda.synchronize()
        ids_equal = bool(np.array_equal(ids_g.cpu().numpy(), ids_o))
        g = acc_g.cpu().numpy().astype(np.float64)[..., :3]
        rel = float(np.sqrt(np.mean((g - acc_o[..., :3]) ** 2)) / np.mean(acc_o[..., :3]))
        assert ids_equal, "bench parity check failed: primary-hit ids differ from the oracle"
        assert rel < 5e-3, f"bench parity check failed: radiance rel RMSE {rel}"  # (the tests hold the 1e-3 bar)
        cpu = {"value": cpu_spp / dt, "unit": "spp/s", "cores": cores, "kind": "port",
               "sample": f"{cpu_spp} spp of the frame ({W * H * cpu_spp} paths, depth {C3_DEPTH}), {dt:.1f} s; same seed on the GPU: "
                         f"primary-hit ids identical ({W * H * cpu_spp} rays), radiance rel RMSE {rel:.2e}"}

    # ================================================================== closest hit on the 1M soup
    closest = None
    ctx = _abi.Context(local)
    if not args.skip_soup:
        closest = soup_leg(args, ctx, torch, _abi, dev, rank, world, local, SOUP_TRIS, barrier, max_over_ranks, hbm_peak, peak_src,
                           profile="soup1m", e2e=True, cpu=(rank == 0 and world == 1 and not args.skip_cpu))
    soup10 = None
    if not args.skip_soup10m:
        soup10 = soup_leg(args, ctx, torch, _abi, dev, rank, world, local, SOUP10_TRIS, barrier, max_over_ranks, hbm_peak, peak_src,
                          profile="soup10m", e2e=False, cpu=False, steps=min(args.steps, 5))
    ctx.close()
    torch.cuda.empty_cache()

    # ================================================================== C5
    c5 = None
    if not args.skip_c5:
        from pyrenderer_b200.mathematics.subdivide import subdivide_scene_arrays
        a = dict(scene.arrays())
        m = a["materials"].copy()
        m[5]["type"], m[5]["ior"], m[5]["two_sided"], m[5]["albedo"] = 3, 1.5, 0, (1.0, 1.0, 1.0)   # ShortBox -> dielectric
        m[6]["type"], m[6]["roughness"], m[6]["albedo"] = 4, 0.0, (0.9, 0.8, 0.6)                  # TallBox -> conductor
        a["materials"] = m
        levels = np.where(np.isin(a["tri_prim"], [5, 6]), 8, 9)   # quads 12 x 4^9, boxes 24 x 4^8 = 4 718 592 triangles
        scene5 = Scene.from_arrays(subdivide_scene_arrays(a, levels))
        nt5 = scene5.arrays()["tris"].shape[0]
        cam5 = Camera(cam.position, cam.looking_at, cam.up, [3840, 2160], fov=cam.fov, aperture=cam.aperture, focal_dist=cam.focal_dist)
        ctx5 = scene5.commit(local)
        if world > 1:
            tracing.init_distributed(scene5, device=local)
        acc5 = torch.zeros((2160, 3840, 4), dtype=torch.float32, device=dev)
        spp5 = max(args.c5_spp, world)

        def frame5(k):
            tracing.render_distributed(scene5, cam5, spp5, max_depth=8, seed=1, spp_begin=k * spp5, device=local, accum=acc5)
        frame5(0)
        ctx5.reset_counters()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        n5 = 3
        for k in range(n5):
            frame5(1 + k)
        g1.record()
        barrier()
        ms5 = max_over_ranks(g0.elapsed_time(g1)) / n5
        c5c = ctx5.counters()
        rays5 = sum_over_ranks((c5c["rays_closest"] + c5c["rays_shadow"]) / n5)
        c5 = {"workload": f"subdivided Cornell box, {nt5} triangles, dielectric + conductor boxes, 3840x2160, max depth 8 (BASELINE configs[4]); "
                          f"{spp5} spp per frame sharded over {world} GPU(s), one all-reduce per frame",
              "triangles": nt5, "bvh": scene5.bvh_stats, "spp_per_frame": spp5, "ms_per_frame": ms5, "spp_per_s": spp5 / (ms5 * 1e-3),
              "mrays_per_s": rays5 / (ms5 * 1e-3) / 1e6,
              "seconds_for_4096_spp_extrapolated": 4096.0 / (spp5 / (ms5 * 1e-3)),
              "mean_radiance": float((acc5[..., :3] / acc5[..., 3:]).mean().item())}
        del acc5
        ctx5.close()

    # ================================================================== C1 (BASELINE configs[0]) next to the CPU port
    c1 = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        import oracle
        a = scene.arrays()
        iview, sw, sh, focal, _, _ = cam.device_record()
        w1 = h1 = 256
        rctx.set_camera(iview, sh * (w1 / h1), sh, focal, w1, h1)
        kw = dict(seed=1, spp_begin=0, spp_end=16, max_depth=5)
        acc1 = torch.zeros((h1, w1, 4), dtype=torch.float32, device=dev)
        rctx.render(rctx.render_params(flags=_abi.RENDER_EXACT_PRIMARY, **kw), acc1)  # warm
        acc1.zero_()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        rctx.render(rctx.render_params(flags=_abi.RENDER_EXACT_PRIMARY, **kw), acc1)
        g1.record()
        torch.cuda.synchronize()
        c1_gpu_ms = g0.elapsed_time(g1)
        ocam = oracle.make_camera(iview, sh * (w1 / h1), sh, focal, w1, h1)
        cores = host_threads()
        t0 = time.perf_counter()
        acc_o = oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"], ocam,
                              oracle.make_params(**kw), nthreads=cores)[0]
        c1_cpu_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()  # the reference's literal setting: joblib n_jobs=4 (main.py:52)
        oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"], ocam,
                      oracle.make_params(**kw), nthreads=4)
        c1_cpu4_ms = (time.perf_counter() - t0) * 1e3
        g = acc1.cpu().numpy().astype(np.float64)[..., :3]
        c1 = {"workload": "cornell-box 256x256, 16 spp, max depth 5 (BASELINE configs[0])",
              "gpu_ms": c1_gpu_ms, "gpu_spp_per_s": 16 / (c1_gpu_ms * 1e-3),
              "cpu_port_ms": c1_cpu_ms, "cpu_port_spp_per_s": 16 / (c1_cpu_ms * 1e-3), "cpu_cores": cores,
              "cpu_port_4_threads_ms": c1_cpu4_ms,
              "rel_rmse_gpu_vs_oracle_equal_seed": float(np.sqrt(np.mean((g - acc_o[..., :3]) ** 2)) / np.mean(acc_o[..., :3]))}

    if rank == 0:
        line = {
            "metric": "cornell-box render spp/s", "value": value, "unit": "spp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": frame_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c3_workload(total), "spp_per_rank": [tracing.shard_samples(total, r, world)[1] - tracing.shard_samples(total, r, world)[0] for r in range(world)],
                       "exact_primary": True, "estimator": "reference (core/tracing.py:116-155)", "rng": "Philox4x32-10 per (pixel, sample, bounce)",
                       "l2": f"inputs larger than L2: every wave (64 spp of every pixel) streams {64 * W * H * 136 / 2**30:.1f} GiB of path state",
                       "api": "pyrenderer_b200.core.tracing.render_distributed -> prt_render_sharded" if world > 1 else
                              "pyrenderer_b200.core.tracing.render_distributed -> prt_render",
                       "bvh": scene.bvh_stats},
            "clocks": clocks,
            "mrays_per_s": rays_frame / (frame_ms * 1e-3) / 1e6,
            "allreduce": {"ms_per_frame": kernel_ms["allreduce"], "bytes": acc_bytes, "library": "NCCL " + str(rctx.comm_info()["nccl_version"]),
                          "share_of_frame": kernel_ms["allreduce"] / frame_ms} if world > 1 else None,
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "spp/s", "h2d_bytes_per_step": acc_bytes, "d2h_bytes_per_step": acc_bytes,
                    "ms_per_step": e2e_ms, "api": "render_distributed with a pinned host accumulation buffer: upload, frame, download, synchronize"},
            "gpu_launches": int(round(launches_frame * args.steps)),
            "gpu_launches_per_frame_this_rank": launches_frame,
            "exact_primary_cost": {"wave_16spp_exact_ms": wave_exact_ms, "wave_16spp_plain_ms": wave_plain_ms,
                                   "overhead_frac": wave_exact_ms / wave_plain_ms - 1.0},
            "roofline": c3_roof,
            "shade_roofline": shade_roof,
            "cpu_baseline": cpu,
            "closest_hit": closest,
            "soup10m": soup10,
            "c5": c5,
            "c1": c1,
        }
        print(json.dumps(line))
    rctx.close()
    if world > 1:
        dist.destroy_process_group()


def soup_leg(args, ctx, torch, _abi, dev, rank, world, local, n_tris, barrier, max_over_ranks, hbm_peak, peak_src,
             profile, e2e, cpu, steps=None):
    """Incoherent closest-hit rays against an n_tris-triangle random soup: BVH build, exact and plain
    traversal (K steps each), roofline, optional host-buffer e2e and CPU baseline / parity check."""
    steps = steps or args.steps
    tris = soup(n_tris)
    tris_d = torch.from_numpy(tris).to(dev)
    build_ms, build_wall = [], []
    for _ in range(5):
        ctx.set_triangles_dev(tris_d, n_tris)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = ctx.build_bvh()
        build_wall.append((time.perf_counter() - t0) * 1e3)
        build_ms.append(st["ms_total"])
    del tris_d
    g = torch.Generator(device=dev)
    g.manual_seed(11 + 1000 * rank)
    batches = []
    for b in range(N_BATCHES):
        r = torch.empty((RAYS_PER_BATCH, 8), dtype=torch.float32, device=dev)
        r[:, 0:3] = torch.rand((RAYS_PER_BATCH, 3), generator=g, device=dev)
        d = torch.randn((RAYS_PER_BATCH, 3), generator=g, device=dev)
        r[:, 4:7] = d / d.norm(dim=1, keepdim=True)
        r[:, 3] = 1e-5
        r[:, 7] = 3.4e38
        batches.append(r)
    hits = torch.empty((RAYS_PER_BATCH, 4), dtype=torch.float32, device=dev)
    out = {"workload": f"soup-{n_tris} closest-hit: {RAYS_PER_BATCH} incoherent rays per step per GPU (origins U[0,1]^3, directions uniform "
                       "on S^2), BVH replicated, no collective (weak)",
           "l2": "inputs larger than L2: 512 MiB of rays + 256 MiB of hits per step, 4 rotating batches",
           "ray_binning": ("on by default: BVH of %.0f MB > 96 MB (rays counting-sorted by origin cell before the traversal, inside the timed step)"
                           if st["n_nodes"] * 64 + n_tris * 40 > (96 << 20) else "off by default: BVH of %.0f MB lives in L2") % ((st["n_nodes"] * 64 + n_tris * 40) / 2**20),
           "bvh": st, "bvh_build_ms_median": float(np.median(build_ms)), "bvh_build_wall_ms_median": float(np.median(build_wall)),
           "bvh_build_mtris_per_s": n_tris / np.median(build_wall) / 1e3, "steps": steps}
    modes = {}
    for mode, flags in (("exact", _abi.TRACE_EXACT), ("fp32", 0)):
        ctx.reset_counters()
        ctx.trace_closest(batches[0], RAYS_PER_BATCH, hits, flags | _abi.TRACE_COUNT)
        c = ctx.counters()
        n_node, n_tri = c["node_visits"] / RAYS_PER_BATCH, c["tri_tests"] / RAYS_PER_BATCH
        for s in range(args.warmup):
            ctx.trace_closest(batches[s % N_BATCHES], RAYS_PER_BATCH, hits, flags)
        ctx.reset_counters()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.profile_begin()
        ev0.record()
        for s in range(steps):
            ctx.trace_closest(batches[(args.warmup + s) % N_BATCHES], RAYS_PER_BATCH, hits, flags)
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
        prof = ctx.profile_end()
        cc = ctx.counters()
        kern_ms = prof["closest"][0] / steps
        # SURVEY 8d per-ray figure with the real fetch sizes: 32 B ray in + 16 B hit out, 64 B per (4-wide) node
        # record visited, 40 B per triangle tested (32 + 8 byte leaf-order record)
        bytes_per_ray = 48.0 + 64.0 * n_node + 40.0 * n_tri
        achieved = bytes_per_ray * RAYS_PER_BATCH / (kern_ms * 1e-3) / 1e9
        ncu = ncu_profile(f"{profile}_{mode}")
        roof = {"kernel": f"prt::trace_persistent_kernel<CLOSEST,{'EXACT' if flags else 'plain'}> (persist.cuh)",
                "achieved": achieved, "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / hbm_peak,
                "frac_meaning": "SURVEY 8(d) nominal: algorithmic bytes / kernel time / HBM peak (the BVH may be L2-resident: see dram_frac)",
                "bytes_per_ray": bytes_per_ray, "n_node": n_node, "n_tri": n_tri, "kernel_ms": kern_ms,
                "fixup_ms": prof["exact_fixup"][0] / steps, "binning_ms": prof["other"][0] / steps, "traffic": None, "bound": "see ncu profile (not captured)", "ncu": ncu}
        if ncu and not ncu.get("stale"):
            roof["traffic"] = ncu["dram_bytes"]
            roof["dram_frac"] = ncu["dram_bytes"] / (ncu["duration_ns"] * 1e-9) / 1e9 / hbm_peak
            l1, l2, dr, iss = (ncu.get("l1_data_pipe_pct", 0.0), ncu.get("l2_throughput_pct", 0.0), ncu.get("dram_throughput_pct", 0.0),
                               ncu.get("issue_active_pct", 0.0))
            top = max((l1, "l1-data-pipe"), (l2, "l2"), (dr, "hbm"), (iss, "issue"))
            roof["bound"] = f"{top[1]} ({top[0]:.0f} % of peak under ncu; L1 data pipe {l1:.0f} %, L2 {l2:.0f} %, DRAM {dr:.0f} %, issue {iss:.0f} %)"
            roof["l1_data_pipe_frac"], roof["simt_lanes_per_warp"] = l1 / 100.0, ncu.get("lanes_per_instruction")
        modes[mode] = {"value": world * RAYS_PER_BATCH / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms,
                       "gpu_launches_per_step": sum(v[1] for v in prof.values()) / steps,
                       "fp64_replayed_rays_per_step": cc["flagged_rays"] / steps, "fp64_in_place_decisions_per_ray": c["f64_decisions"] / RAYS_PER_BATCH,
                       "hit_fraction": float((hits[:, 3].view(torch.int32) >= 0).float().mean().item()), "roofline": roof}
    out["exact"], out["fp32"] = modes["exact"], modes["fp32"]
    out["value"], out["unit"], out["metric"] = modes["exact"]["value"], "Mrays/s", "closest-hit Mrays/s, ids bit-exact (PRT_TRACE_EXACT)"

    if e2e:  # host rays in, host hits out, exact mode, through the C-ABI host entry point
        e2e_steps = max(2, min(steps, 4))
        host_batch = batches[0].cpu().numpy()
        pinned = torch.from_numpy(host_batch).pin_memory().numpy()
        hits_host = torch.empty((RAYS_PER_BATCH, 4), dtype=torch.float32).pin_memory().numpy().view(_abi.HIT_DTYPE).reshape(-1)
        res = {}
        for mode, flags in (("exact", _abi.TRACE_EXACT), ("fp32", 0)):
            ctx.trace_closest_host(pinned, flags, out=hits_host)  # warm: staging buffers, copy streams
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                ctx.trace_closest_host(pinned, flags, out=hits_host)
            barrier()
            ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
            res[mode] = {"value": world * RAYS_PER_BATCH / (ms * 1e-3) / 1e6, "ms_per_step": ms}
        out["e2e"] = {"value": res["exact"]["value"], "unit": "Mrays/s", "h2d_bytes_per_step": RAYS_PER_BATCH * 32,
                      "d2h_bytes_per_step": RAYS_PER_BATCH * 16, "ms_per_step": res["exact"]["ms_per_step"], "fp32": res["fp32"],
                      "api": "prt_trace_closest_host (pinned host rays -> host hits), 2^21-ray chunks pipelined over 4 streams",
                      "pcie_GBps_per_gpu": (RAYS_PER_BATCH * 48) / (res["exact"]["ms_per_step"] * 1e-3) / 1e9}
        del pinned, host_batch, hits_host
    if cpu:  # CPU port on a bounded sample; doubles as the parity check of the TIMED (exact) kernel
        import oracle
        cores = host_threads()
        n = max(256, 640 * cores)  # ~10-15 s of brute force with every host thread busy
        sample = host_rays(n, seed=99)
        oracle.closest_hit(tris[:1000], sample[:8], nthreads=cores)
        t0 = time.perf_counter()
        ids_o, _, _, _ = oracle.closest_hit(tris, sample, nthreads=cores)
        dt = time.perf_counter() - t0
        sd = torch.from_numpy(sample).to(dev)
        hd = torch.empty((n, 4), dtype=torch.float32, device=dev)
        ctx.trace_closest(sd, n, hd, _abi.TRACE_EXACT)
        torch.cuda.synchronize()
        ids_x = hd[:, 3].view(torch.int32).cpu().numpy()
        ctx.trace_closest(sd, n, hd, 0)
        torch.cuda.synchronize()
        n_bad = int(np.sum(hd[:, 3].view(torch.int32).cpu().numpy() != ids_o))
        assert np.array_equal(ids_x, ids_o), "bench parity check failed: exact-mode ids differ from the oracle"
        out["cpu_baseline"] = {"value": n / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                               "sample": f"{n} rays x {n_tris} triangles brute force (reference algorithm), {dt:.1f} s; ids == the timed exact "
                                         f"kernel on all {n} rays; plain FP32 kernel: {n_bad} mismatches"}
    del batches, hits
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
