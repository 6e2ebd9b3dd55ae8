#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path-tracing core.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[3] + north_star target ">= 2 Grays/s closest-hit on a
1M-triangle scene per B200"): incoherent closest-hit rays against a 1 000 000-triangle random
soup.  One STEP = one batch of 2^24 rays through prt_trace_closest (one launch of the traversal
kernel).  `value` = Mrays/s summed over all GPUs with rays resident in HBM; `e2e` = the same
through prt_trace_closest_host (pinned host rays in, host hits out, copies inside the timed
region).  N > 1: every rank traces its own ray batches over a replicated BVH (no data-path
collective -> "weak").

Second leg in the same line (`render`): BASELINE.json configs[2], Cornell box 1024x1024, max
depth 8, sample-sharded -- every step each rank renders `spp_per_step` samples of every pixel
and the fp32 accumulation buffers are summed with ONE NCCL all-reduce per step (N > 1).

`--impl reference` times the reference's CPU algorithm (the oracle port: brute-force
Moller-Trumbore over all triangles, mathematics/intersection.py) on a bounded sample of the
same workload with all host threads.
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
RAYS_PER_BATCH = 1 << 24
N_BATCHES = 4
HBM_FALLBACK_GBS = 6650.0


NCU_DRAM_BYTES_PER_LAUNCH = 3.600135e9 + 0.378812e9  # gpurun r41 capture


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


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 100 ms from `start()` on; `summary()`
    keeps the samples taken between `mark_begin()` and `mark_end()` (the timed region); if the
    region was shorter than one sampling period it falls back to the samples since start()
    (warm-up + timed region, all under load)."""
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


def run_reference(args):
    """CPU arm: oracle port of the reference's brute-force closest hit, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    tris = soup(SOUP_TRIS)
    cores = oracle.num_threads()
    n = max(64, 64 * cores)  # rays per step (~1 s of brute force per step with every thread busy)
    times = []
    for s in range(args.warmup + args.steps):
        rays = host_rays(n, seed=1000 + s)
        t0 = time.perf_counter()
        oracle.closest_hit(tris, rays)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    v = n / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": "closest-hit Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"soup-{SOUP_TRIS} closest-hit, incoherent rays", "rays_per_step": n,
                       "note": "reference algorithm = test every triangle (core/scene.py:66-73); its BVH does not run at HEAD"},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"{n} rays x {SOUP_TRIS} triangles per step"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--render-spp", type=int, default=16, help="Cornell samples per pixel per step per GPU")
    ap.add_argument("--skip-render", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pyrenderer_b200 import _abi
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    from pyrenderer_b200.main import DEFAULT_SCENE

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

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = _abi.Context(local)

    # ------------------------------------------------------------------ soup leg
    tris = soup(SOUP_TRIS)
    tris_d = torch.from_numpy(tris).to(dev)
    build_ms = []
    for _ in range(5):
        ctx.set_triangles_dev(tris_d, SOUP_TRIS)
        st = ctx.build_bvh()
        build_ms.append(st["ms_total"])
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
    # counters (N_node, N_tri per ray) from the instrumented twin, outside the timed region
    ctx.reset_counters()
    ctx.trace_closest(batches[0], RAYS_PER_BATCH, hits, _abi.TRACE_COUNT)
    c = ctx.counters()
    n_node = c["node_visits"] / RAYS_PER_BATCH
    n_tri = c["tri_tests"] / RAYS_PER_BATCH
    # SURVEY 8d per-ray figure with the real fetch sizes: 32 B ray in + 16 B hit out, 64 B per
    # (4-wide) node record visited, 40 B per triangle tested (32 + 8 byte leaf-order record)
    bytes_per_ray = 48.0 + 64.0 * n_node + 40.0 * n_tri
    hit_frac = float((hits[:, 3].view(torch.int32) >= 0).float().mean().item())

    clk = ClockSampler(local).start()
    for s in range(args.warmup):
        ctx.trace_closest(batches[s % N_BATCHES], RAYS_PER_BATCH, hits, 0)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    clk.mark_begin()
    ev[0].record()
    for s in range(args.steps):
        ctx.trace_closest(batches[(args.warmup + s) % N_BATCHES], RAYS_PER_BATCH, hits, 0)
        ev[s + 1].record()
    barrier()
    clk.mark_end()
    clk.stop()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    kern_ms = float(np.mean([ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]))
    ms_per_step = total_ms / args.steps
    value = world * RAYS_PER_BATCH / (ms_per_step * 1e-3) / 1e6
    clocks = clk.summary()
    achieved = bytes_per_ray * RAYS_PER_BATCH / (kern_ms * 1e-3) / 1e9

    # e2e: host rays in, host hits out, through the C-ABI host entry point
    e2e_steps = max(2, min(args.steps, 4))
    host_batch = batches[0].cpu().numpy()
    pinned = torch.from_numpy(host_batch).pin_memory().numpy()
    hits_host = torch.empty((RAYS_PER_BATCH, 4), dtype=torch.float32).pin_memory().numpy().view(_abi.HIT_DTYPE).reshape(-1)
    ctx.trace_closest_host(pinned, 0, out=hits_host)  # warm: staging buffers, copy streams
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.trace_closest_host(pinned, 0, out=hits_host)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_value = world * RAYS_PER_BATCH / (e2e_ms * 1e-3) / 1e6
    del pinned, host_batch, hits_host
    for b in batches[1:]:
        del b
    batches = batches[:1]
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ render leg
    render = None
    launches_render = 0
    if not args.skip_render:
        scene, cam = read_file(DEFAULT_SCENE)
        a = scene.arrays()
        rctx = _abi.Context(local)
        rctx.set_triangles(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"])
        rctx.build_bvh()
        iview, sw, sh, focal, W, H = cam.device_record()
        rctx.set_camera(iview, sw, sh, focal, W, H)
        depth, spp = 8, args.render_spp
        accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)

        def render_step(step):
            s0 = (step * world + rank) * spp  # disjoint Philox sample ranges per rank and step
            rctx.render(rctx.render_params(seed=1, spp_begin=s0, spp_end=s0 + spp, max_depth=depth), accum)
            if world > 1:
                dist.all_reduce(accum, op=dist.ReduceOp.SUM)

        for s in range(args.warmup):
            render_step(s)
        rctx.reset_counters()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.steps):
            render_step(args.warmup + s)
        e1.record()
        barrier()
        r_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        rc = rctx.counters()
        rays_step = (rc["rays_closest"] + rc["rays_shadow"]) / args.steps
        t = torch.tensor([rays_step], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        # e2e: host accumulation buffer through prt_render_host (upload + render + download)
        host_acc = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory().numpy()
        rctx.render_host(rctx.render_params(seed=1, spp_begin=0, spp_end=1, max_depth=depth), host_acc)  # warm
        barrier()
        t0 = time.perf_counter()
        for k in range(2):
            rctx.render_host(rctx.render_params(seed=1, spp_begin=k * spp, spp_end=(k + 1) * spp, max_depth=depth), host_acc)
        r_e2e_ms = (time.perf_counter() - t0) * 1e3 / 2
        waves = (spp * W * H + (16 << 20) - 1) // (16 << 20)
        launches_render = waves * (2 + 4 * depth)
        render = {"workload": f"cornell-box {W}x{H}, max depth {depth}, {spp} spp/step/GPU, sample-sharded"
                              + (", 1 NCCL all-reduce of the fp32 accum per step" if world > 1 else ""),
                  "ms_per_step": r_ms, "mrays_per_s": float(t.item()) / (r_ms * 1e-3) / 1e6,
                  "spp_per_s": world * spp / (r_ms * 1e-3),
                  "rays_closest_per_step": rc["rays_closest"] / args.steps,
                  "rays_shadow_per_step": rc["rays_shadow"] / args.steps,
                  "e2e_host_ms_per_step": r_e2e_ms, "e2e_spp_per_s": spp / (r_e2e_ms * 1e-3),
                  "gpu_launches_per_step": launches_render}
        # BASELINE configs[0] -- the reference's own CPU-runnable case: Cornell 256x256, 16 spp, depth 5.
        # GPU through the same C-ABI call; CPU = oracle port of main.py's loop (all host threads);
        # same seed => same paths, so the two images are also a parity spot check.
        if rank == 0 and world == 1 and not args.skip_cpu:
            import oracle
            w1 = h1 = 256
            rctx.set_camera(iview, sh * (w1 / h1), sh, focal, w1, h1)
            kw = dict(seed=1, spp_begin=0, spp_end=16, max_depth=5)
            acc1 = torch.zeros((h1, w1, 4), dtype=torch.float32, device=dev)
            rctx.render(rctx.render_params(**kw), acc1)  # warm
            acc1.zero_()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            rctx.render(rctx.render_params(**kw), acc1)
            g1.record()
            torch.cuda.synchronize()
            c1_gpu_ms = g0.elapsed_time(g1)
            ocam = oracle.make_camera(iview, sh * (w1 / h1), sh, focal, w1, h1)
            t0 = time.perf_counter()
            acc_o = oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"], ocam,
                                  oracle.make_params(**kw))[0]
            c1_cpu_ms = (time.perf_counter() - t0) * 1e3
            all_threads = oracle.num_threads()
            t0 = time.perf_counter()  # the reference's literal setting: joblib n_jobs=4 (main.py:52)
            oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"], ocam,
                          oracle.make_params(**kw), nthreads=4)
            c1_cpu4_ms = (time.perf_counter() - t0) * 1e3
            # omp_set_num_threads is sticky: hand all threads back before the soup baseline below
            oracle.closest_hit(a["tris"][:1], np.zeros((1, 8), np.float32), nthreads=all_threads)
            g = acc1.cpu().numpy().astype(np.float64)[..., :3]
            c1 = {"workload": "cornell-box 256x256, 16 spp, max depth 5 (BASELINE configs[0])",
                  "gpu_ms": c1_gpu_ms, "gpu_spp_per_s": 16 / (c1_gpu_ms * 1e-3),
                  "cpu_port_ms": c1_cpu_ms, "cpu_port_spp_per_s": 16 / (c1_cpu_ms * 1e-3), "cpu_cores": all_threads,
                  "cpu_port_4_threads_ms": c1_cpu4_ms,
                  "rel_rmse_gpu_vs_oracle_equal_seed": float(np.sqrt(np.mean((g - acc_o[..., :3]) ** 2)) / np.mean(acc_o[..., :3]))}
            render["c1"] = c1
        rctx.close()

    # ------------------------------------------------------------------ CPU baseline (rank 0, N == 1)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        import oracle
        cores = oracle.num_threads()
        n = max(256, 640 * cores)  # ~10-15 s of brute force with every host thread busy
        sample = host_rays(n, seed=99)
        oracle.closest_hit(tris[:1000], sample[:8])
        t0 = time.perf_counter()
        ids_o, _, _, _ = oracle.closest_hit(tris, sample)
        dt = time.perf_counter() - t0
        # the sample doubles as a parity spot check of the timed configuration
        h = ctx.trace_closest_host(sample, _abi.TRACE_EXACT)
        assert np.array_equal(h["tri"], ids_o), "bench parity spot check failed"
        cpu = {"value": n / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
               "sample": f"{n} rays x {SOUP_TRIS} triangles brute force (reference algorithm), {dt:.1f} s; ids == GPU exact mode"}

    if rank == 0:
        line = {
            "metric": "closest-hit Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"soup-{SOUP_TRIS} closest-hit: {RAYS_PER_BATCH} incoherent rays per step per GPU "
                                   "(origins U[0,1]^3, directions uniform on S^2), BVH replicated",
                       "l2": "inputs larger than L2: 512 MiB of rays + 256 MiB of hits per step, 4 rotating batches",
                       "bvh_build_ms_median": float(np.median(build_ms)), "bvh": st, "hit_fraction": hit_frac},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": RAYS_PER_BATCH * 32,
                    "d2h_bytes_per_step": RAYS_PER_BATCH * 16, "ms_per_step": e2e_ms,
                    "api": "prt_trace_closest_host (pinned host rays -> host hits)"},
            "gpu_launches": args.steps * 1,  # timed region of `value`: one trace_persistent_kernel per step (render leg: see render.gpu_launches_per_step)
            "roofline": {"bound": "hbm", "kernel": "prt::trace_persistent_kernel<CLOSEST> (traverse.cu / persist.cuh)", "achieved": achieved,
                         "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / hbm_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE 2^24-ray launch of this kernel
                         # (ncu --set full, profiles/r1_trace_persistent_final_ncu.txt): 12x below the
                         # algorithmic bytes because nodes + triangles (70 MB) live in the 126 MB L2
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH if RAYS_PER_BATCH == (1 << 24) else None,
                         "traffic_source": "profiles/r1_trace_persistent_final_ncu.txt",
                         "bytes_per_ray": bytes_per_ray, "n_node": n_node, "n_tri": n_tri,
                         "kernel_ms": kern_ms,
                         # what actually bounds the kernel (same capture): the L1 data pipe moves one
                         # 32-byte sector per cycle per SM for divergent lanes
                         "l1_data_pipe_frac": 0.82, "simt_lanes_per_warp": 18.2, "issue_slots_busy": 0.68},
            "cpu_baseline": cpu,
            "render": render,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
