"""Tuning sweep on soup-1M: BVH options and persistent-traversal knobs -> Mrays/s, N_node, N_tri.
usage: python profiles/sweep.py bvh | knobs"""
import os, subprocess, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import soup
N = 1 << 22

def rays(dev):
    g = torch.Generator(device=dev); g.manual_seed(11)
    r = torch.empty((N, 8), dtype=torch.float32, device=dev)
    r[:, 0:3] = torch.rand((N, 3), generator=g, device=dev)
    d = torch.randn((N, 3), generator=g, device=dev)
    r[:, 4:7] = d / d.norm(dim=1, keepdim=True)
    r[:, 3] = 1e-5; r[:, 7] = 3.4e38
    return r

def measure(ctx, r, hits, reps=3):
    from pyrenderer_b200 import _abi
    ctx.reset_counters(); ctx.trace_closest(r, N, hits, _abi.TRACE_COUNT); c = ctx.counters()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.trace_closest(r, N, hits, 0); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    if c["warp_iters"]:
        print(f"    util: iters/warp-ray {c['warp_iters'] * 32 / N:.1f}  node lanes/iter {c['node_lane_iters'] / c['warp_iters']:.1f}  "
              f"leaf phases/iter {c['leaf_phases'] / c['warp_iters']:.3f}  lanes/leaf phase {c['leaf_lane_phases'] / max(1, c['leaf_phases']):.1f}")
    return N / best / 1e3, c["node_visits"] / N, c["tri_tests"] / N

def main():
    from pyrenderer_b200 import _abi
    dev = torch.device("cuda", 0)
    tris = torch.from_numpy(soup(1_000_000)).to(dev)
    r = rays(dev); hits = torch.empty((N, 4), dtype=torch.float32, device=dev)
    mode = sys.argv[1] if len(sys.argv) > 1 else "bvh"
    if mode == "bvh":
        ctx = _abi.Context(0)
        for leaf in (1, 2, 4, 7):
            for cn, ct in ((1.0, 1.0), (1.0, 0.5), (1.0, 2.0)):
                for rot in (0, 1):
                    ctx.set_triangles_dev(tris, 1_000_000)
                    st = ctx.build_bvh(max_leaf_tris=leaf, cost_node=cn, cost_tri=ct, rotations=rot)
                    m, nn, nt = measure(ctx, r, hits)
                    print(f"leaf {leaf} cn {cn} ct {ct} rot {rot}: {m:8.1f} Mrays/s  N_node {nn:6.1f} N_tri {nt:5.1f} nodes {st['n_nodes']} depth {st['depth']} sah {st['sah_cost']:.1f} build {st['ms_total']:.2f} ms", flush=True)
    elif mode == "rot":
        ctx = _abi.Context(0)
        for rot in (0, 1, 2, 3, 4, 6):
            ctx.set_triangles_dev(tris, 1_000_000)
            st = ctx.build_bvh(max_leaf_tris=1, rotations=rot)
            m, nn, nt = measure(ctx, r, hits)
            print(f"rotation passes {rot}: {m:8.1f} Mrays/s  N_node {nn:6.2f} N_tri {nt:5.2f} sah {st['sah_cost']:.1f} build {st['ms_total']:.2f} ms (refit {st['ms_refit']:.2f})", flush=True)
    elif mode == "treelet":
        ctx = _abi.Context(0)
        for tl in (0, 1):
            for rot in (0, 1):
                ctx.set_triangles_dev(tris, 1_000_000)
                st = ctx.build_bvh(treelets=tl, rotations=rot)
                ms = []
                for _ in range(4):
                    ms.append(ctx.build_bvh(treelets=tl, rotations=rot)["ms_total"])
                m, nn, nt = measure(ctx, r, hits)
                print(f"treelets {tl} rotations {rot}: {m:8.1f} Mrays/s  N_node {nn:6.2f} N_tri {nt:5.2f} nodes {st['n_nodes']} depth {st['depth']} sah {st['sah_cost']:.1f} build {np.median(ms):.2f} ms (hierarchy+treelets {st['ms_hierarchy']:.2f})", flush=True)
    elif mode == "any":
        ctx = _abi.Context(0)
        ctx.set_triangles_dev(tris, 1_000_000); ctx.build_bvh()
        occ = torch.empty(N, dtype=torch.uint8, device=dev)
        for tmax in (0.05, 0.2, 3.4e38):
            r[:, 7] = tmax
            best = 1e9
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ctx.trace_any(r, N, occ, 0); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"any-hit soup-1M tmax {tmax:g}: {N / best / 1e3:8.1f} Mrays/s  occluded {occ.float().mean().item():.3f}", flush=True)
    elif mode == "util":
        ctx = _abi.Context(0)
        ctx.set_triangles_dev(tris, 1_000_000)
        print(ctx.build_bvh(max_leaf_tris=1))
        print(measure(ctx, r, hits))
    elif mode == "render_wave":
        from pyrenderer_b200.io_utils.read_tungsten import read_file
        from pyrenderer_b200.main import DEFAULT_SCENE
        scene, cam = read_file(DEFAULT_SCENE)
        a = scene.arrays()
        for leaf in (1, 2, 4):
            for wave in (1 << 20, 4 << 20, 16 << 20):
                ctx = _abi.Context(0)
                ctx.set_triangles(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"])
                st = ctx.build_bvh(max_leaf_tris=leaf)
                ctx.set_wave_paths(wave)
                iview, sw, sh, focal, W, H = cam.device_record()
                ctx.set_camera(iview, sw, sh, focal, W, H)
                acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
                best = 1e9
                for rep in range(3):
                    ctx.reset_counters()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); ctx.render(ctx.render_params(seed=1, spp_begin=16 * rep, spp_end=16 * rep + 16, max_depth=8), acc); e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                c = ctx.counters()
                print(f"render leaf {leaf} nodes {st['n_nodes']} wave {wave >> 20}M: {best:7.2f} ms / 16 spp  {(c['rays_closest'] + c['rays_shadow']) / best / 1e3:8.1f} Mrays/s", flush=True)
                ctx.close()
    elif mode in ("render", "render1"):
        from pyrenderer_b200.io_utils.read_tungsten import read_file
        from pyrenderer_b200.main import DEFAULT_SCENE
        scene, cam = read_file(DEFAULT_SCENE)
        a = scene.arrays()
        for ri, lb in (((6, 8), (16, 4), (10, 33), (16, 33), (16, 16), (24, 33), (16, 8), (12, 12), (20, 20)) if mode == "render" else ((16, 8),)):
            os.environ["PRT_REFILL_IDLE"] = str(ri); os.environ["PRT_LEAF_BATCH"] = str(lb)
            ctx = _abi.Context(0)
            ctx.set_triangles(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"])
            ctx.build_bvh()
            iview, sw, sh, focal, W, H = cam.device_record()
            ctx.set_camera(iview, sw, sh, focal, W, H)
            acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
            best = 1e9
            for rep in range(3):
                ctx.reset_counters()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ctx.render(ctx.render_params(seed=1, spp_begin=8 * rep, spp_end=8 * rep + 8, max_depth=8), acc); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            c = ctx.counters()
            print(f"render refill_idle {ri:2d} leaf_batch {lb:2d}: {best:7.2f} ms / 8 spp  {(c['rays_closest'] + c['rays_shadow']) / best / 1e3:8.1f} Mrays/s", flush=True)
            ctx.close()
    else:
        for ri in (4, 6, 8, 12):
            for lb in (4, 8, 12, 16, 20, 24):
                os.environ["PRT_REFILL_IDLE"] = str(ri); os.environ["PRT_LEAF_BATCH"] = str(lb)
                ctx = _abi.Context(0)
                ctx.set_triangles_dev(tris, 1_000_000); ctx.build_bvh()
                m, nn, nt = measure(ctx, r, hits)
                print(f"refill_idle {ri:2d} leaf_batch {lb:2d}: {m:8.1f} Mrays/s", flush=True)
                ctx.close()

main()
