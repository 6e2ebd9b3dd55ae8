"""BASELINE configs 4 (10M soup) and 5 (subdivided Cornell, 4K) on one B200.
usage: python profiles/prof_configs.py c4 | c5 [spp per render] [Mi paths per wave]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import soup
from pyrenderer_b200 import _abi

dev = torch.device("cuda", 0)
mode = sys.argv[1]
ctx = _abi.Context(0)
if mode == "c4":
    for n in (1_000_000, 10_000_000):
        tris = torch.from_numpy(soup(n)).to(dev)
        ms = []
        for _ in range(5):
            ctx.set_triangles_dev(tris, n)
            st = ctx.build_bvh()
            ms.append(st["ms_total"])
        N = 1 << 24
        g = torch.Generator(device=dev); g.manual_seed(11)
        best = 1e9
        hits = torch.empty((N, 4), dtype=torch.float32, device=dev)
        for b in range(4):
            r = torch.empty((N, 8), dtype=torch.float32, device=dev)
            r[:, 0:3] = torch.rand((N, 3), generator=g, device=dev)
            d = torch.randn((N, 3), generator=g, device=dev)
            r[:, 4:7] = d / d.norm(dim=1, keepdim=True); r[:, 3] = 1e-5; r[:, 7] = 3.4e38
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.trace_closest(r, N, hits, 0); e1.record(); torch.cuda.synchronize()
            if b > 0:
                best = min(best, e0.elapsed_time(e1))
        ctx.reset_counters(); ctx.trace_closest(r, N, hits, _abi.TRACE_COUNT); c = ctx.counters()
        print(json.dumps({"config": f"C4 soup {n}", "build_ms_median": float(np.median(ms)), "build_mtris_per_s": n / np.median(ms) / 1e3,
                          "bvh": st, "closest_mrays_per_s": N / best / 1e3, "n_node": c["node_visits"] / N, "n_tri": c["tri_tests"] / N}), flush=True)
        del tris, r
else:
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    from pyrenderer_b200.main import DEFAULT_SCENE
    from pyrenderer_b200.mathematics.subdivide import subdivide_scene_arrays
    scene, cam = read_file(DEFAULT_SCENE)
    a = scene.arrays()
    a["materials"] = a["materials"].copy()
    m = a["materials"]
    m[5]["type"], m[5]["ior"], m[5]["two_sided"], m[5]["albedo"] = 3, 1.5, 0, (1.0, 1.0, 1.0)   # ShortBox -> dielectric
    m[6]["type"], m[6]["roughness"], m[6]["albedo"] = 4, 0.0, (0.9, 0.8, 0.6)                  # TallBox -> conductor
    levels = np.where(np.isin(a["tri_prim"], [5, 6]), 8, 9)   # quads 12 x 4^9, boxes 24 x 4^8 = 4 718 592 triangles
    b = subdivide_scene_arrays(a, levels)
    nt = b["tris"].shape[0]
    ctx.set_triangles(b["tris"], b["normals"], b["tri_material"], b["materials"], b["light_tris"])
    st = ctx.build_bvh()
    iview, sw, sh, focal, _, _ = cam.device_record()
    W, H = 3840, 2160
    ctx.set_camera(iview, sh * (W / H), sh, focal, W, H)
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    if len(sys.argv) > 3:
        ctx.set_wave_paths(int(sys.argv[3]) << 20)  # Mi paths per wave
    res = []
    for rep in range(3):
        ctx.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.render(ctx.render_params(seed=1, spp_begin=rep * spp, spp_end=(rep + 1) * spp, max_depth=8), acc); e1.record()
        torch.cuda.synchronize()
        c = ctx.counters()
        res.append((e0.elapsed_time(e1), c["rays_closest"] + c["rays_shadow"]))
    ms, rays = min(res)
    img = (acc[..., :3] / acc[..., 3:]).mean().item()
    print(json.dumps({"config": f"C5 subdivided Cornell {nt} tris, {W}x{H}, depth 8, dielectric+conductor", "bvh": st,
                      "ms_per_spp": ms / spp, "spp_per_s": spp / ms * 1e3, "mrays_per_s": rays / ms / 1e3,
                      "light_tris": int(b["light_tris"].shape[0]), "mean_radiance": img}))
