"""rotations = 0 vs 1 after the SAH treelets: build time and traversal on soup-10M, the C5 scene and the Cornell render."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import soup, load_cornell
from pyrenderer_b200 import _abi
from pyrenderer_b200.mathematics.subdivide import subdivide_scene_arrays
dev = torch.device("cuda", 0)
ctx = _abi.Context(0)
N = 1 << 24
g = torch.Generator(device=dev); g.manual_seed(11)
r = torch.empty((N, 8), dtype=torch.float32, device=dev)
r[:, 0:3] = torch.rand((N, 3), generator=g, device=dev)
d = torch.randn((N, 3), generator=g, device=dev)
r[:, 4:7] = d / d.norm(dim=1, keepdim=True); r[:, 3] = 1e-5; r[:, 7] = 3.4e38
hits = torch.empty((N, 4), dtype=torch.float32, device=dev)
tris = torch.from_numpy(soup(10_000_000)).to(dev)
for rot in (0, 1):
    ctx.set_triangles_dev(tris, 10_000_000)
    ms = [ctx.build_bvh(rotations=rot)["ms_wall"] for _ in range(4)]
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.trace_closest(r, N, hits, 0); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ctx.reset_counters(); ctx.trace_closest(r, N, hits, _abi.TRACE_COUNT); c = ctx.counters()
    print(f"soup-10M rotations {rot}: build wall {np.median(ms[1:]):.2f} ms, {N / best / 1e3:.0f} Mrays/s, N_node {c['node_visits'] / N:.2f}", flush=True)
del tris, r, hits
scene, cam = load_cornell()
a = dict(scene.arrays())
m = a["materials"].copy()
m[5]["type"], m[5]["ior"], m[5]["two_sided"], m[5]["albedo"] = 3, 1.5, 0, (1.0, 1.0, 1.0)
m[6]["type"], m[6]["roughness"], m[6]["albedo"] = 4, 0.0, (0.9, 0.8, 0.6)
a["materials"] = m
b = subdivide_scene_arrays(a, np.where(np.isin(a["tri_prim"], [5, 6]), 8, 9))
iview, sw, sh, focal, _, _ = cam.device_record()
for name, arr, (W, H), spp in (("C5", b, (3840, 2160), 4), ("cornell", scene.arrays(), (1024, 1024), 16)):
    for rot in (0, 1):
        ctx.set_triangles(arr["tris"], arr["normals"], arr["tri_material"], arr["materials"], arr["light_tris"])
        ms = [ctx.build_bvh(rotations=rot)["ms_wall"] for _ in range(3)]
        ctx.set_camera(iview, sh * (W / H), sh, focal, W, H)
        acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        best = 1e9
        for k in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.render(ctx.render_params(seed=1, spp_begin=k * spp, spp_end=(k + 1) * spp, max_depth=8, flags=1), acc); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name} rotations {rot}: build wall {np.median(ms[1:]):.2f} ms, {spp} spp in {best:.2f} ms", flush=True)
        del acc
