"""Exact (PRT_TRACE_EXACT) vs plain FP32 closest hit on the persistent kernel: Mrays/s and flagged fraction.
soup-1M, 2^24 incoherent rays per launch (the bench step), and Cornell 1024^2 jittered primaries x 16 spp.
usage: python profiles/prof_exact.py"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import soup
from pyrenderer_b200 import _abi

dev = torch.device("cuda", 0)
ctx = _abi.Context(0)

def timed(rays, n, hits, flags, reps=4):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.trace_closest(rays, n, hits, flags); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def report(label, rays, n):
    hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    hx = torch.empty((n, 4), dtype=torch.float32, device=dev)
    ms_f = timed(rays, n, hits, 0)
    ctx.trace_closest(rays, n, hx, _abi.TRACE_EXACT)  # warm: the exact mode's flag list is allocated on first use
    torch.cuda.synchronize()
    ctx.reset_counters()
    ms_x = timed(rays, n, hx, _abi.TRACE_EXACT)
    flagged = ctx.counters()["flagged_rays"] / 4
    ctx.reset_counters(); ctx.trace_closest(rays, n, hx, _abi.TRACE_EXACT | _abi.TRACE_COUNT); c = ctx.counters()
    ctx.reset_counters(); ctx.trace_closest(rays, n, hits, _abi.TRACE_COUNT); c0 = ctx.counters()
    mism = float((hits[:, 3].view(torch.int32) != hx[:, 3].view(torch.int32)).float().mean().item())
    print(json.dumps({"case": label, "rays": n, "fp32_ms": ms_f, "fp32_mrays_s": n / ms_f / 1e3, "exact_ms": ms_x,
                      "exact_mrays_s": n / ms_x / 1e3, "flagged_fraction": flagged / n, "n_node_exact": c["node_visits"] / n, "n_tri_exact": c["tri_tests"] / n,
                      "n_node_fp32": c0["node_visits"] / n, "n_tri_fp32": c0["tri_tests"] / n, "fp32_vs_exact_id_mismatch": mism}), flush=True)

N = 1 << 24
ctx.set_triangles_dev(torch.from_numpy(soup(1_000_000)).to(dev), 1_000_000)
ctx.build_bvh()
g = torch.Generator(device=dev); g.manual_seed(11)
r = torch.empty((N, 8), dtype=torch.float32, device=dev)
r[:, 0:3] = torch.rand((N, 3), generator=g, device=dev)
d = torch.randn((N, 3), generator=g, device=dev)
r[:, 4:7] = d / d.norm(dim=1, keepdim=True); r[:, 3] = 1e-5; r[:, 7] = 3.4e38
report("soup-1M", r, N)
del r
if len(sys.argv) > 1 and sys.argv[1] == "soup":
    sys.exit(0)
from pyrenderer_b200.io_utils.read_tungsten import read_file
from pyrenderer_b200.main import DEFAULT_SCENE
scene, cam = read_file(DEFAULT_SCENE)
a = scene.arrays()
ctx.set_triangles(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"])
ctx.build_bvh()
iview, sw, sh, focal, W, H = cam.device_record()
ctx.set_camera(iview, sw, sh, focal, W, H)
rays = torch.empty((H * W * 16, 8), dtype=torch.float32, device=dev)
ctx.generate_rays(rays, seed=1, s0=0, s1=16, jitter=True)
report("cornell-1024x1024x16spp primaries", rays, H * W * 16)
