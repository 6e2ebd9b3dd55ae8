"""Small driver for ncu: Cornell 1024x1024, depth 8, 16 spp waves (the bench's headline frame is 64 of them).
Prints the number of rays of every bounce (closest / shadow) so that a captured launch can be normalised per ray.
usage: python profiles/prof_render.py [n_waves] [classes] [bvh option=value ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_cornell  # noqa: E402
from pyrenderer_b200 import _abi  # noqa: E402

n_waves = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scene, cam = load_cornell()
bvh_kw = {a.split("=")[0]: float(a.split("=")[1]) for a in sys.argv[3:] if "=" in a}  # e.g. max_leaf_tris=7 cost_tri=1
SPW = int(bvh_kw.pop("spp_wave", 16))  # samples per pixel per wave (16 = the library default of 2^24 paths at 1024^2)
ctx = scene.commit(0, **bvh_kw)
if bvh_kw:
    print("bvh", bvh_kw, {k: scene.bvh_stats[k] for k in ("n_nodes", "depth", "sah_cost")})
ctx.set_camera(*cam.device_record())
W, H = cam.get_resolution()
acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
per_bounce = []
prev = (0, 0)
for depth in range(1, 9):  # rays per bounce: difference of the counters of depth-d and depth-(d-1) renders of the same samples
    ctx.reset_counters()
    ctx.render(ctx.render_params(seed=1, spp_begin=0, spp_end=16, max_depth=depth, flags=_abi.RENDER_EXACT_PRIMARY), acc)
    c = ctx.counters()
    per_bounce.append((c["rays_closest"] - prev[0], c["rays_shadow"] - prev[1]))
    prev = (c["rays_closest"], c["rays_shadow"])
print(json.dumps({"rays_per_bounce_closest": [p[0] for p in per_bounce], "rays_per_bounce_shadow": [p[1] for p in per_bounce]}))
torch.cuda.synchronize()
if SPW != 16:
    ctx.set_wave_paths(SPW * W * H)
if len(sys.argv) > 2 and sys.argv[2] == "classes":  # per-kernel-class ms of n_waves 16-spp waves (cudaEvent pairs, best of 3)
    best = None
    for rep in range(3):
        ctx.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n_waves):
            ctx.render(ctx.render_params(seed=1, spp_begin=SPW * k, spp_end=SPW * k + SPW, max_depth=8, flags=_abi.RENDER_EXACT_PRIMARY), acc)
        e1.record()
        torch.cuda.synchronize()
        prof = {k: round(v[0] / n_waves * 16 / SPW, 4) for k, v in ctx.profile_end().items()}  # per 16 spp
        prof["wave_ms"] = round(e0.elapsed_time(e1) / n_waves * 16 / SPW, 4)
        if best is None or prof["wave_ms"] < best["wave_ms"]:
            best = prof
    print(json.dumps(best))
    sys.exit(0)
for k in range(n_waves):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.render(ctx.render_params(seed=1, spp_begin=16 * k, spp_end=16 * k + 16, max_depth=8, flags=_abi.RENDER_EXACT_PRIMARY), acc)
    e1.record()
    torch.cuda.synchronize()
    print(f"wave {k}: {e0.elapsed_time(e1):.3f} ms")
