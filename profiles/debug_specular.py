"""Where do GPU and oracle renders differ when specular materials are present?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from pyrenderer_b200 import _abi
from pyrenderer_b200.io_utils.read_tungsten import read_file
from pyrenderer_b200.main import DEFAULT_SCENE

scene, cam = read_file(DEFAULT_SCENE)
base = scene.arrays()
W = H = 64
iview, sw, sh, focal, _, _ = cam.device_record()
ocam = oracle.make_camera(iview, sh, sh, focal, W, H)
ctx = _abi.Context(0)

def run(label, edit, depth=8, spp=32):
    a = {k: v.copy() for k, v in base.items()}
    edit(a["materials"])
    ctx.set_triangles(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"])
    ctx.build_bvh()
    ctx.set_camera(iview, sh, sh, focal, W, H)
    kw = dict(seed=11, spp_begin=0, spp_end=spp, max_depth=depth)
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    ctx.render(ctx.render_params(**kw), acc)
    torch.cuda.synchronize()
    g = acc.cpu().numpy()[..., :3].astype(np.float64)
    o, _, st = oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"], ocam, oracle.make_params(**kw))
    o = o[..., :3]
    d = np.abs(g - o).max(axis=2)
    rel = np.sqrt(np.mean((g - o) ** 2)) / np.mean(o)
    bad = np.argwhere(d > 1e-3 * o.mean() * spp)
    print(f"{label}: relRMSE {rel:.3e} maxdiff {d.max():.3e} (mean sum {o.mean():.3f}) bad px {len(bad)} / {W*H}; first {bad[:6].tolist()}")
    for (y, x) in bad[:4]:
        print("   px", y, x, "gpu", g[y, x], "oracle", o[y, x])

def none(m): pass
def mirror(m): m[2]["type"] = 2
def diel(m): m[5]["type"], m[5]["ior"], m[5]["two_sided"], m[5]["albedo"] = 3, 1.5, 0, (1.0, 1.0, 1.0)
def cond(m): m[6]["type"], m[6]["roughness"], m[6]["albedo"] = 4, 0.15, (0.9, 0.8, 0.6)
def cond0(m): m[6]["type"], m[6]["roughness"], m[6]["albedo"] = 4, 0.0, (0.9, 0.8, 0.6)
for label, f in (("lambert", none), ("mirror", mirror), ("dielectric", diel), ("conductor.15", cond), ("conductor0", cond0)):
    run(label, f)
    run(label + " d3", f, depth=3)
