"""BVH build phases (median of 7 builds) + tree quality (record visits / triangle tests per ray of the plain
persistent kernel, 2^22 incoherent rays) on the 1M and 10M soups.
usage: python profiles/prof_build.py [n_tris ...] [rotations=R]"""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import soup
from pyrenderer_b200 import _abi

dev = torch.device("cuda", 0)
ctx = _abi.Context(0)
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1_000_000, 10_000_000]
kw = {a.split("=")[0]: int(a.split("=")[1]) for a in sys.argv[1:] if "=" in a}
N = 1 << 22
g = torch.Generator(device=dev); g.manual_seed(11)
r = torch.empty((N, 8), dtype=torch.float32, device=dev)
r[:, 0:3] = torch.rand((N, 3), generator=g, device=dev)
d = torch.randn((N, 3), generator=g, device=dev)
r[:, 4:7] = d / d.norm(dim=1, keepdim=True); r[:, 3] = 1e-5; r[:, 7] = 3.4e38
hits = torch.empty((N, 4), dtype=torch.float32, device=dev)
for nt in sizes:
    t = torch.from_numpy(soup(nt)).to(dev)
    sts, walls = [], []
    for _ in range(7):
        ctx.set_triangles_dev(t, nt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sts.append(ctx.build_bvh(**kw))
        walls.append((time.perf_counter() - t0) * 1e3)
    med = {k: float(np.median([s[k] for s in sts])) for k in sts[0] if k.startswith("ms_")}
    ctx.reset_counters()
    ctx.trace_closest(r, N, hits, _abi.TRACE_COUNT)
    torch.cuda.synchronize()
    c = ctx.counters()
    hx = torch.empty_like(hits)
    ctx.trace_closest(r, N, hx, _abi.TRACE_EXACT)  # bit-exact ids do not depend on the tree
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.trace_closest(r, N, hits, 0); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"n_tris": nt, "opts": kw, "wall_ms_median": float(np.median(walls)), **med,
                      "n_nodes": sts[-1]["n_nodes"], "depth": sts[-1]["depth"], "sah_cost": sts[-1]["sah_cost"],
                      "n_node_per_ray": c["node_visits"] / N, "n_tri_per_ray": c["tri_tests"] / N,
                      "hits_checksum": int(hx[:, 3].view(torch.int32).to(torch.int64).sum().item()),
                      "mrays_s_2p22": N / best / 1e3}), flush=True)
    del t
