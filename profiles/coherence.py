"""Experiment: how much faster is the soup-1M closest-hit kernel when the SAME incoherent rays arrive
in a spatially binned order (origin cell + direction octant)?  Sorting here is done with torch.sort
outside the timed region -- this measures only the potential of reordering.
usage: python profiles/coherence.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import soup
from pyrenderer_b200 import _abi
N = 1 << 24
dev = torch.device("cuda", 0)
ctx = _abi.Context(0)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx.set_triangles_dev(torch.from_numpy(soup(NT)).to(dev), NT)
ctx.build_bvh()
g = torch.Generator(device=dev); g.manual_seed(11)
r = torch.empty((N, 8), dtype=torch.float32, device=dev)
r[:, 0:3] = torch.rand((N, 3), generator=g, device=dev)
d = torch.randn((N, 3), generator=g, device=dev)
r[:, 4:7] = d / d.norm(dim=1, keepdim=True)
r[:, 3] = 1e-5; r[:, 7] = 3.4e38
hits = torch.empty((N, 4), dtype=torch.float32, device=dev)

def spread(v, bits):
    out = torch.zeros_like(v)
    for b in range(bits):
        out |= ((v >> b) & 1) << (3 * b)
    return out

def timed(rays):
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.trace_closest(rays, N, hits, 0); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return N / best / 1e3

print(f"unsorted: {timed(r):8.1f} Mrays/s", flush=True)
for bits in (3, 5, 7):
    for octant in (0, 1):
        q = (r[:, 0:3].clamp(0, 0.999999) * (1 << bits)).to(torch.int64)
        key = (spread(q[:, 0], bits) << 2) | (spread(q[:, 1], bits) << 1) | spread(q[:, 2], bits)
        if octant:
            o = ((r[:, 4] < 0).to(torch.int64) << 2) | ((r[:, 5] < 0).to(torch.int64) << 1) | (r[:, 6] < 0).to(torch.int64)
            key = (key << 3) | o
        perm = torch.argsort(key)
        rs = r[perm].contiguous()
        print(f"cells {1 << bits}^3 octant {octant} ({3 * bits + 3 * octant} key bits): {timed(rs):8.1f} Mrays/s", flush=True)
