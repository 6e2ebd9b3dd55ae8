"""Small driver for ncu: N-triangle soup, closest-hit batches of 2^24 incoherent rays (= one bench.py step).
usage: python profiles/prof_trace.py [n_launches] [exact|fp32] [n_tris] [log2 rays per launch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import soup  # noqa: E402
from pyrenderer_b200 import _abi  # noqa: E402

n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
flags = _abi.TRACE_EXACT if (len(sys.argv) > 2 and sys.argv[2] == "exact") else 0
n_tris = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
N = 1 << (int(sys.argv[4]) if len(sys.argv) > 4 else 24)  # the bench launch: 2^24 rays
dev = torch.device("cuda", 0)
ctx = _abi.Context(0)
ctx.set_triangles_dev(torch.from_numpy(soup(n_tris)).to(dev), n_tris)
print(ctx.build_bvh())
g = torch.Generator(device=dev)
g.manual_seed(11)
r = torch.empty((N, 8), dtype=torch.float32, device=dev)
r[:, 0:3] = torch.rand((N, 3), generator=g, device=dev)
d = torch.randn((N, 3), generator=g, device=dev)
r[:, 4:7] = d / d.norm(dim=1, keepdim=True)
r[:, 3] = 1e-5
r[:, 7] = 3.4e38
hits = torch.empty((N, 4), dtype=torch.float32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n_launch):
    e0.record()
    ctx.trace_closest(r, N, hits, flags)
    e1.record()
    torch.cuda.synchronize()
    print(f"launch {i}: {e0.elapsed_time(e1):.3f} ms  {N / e0.elapsed_time(e1) / 1e3:.1f} Mrays/s")
