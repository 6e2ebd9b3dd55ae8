import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import soup
from pyrenderer_b200 import _abi
N = 1 << 24
dev = torch.device("cuda", 0)
ctx = _abi.Context(0)
g = torch.Generator(device=dev); g.manual_seed(11)
r = torch.empty((N, 8), dtype=torch.float32, device=dev)
r[:, 0:3] = torch.rand((N, 3), generator=g, device=dev)
d = torch.randn((N, 3), generator=g, device=dev)
r[:, 4:7] = d / d.norm(dim=1, keepdim=True); r[:, 3] = 1e-5; r[:, 7] = 3.4e38
hits = torch.empty((N, 4), dtype=torch.float32, device=dev)
for nt in (1_000_000, 10_000_000):
    ctx.set_triangles_dev(torch.from_numpy(soup(nt)).to(dev), nt); ctx.build_bvh()
    for name, fl in (("fp32 nobin", _abi.TRACE_NO_BIN), ("fp32 bin", _abi.TRACE_BIN), ("exact nobin", 1 | _abi.TRACE_NO_BIN), ("exact bin", 1 | _abi.TRACE_BIN), ("fp32 default", 0)):
        ctx.trace_closest(r, N, hits, fl); torch.cuda.synchronize()
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.trace_closest(r, N, hits, fl); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"soup {nt}: {name:12s} {best:7.3f} ms  {N / best / 1e3:7.1f} Mrays/s", flush=True)
