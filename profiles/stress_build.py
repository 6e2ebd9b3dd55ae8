"""Builder stress run: many scene sizes and shapes, bit-exact BVH ids against the exhaustive GPU answer
(PRT_TRACE_BRUTE | PRT_TRACE_EXACT walks every triangle with the same exact tests, no BVH), plus the
triangle count the emitted leaves hold.  The pytest suite covers the boundaries (tests/test_gpu_trace.py);
this sweeps a few hundred sizes once after a builder change.
usage: python profiles/stress_build.py [n_cases]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrenderer_b200 import _abi

EXACT, BRUTE = _abi.TRACE_EXACT, _abi.TRACE_BRUTE
dev = torch.device("cuda", 0)
ctx = _abi.Context(0)
rng = np.random.default_rng(2026)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300


def scene(nt, kind):
    if kind == 0:    # soup
        h = 0.75 * nt ** (-1.0 / 3.0)
        c = rng.uniform(0, 1, (nt, 1, 3)); e = rng.uniform(-h, h, (nt, 2, 3))
        return np.concatenate([c, c + e[:, :1], c + e[:, 1:]], 1).astype(np.float32)
    if kind == 1:    # clustered
        k = max(1, nt // 500)
        cen = rng.uniform(0, 1, (k, 1, 3))
        return (cen[rng.integers(0, k, nt)] + rng.normal(0, 0.003, (nt, 3, 3))).astype(np.float32)
    if kind == 2:    # grid of axis-aligned quads split in two (coplanar pairs, shared edges, duplicates of some)
        m = int(np.ceil(np.sqrt(nt / 2))) + 1
        xs = np.linspace(0, 1, m + 1, dtype=np.float32)
        t = []
        for i in range(m):
            for j in range(m):
                a, b, c, d = (xs[i], xs[j], 0.5), (xs[i + 1], xs[j], 0.5), (xs[i + 1], xs[j + 1], 0.5), (xs[i], xs[j + 1], 0.5)
                t.append((a, b, c)); t.append((a, c, d))
        t = np.asarray(t, np.float32)[:nt]
        if nt > 8:
            t[rng.integers(0, nt, nt // 8)] = t[rng.integers(0, nt, nt // 8)]
        return t
    # mixed scales: a few huge triangles over many tiny ones + degenerate (zero-area) ones
    t = scene(nt, 0)
    big = rng.integers(0, nt, max(1, nt // 50))
    t[big] = rng.uniform(-2, 3, (big.size, 3, 3)).astype(np.float32)
    deg = rng.integers(0, nt, max(1, nt // 40))
    t[deg, 2] = t[deg, 1]
    return t


def rays(n):
    r = np.empty((n, 8), np.float32)
    r[:, 0:3] = rng.uniform(-0.2, 1.2, (n, 3))
    d = rng.normal(size=(n, 3)); r[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    r[: n // 4, 6] = 0.0                       # a quarter in-plane / axis-parallel
    r[: n // 8, 5] = 0.0
    nrm = np.linalg.norm(r[:, 4:7], axis=1, keepdims=True); r[:, 4:7] /= np.maximum(nrm, 1e-20)
    r[:, 3] = 1e-5; r[:, 7] = 3.4e38
    return r


bad = 0
sizes = sorted(set([1, 2, 3, 4, 5, 6, 7, 8, 9, 31, 32, 33, 127, 128, 129, 255, 256, 257, 4095, 4096, 4097]
                   + [int(x) for x in np.exp(rng.uniform(np.log(2), np.log(60000), n_cases))]))
for k, nt in enumerate(sizes):
    kind = k % 4
    tris = scene(nt, kind)
    nr = 4096 if nt > 20000 else 16384
    r = torch.from_numpy(rays(nr)).to(dev)
    opts = [dict(), dict(max_leaf_tris=1), dict(max_leaf_tris=7, cost_tri=0.5), dict(rotations=0), dict(rotations=2)][k % 5]
    ctx.set_triangles(tris)
    st = ctx.build_bvh(**opts)
    hb = torch.empty((nr, 4), dtype=torch.float32, device=dev); hx = torch.empty_like(hb)
    ctx.trace_closest(r, nr, hb, EXACT | BRUTE)
    ctx.trace_closest(r, nr, hx, EXACT)
    torch.cuda.synchronize()
    same = bool(torch.equal(hb[:, 3].view(torch.int32), hx[:, 3].view(torch.int32))) and bool(torch.equal(hb[:, 0], hx[:, 0]))
    if not same or st["n_tris"] != nt or 3 * st["depth"] + 1 > 128:
        bad += 1
        print("MISMATCH", nt, kind, opts, st, flush=True)
for nt, kind in ((300_000, 1), (1_000_000, 3), (2_000_000, 2)):  # a few large ones, sampled rays, BVH vs exhaustive
    tris = scene(nt, kind)
    r = torch.from_numpy(rays(2048)).to(dev)
    ctx.set_triangles(tris); st = ctx.build_bvh()
    hb = torch.empty((2048, 4), dtype=torch.float32, device=dev); hx = torch.empty_like(hb)
    ctx.trace_closest(r, 2048, hb, EXACT | BRUTE); ctx.trace_closest(r, 2048, hx, EXACT); torch.cuda.synchronize()
    same = bool(torch.equal(hb[:, 3].view(torch.int32), hx[:, 3].view(torch.int32)))
    print("large", nt, kind, "ok" if same else "MISMATCH", {k: st[k] for k in ("n_nodes", "depth", "ms_wall", "morton_bits")}, flush=True)
    bad += 0 if same else 1
print(json.dumps({"cases": len(sizes) + 3, "mismatches": bad}))
sys.exit(1 if bad else 0)
