"""GPU parity tests for the intersection / BVH path, through the C ABI.

Contract (BASELINE.json north_star): triangle ids bit-exact against the
reference's intersection code (here: the oracle, itself bit-exact against the
reference's numba kernel -- tests/test_oracle_golden.py); t within 4 ulp(f32).
PRT_TRACE_EXACT is the parity mode and runs on the SAME persistent-warp kernel
as the throughput mode (csrc/persist.cuh): FP32 watertight traversal with forward
error bounds, every low-margin triangle decided in place with the reference's
FP64 formula, what still cannot be ordered replayed in FP64.  The plain FP32
mode is also measured and must disagree on at most a tiny, reported fraction of
rays (thresholds = 3x the measured fraction, gpurun r2-8).
"""
import numpy as np
import pytest

import oracle
from conftest import CUBE_OBJ, make_rays, random_rays, random_soup

pytestmark = pytest.mark.gpu

F32MAX = 3.4028234663852886e+38
EXACT, COUNT, BRUTE = 1, 2, 4


def _torch():
    import torch
    return torch


def gpu_closest(ctx, rays, flags):
    torch = _torch()
    r = torch.from_numpy(np.ascontiguousarray(rays, np.float32)).cuda()
    h = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    ctx.trace_closest(r, rays.shape[0], h, flags)
    torch.cuda.synchronize()
    h = h.cpu().numpy()
    return h[:, 3].copy().view(np.int32), h[:, 0].astype(np.float64), h[:, 1], h[:, 2]


def check_against_oracle(ctx, tris, rays, flags, label, ref=None):
    ids_o, t_o, u_o, v_o = ref if ref is not None else oracle.closest_hit(tris, rays)
    ctx.reset_counters()
    ids_g, t_g, u_g, v_g = gpu_closest(ctx, rays, flags)
    bad = np.nonzero(ids_g != ids_o)[0]
    assert bad.size == 0, f"{label}: {bad.size} id mismatches, first {bad[:5]} gpu {ids_g[bad[:5]]} ref {ids_o[bad[:5]]}"
    hit = ids_o >= 0
    rel = np.abs(t_g[hit] - t_o[hit]) / np.maximum(np.abs(t_o[hit]), 1e-30)
    assert rel.max() <= 4 * 2.0 ** -23, f"{label}: t off by {rel.max() / 2.0 ** -23:.2f} ulp"
    assert np.abs(u_g[hit] - u_o[hit]).max() < 1e-3 and np.abs(v_g[hit] - v_o[hit]).max() < 1e-3
    c = ctx.counters()
    print(f"[{label}] rays {rays.shape[0]} hits {hit.sum()} flagged(FP64 replay) {c['flagged_rays']}")
    return c["flagged_rays"]


def test_abi_roundtrip_and_errors(gpu_ctx):
    from pyrenderer_b200 import _abi
    ctx = _abi.Context(0)
    rays = random_rays(64)
    torch = _torch()
    r = torch.from_numpy(rays).cuda()
    h = torch.empty((64, 4), dtype=torch.float32, device="cuda")
    with pytest.raises(_abi.PrtError) as e:  # trace before a scene exists
        ctx.trace_closest(r, 64, h, 0)
    assert "[-3]" in str(e.value) and "scene" in str(e.value)
    ctx.set_triangles(random_soup(10))
    with pytest.raises(_abi.PrtError) as e:  # BVH traversal before the build
        ctx.trace_closest(r, 64, h, 0)
    assert "BVH not built" in str(e.value)
    ctx.trace_closest(r, 64, h, BRUTE)  # brute force needs no BVH
    with pytest.raises(_abi.PrtError):
        ctx.set_triangles(random_soup(4), light_tris=[9])  # light id out of range
    with pytest.raises(_abi.PrtError):
        ctx.build_bvh(max_leaf_tris=9)
    ctx.close()
    with pytest.raises(_abi.PrtError):
        _abi.Context(99)


@pytest.mark.parametrize("case", ["cornell", "soup64"])
def test_golden_rays(gpu_ctx, golden, case):
    """The rays the REFERENCE's numba kernel was run on (make_golden.py)."""
    tris = golden[f"ch_{case}_tris"]
    rays = make_rays(golden[f"ch_{case}_o"], golden[f"ch_{case}_d"], tmin=1.1754943508222875e-38, tmax=F32MAX)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build_bvh()
    for flags, label in ((EXACT | BRUTE, "brute"), (EXACT, "bvh")):
        ids_g, t_g, _, _ = gpu_closest(gpu_ctx, rays, flags)
        assert np.array_equal(ids_g, golden[f"ch_{case}_ids"]), label
        hit = ids_g >= 0
        rel = np.abs(t_g[hit] - golden[f"ch_{case}_t"][hit]) / golden[f"ch_{case}_t"][hit]
        assert rel.max() <= 4 * 2.0 ** -23


def test_cornell_primary_ids_1024(gpu_ctx, cornell):
    """Every primary ray of the 1024x1024 Cornell camera (pixel centres): device ray
    generation bit-equal to the oracle camera, ids bit-exact in EXACT mode."""
    torch = _torch()
    scene, cam = cornell
    a = scene.arrays()
    gpu_ctx.set_triangles(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"])
    st = gpu_ctx.build_bvh()
    assert st["n_tris"] == 36
    iview, sw, sh, focal, W, H = cam.device_record()
    gpu_ctx.set_camera(iview, sw, sh, focal, W, H)
    rays_d = torch.empty((H * W, 8), dtype=torch.float32, device="cuda")
    gpu_ctx.generate_rays(rays_d, seed=0, s0=0, s1=1, jitter=False, tmin=1e-5, tmax=99999.9)
    torch.cuda.synchronize()
    rays = rays_d.cpu().numpy()
    ocam = oracle.make_camera(iview, sw, sh, focal, W, H)
    rays_o = oracle.generate_rays(ocam, jitter=False).reshape(-1, 8)
    assert np.array_equal(rays.view(np.uint32), rays_o.view(np.uint32)), "device raygen != oracle camera"
    ref = oracle.closest_hit(a["tris"], rays)
    flagged = check_against_oracle(gpu_ctx, a["tris"], rays, EXACT, "cornell-1024-bvh", ref)
    check_against_oracle(gpu_ctx, a["tris"], rays, EXACT | BRUTE, "cornell-1024-brute", ref)
    assert flagged < 0.01 * rays.shape[0]
    # plain FP32 mode: report, and bound, the disagreement
    ids_o = ref[0]
    ids_f, _, _, _ = gpu_closest(gpu_ctx, rays, 0)
    frac = np.mean(ids_f != ids_o)
    print(f"[cornell-1024] FP32-only id mismatch fraction {frac:.2e}")
    assert frac < 5e-4  # measured 1.74e-4: pixel-centre rays that run exactly along the quads' diagonals


def test_cube_obj_primary_ids_1024(gpu_ctx):
    """BASELINE config 2: media/cube.obj, 1024x1024 primary rays, camera of SURVEY 8d."""
    from pyrenderer_b200.core.camera import Camera
    from pyrenderer_b200.io_utils.read_tungsten import read_obj
    torch = _torch()
    v, f = read_obj(CUBE_OBJ)
    tris = v[f].astype(np.float32)
    cam = Camera([2.6, 2.1, 3.4], [0.5, 0.5, 0.5], [0, 1, 0], [1024, 1024], fov=35.0)
    iview, sw, sh, focal, W, H = cam.device_record()
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build_bvh()
    gpu_ctx.set_camera(iview, sw, sh, focal, W, H)
    rays_d = torch.empty((H * W, 8), dtype=torch.float32, device="cuda")
    gpu_ctx.generate_rays(rays_d, jitter=False, tmin=1e-5, tmax=F32MAX)
    torch.cuda.synchronize()
    rays = rays_d.cpu().numpy()
    ocam = oracle.make_camera(iview, sw, sh, focal, W, H)
    assert np.array_equal(rays.view(np.uint32),
                          oracle.generate_rays(ocam, jitter=False, tmax=F32MAX).reshape(-1, 8).view(np.uint32))
    check_against_oracle(gpu_ctx, tris, rays, EXACT, "cube-1024-bvh")
    ids, _, _, _ = gpu_closest(gpu_ctx, rays, EXACT)
    assert 0.05 < np.mean(ids >= 0) < 0.9 and set(np.unique(ids)) <= set(range(-1, 12))


@pytest.mark.parametrize("nt,nr,leaf", [(1, 2000, 4), (2, 2000, 4), (37, 20000, 1), (5000, 100000, 4), (20000, 100000, 7)])
def test_random_soup_bvh_vs_oracle(gpu_ctx, nt, nr, leaf):
    tris = random_soup(nt, seed=nt)
    if nt <= 2:
        tris *= 0.0
        tris += random_soup(nt, seed=5) * 0.5 + 0.25  # big triangles so that rays hit
    rays = random_rays(nr, seed=nt + 1)
    gpu_ctx.set_triangles(tris)
    st = gpu_ctx.build_bvh(max_leaf_tris=leaf)
    assert st["n_tris"] == nt and 3 * st["depth"] + 1 <= 128 and st["morton_sorted"] == 1
    ref = oracle.closest_hit(tris, rays)
    check_against_oracle(gpu_ctx, tris, rays, EXACT, f"soup{nt}-bvh", ref)
    check_against_oracle(gpu_ctx, tris, rays, EXACT | BRUTE, f"soup{nt}-brute", ref)
    ids_o = ref[0]
    ids_f, _, _, _ = gpu_closest(gpu_ctx, rays, 0)
    frac = np.mean(ids_f != ids_o)
    print(f"[soup{nt}] FP32-only id mismatch fraction {frac:.2e}")
    assert frac < 5e-5  # measured 0 on every one of these soups (<= 100k rays)


def test_rotations_and_leaf_sizes_do_not_change_hits(gpu_ctx):
    tris = random_soup(30000, seed=3)
    rays = random_rays(100000, seed=4)
    ref = None
    visits = {}
    for leaf, rot, treelets in ((1, 0, 0), (1, 1, 0), (4, 1, 0), (7, 0, 0), (1, 0, 1), (4, 1, 1), (7, 2, 1)):
        gpu_ctx.set_triangles(tris)
        st = gpu_ctx.build_bvh(max_leaf_tris=leaf, rotations=rot, treelets=treelets)
        assert st["n_tris"] == 30000 and 3 * st["depth"] + 1 <= 128
        ids, t, _, _ = gpu_closest(gpu_ctx, rays, EXACT)
        if ref is None:
            ref = (ids, t)
        assert np.array_equal(ids, ref[0]) and np.array_equal(t, ref[1]), (leaf, rot, treelets)
        gpu_ctx.reset_counters()
        gpu_closest(gpu_ctx, rays, COUNT)
        visits[(leaf, rot, treelets)] = gpu_ctx.counters()["node_visits"] / rays.shape[0]
    # the SAH treelets (every maximal subtree of <= 128 triangles rebuilt with binned SAH) are
    # there to cut record visits per ray; same hits, fewer visits
    print(f"[treelets] record visits per ray: LBVH {visits[(1, 0, 0)]:.2f} -> treelets {visits[(1, 0, 1)]:.2f}")
    assert visits[(1, 0, 1)] < 0.99 * visits[(1, 0, 0)]


def test_treelets_on_clustered_and_tiny_scenes(gpu_ctx):
    """Treelet edge cases: whole tree smaller than one treelet (3..130 triangles), coincident
    centroids (all planes degenerate -> median split), and a clustered soup."""
    rng = np.random.default_rng(17)

    def check(tris, rays, label):
        gpu_ctx.set_triangles(tris)
        st = gpu_ctx.build_bvh()
        assert st["n_tris"] == tris.shape[0]
        check_against_oracle(gpu_ctx, tris, rays, EXACT, label)

    for nt in (3, 4, 5, 63, 64, 65, 127, 128, 129, 130):
        check(random_soup(nt, seed=100 + nt), random_rays(3000, seed=nt), f"tiny{nt}")
    one = random_soup(1, seed=5)
    tris = np.repeat(one, 200, axis=0)  # 200 coincident triangles: lowest id must win everywhere
    rays = random_rays(2000, seed=6)
    rays[:, 0:3] = tris[0].mean(axis=0) - 0.5 * rays[:, 4:7]
    check(tris, rays, "coincident200")
    centres = rng.uniform(0, 1, (40, 1, 3))
    tris = (centres[rng.integers(0, 40, 20000)] + rng.normal(0, 0.004, (20000, 3, 3))).astype(np.float32)
    check(tris, random_rays(20000, seed=8), "clustered20000")


def test_deep_stacks_spill_and_unspill(gpu_ctx):
    """4000 large triangles that all overlap: nearly every child box is hit at every level, so the
    per-lane stacks grow far past the 16 shared-memory levels and exercise the spill / unspill
    path (bvh.cuh sstack_spill) in every kernel family: persistent closest / any (plain FP32),
    one-ray-per-thread exact closest, all-hits (never culls: deepest) -- all against the oracle."""
    torch = _torch()
    rng = np.random.default_rng(77)
    c = rng.uniform(0.3, 0.7, (4000, 1, 3))
    tris = (c + rng.uniform(-0.6, 0.6, (4000, 3, 3))).astype(np.float32)
    rays = random_rays(6000, seed=78)
    gpu_ctx.set_triangles(tris)
    st = gpu_ctx.build_bvh(max_leaf_tris=1)
    assert 3 * st["depth"] + 1 <= 128
    ids_o, t_o, _, _ = oracle.closest_hit(tris, rays)
    cnt_o, sums_o = oracle.all_hits(tris, rays)
    occ_o = oracle.any_hit(tris, rays)
    assert cnt_o.mean() > 100  # every ray crosses hundreds of triangles
    check_against_oracle(gpu_ctx, tris, rays, EXACT, "overlap4000", ref=oracle.closest_hit(tris, rays))
    r = torch.from_numpy(rays).cuda()
    cnt = torch.empty(rays.shape[0], dtype=torch.int32, device="cuda")
    sums = torch.empty(rays.shape[0], dtype=torch.int64, device="cuda")
    occ = torch.empty(rays.shape[0], dtype=torch.uint8, device="cuda")
    gpu_ctx.trace_all(r, rays.shape[0], cnt, sums, EXACT)
    gpu_ctx.trace_any(r, rays.shape[0], occ, 0)
    torch.cuda.synchronize()
    assert np.array_equal(cnt.cpu().numpy().view(np.uint32), cnt_o) and np.array_equal(sums.cpu().numpy().view(np.uint64), sums_o)
    assert np.mean(occ.cpu().numpy() != occ_o) < 1e-3
    ids_f, t_f, _, _ = gpu_closest(gpu_ctx, rays, 0)  # persistent kernel, plain FP32
    frac = np.mean(ids_f != ids_o)
    print(f"[overlap4000] mean hits per ray {cnt_o.mean():.0f}, FP32-only closest id mismatch fraction {frac:.2e}")
    assert frac < 5e-4  # measured 1.67e-4 (199 hits per ray: near-ties between overlapping triangles)
    same = ids_f == ids_o
    assert np.abs(t_f[same & (ids_o >= 0)] - t_o[same & (ids_o >= 0)]).max() < 1e-4


def test_any_and_all_hits(gpu_ctx):
    torch = _torch()
    tris = random_soup(3000, seed=9)
    rays = random_rays(50000, seed=10, tmax=0.35)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build_bvh()
    r = torch.from_numpy(rays).cuda()
    occ = torch.empty(rays.shape[0], dtype=torch.uint8, device="cuda")
    cnt = torch.empty(rays.shape[0], dtype=torch.int32, device="cuda")
    sums = torch.empty(rays.shape[0], dtype=torch.int64, device="cuda")
    occ_o = oracle.any_hit(tris, rays)
    cnt_o, sums_o = oracle.all_hits(tris, rays)
    assert 0.2 < occ_o.mean() < 0.98 and cnt_o.max() >= 3
    for flags in (EXACT, EXACT | BRUTE):
        gpu_ctx.trace_any(r, rays.shape[0], occ, flags)
        gpu_ctx.trace_all(r, rays.shape[0], cnt, sums, flags)
        torch.cuda.synchronize()
        assert np.array_equal(occ.cpu().numpy(), occ_o)
        assert np.array_equal(cnt.cpu().numpy().view(np.uint32), cnt_o)  # the full hit SET per ray
        assert np.array_equal(sums.cpu().numpy().view(np.uint64), sums_o)
    # throughput (plain FP32, persistent-warp) any-hit kernel: disagreement must be tiny
    gpu_ctx.trace_any(r, rays.shape[0], occ, 0)
    torch.cuda.synchronize()
    frac = np.mean(occ.cpu().numpy() != occ_o)
    print(f"[any-hit] FP32-only mismatch fraction {frac:.2e}")
    assert frac < 1e-4
    # counted twin of the throughput closest-hit kernel returns the same hits as the plain one
    h0 = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    h1 = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    gpu_ctx.reset_counters()
    gpu_ctx.trace_closest(r, rays.shape[0], h0, 0)
    gpu_ctx.trace_closest(r, rays.shape[0], h1, COUNT)
    torch.cuda.synchronize()
    assert torch.equal(h0.view(torch.int32), h1.view(torch.int32))  # bit compare (tri == -1 is a NaN pattern)
    c = gpu_ctx.counters()
    assert c["rays_closest"] == rays.shape[0] and c["node_visits"] > rays.shape[0] and c["tri_tests"] > 0


def test_accept_rule_known_answers_on_the_gpu(gpu_ctx):
    """The crafted cases of tests/test_oracle_golden.py::test_oracle_edge_cases_and_accept_rule through the C ABI:
    rays through edges and vertices (inclusive), t exactly at tmin / tmax (inclusive), stacked identical triangles
    (lowest id wins, the hit SET holds all of them), a quad's shared diagonal -- closest, any and all hits, BVH and
    exhaustive walk, against the oracle."""
    torch = _torch()
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
    stack = np.concatenate([tri + np.float32([0, 0, -1]), tri, tri, tri + np.float32([0, 0, 0.5]), tri + np.float32([0, 0, 0.5])])
    quad = np.array([[[0, 0, 0], [1, 0, 0], [1, 1, 0]], [[0, 0, 0], [1, 1, 0], [0, 1, 0]]], np.float32)
    pts = [[0.25, 0.25], [0.5, 0.0], [0.0, 0.5], [0.5, 0.5], [0.0, 0.0], [1.0, 0.0], [0.0, 1.0],
           [0.5, -1e-6], [-1e-6, 0.5], [0.5 + 1e-6, 0.5 + 1e-6]]
    o = [[x, y, 1.0] for x, y in pts] + [[x, y, -2.0] for x, y in pts] + [[-1, 0.25, 0.0], [0.25, 0.25, 1.0]]
    d = [[0, 0, -1]] * len(pts) + [[0, 0, 1]] * len(pts) + [[1, 0, 0], [0, 0, 1]]
    rays = make_rays(o, d)
    rng_t = make_rays([[0.25, 0.25, 1.0]] * 4, [[0, 0, -1]] * 4)
    rng_t[:, 3] = [1.0, np.nextafter(np.float32(1.0), np.float32(2.0)), 0.0, 0.0]
    rng_t[:, 7] = [9.0, 9.0, 1.0, np.nextafter(np.float32(1.0), np.float32(0.0))]
    rays = np.concatenate([rays, rng_t])
    for name, tris in (("tri", tri), ("stack", stack), ("quad", quad), ("quad-reversed", quad[::-1].copy())):
        gpu_ctx.set_triangles(tris)
        gpu_ctx.build_bvh()
        ref = oracle.closest_hit(tris, rays)
        assert np.any(ref[0] >= 0) and np.any(ref[0] < 0)
        check_against_oracle(gpu_ctx, tris, rays, EXACT, f"accept/{name}", ref)
        check_against_oracle(gpu_ctx, tris, rays, EXACT | BRUTE, f"accept/{name}/brute", ref)
        r = torch.from_numpy(rays).cuda()
        occ = torch.empty(rays.shape[0], dtype=torch.uint8, device="cuda")
        cnt = torch.empty(rays.shape[0], dtype=torch.int32, device="cuda")
        sums = torch.empty(rays.shape[0], dtype=torch.int64, device="cuda")
        cnt_o, sums_o = oracle.all_hits(tris, rays)
        for flags in (EXACT, EXACT | BRUTE):
            gpu_ctx.trace_any(r, rays.shape[0], occ, flags)
            gpu_ctx.trace_all(r, rays.shape[0], cnt, sums, flags)
            torch.cuda.synchronize()
            assert np.array_equal(occ.cpu().numpy().astype(bool), ref[0] >= 0), name
            assert np.array_equal(cnt.cpu().numpy().view(np.uint32), cnt_o), name
            assert np.array_equal(sums.cpu().numpy().view(np.uint64), sums_o), name
    ref = oracle.closest_hit(stack, rays)
    assert ref[0][0] == 3 and ref[0][len(pts)] == 0  # from above: the lower id of the two at z = 0.5; from below: id 0


def test_edge_cases(gpu_ctx, cornell):
    torch = _torch()
    # empty scene
    gpu_ctx.set_triangles(np.zeros((0, 3, 3), np.float32))
    gpu_ctx.build_bvh()
    ids, _, _, _ = gpu_closest(gpu_ctx, random_rays(100), EXACT)
    assert np.all(ids == -1)
    ids, _, _, _ = gpu_closest(gpu_ctx, random_rays(100), BRUTE)
    assert np.all(ids == -1)
    # zero rays
    gpu_ctx.trace_closest(torch.empty((0, 8), device="cuda"), 0, torch.empty((0, 4), device="cuda"), 0)
    # degenerate (zero-area, duplicated) triangles mixed into a soup + coincident centroids
    tris = random_soup(500, seed=21)
    tris[10] = tris[10, 0]  # point
    tris[11, 2] = tris[11, 1]  # segment
    tris[20:40] = tris[20]  # 20 identical triangles: identical Morton codes AND exact t ties
    rays = random_rays(30000, seed=22)
    # aim a few thousand rays exactly at the duplicated triangle
    c = tris[20].mean(axis=0)
    d = c[None] - rays[:3000, 0:3]
    rays[:3000, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build_bvh(max_leaf_tris=2)
    check_against_oracle(gpu_ctx, tris, rays, EXACT, "degenerate")
    ids, _, _, _ = gpu_closest(gpu_ctx, rays, EXACT)
    assert np.sum(ids == 20) > 100 and not np.any((ids > 20) & (ids < 40))  # lowest id wins ties
    # Cornell: axis-aligned rays, rays through shared edges / vertices, origins on surfaces
    scene, _ = cornell
    a = scene.arrays()
    gpu_ctx.set_triangles(a["tris"])
    gpu_ctx.build_bvh()
    o, d = [], []
    for x in np.linspace(-0.9, 0.9, 37):
        for y in np.linspace(0.1, 1.9, 37):
            o.append([x, y, 0.9]); d.append([0, 0, -1])       # axis aligned, hits back wall / boxes
            o.append([x, 0.0, y - 1.0]); d.append([0, 1, 0])  # origin ON the floor
            o.append([0.0, 1.0, 0.5]); d.append([x, y - 1.0, -0.5])  # through the wall diagonals
    o = np.array(o, np.float32); d = np.array(d, np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = make_rays(o, d.astype(np.float32), tmin=1e-5, tmax=99999.9)
    check_against_oracle(gpu_ctx, a["tris"], rays, EXACT, "cornell-edges")
    check_against_oracle(gpu_ctx, a["tris"], rays, EXACT | BRUTE, "cornell-edges-brute")
    # axis-parallel rays whose zero components are NEGATIVE zeros (e.g. the product -1 * 0.0): the slab
    # test's octant must follow the sign of the clamped reciprocal (-1e20), or near and far planes swap
    # and a ray travelling inside a slab misses every box
    n = rays.shape[0] // 3
    rz = rays[0::3][:n].copy()   # the (0, 0, -1) family
    rz[:, 4] = -0.0
    rz[:, 5] = -0.0
    ry = rays[1::3][:n].copy()   # the (0, 1, 0) family, origins on the floor
    ry[:, 4] = -0.0
    ry[:, 6] = -0.0
    neg = np.concatenate([rz, ry])
    assert np.all(np.signbit(neg[:, 4]))
    ref = oracle.closest_hit(a["tris"], neg)
    assert np.mean(ref[0] >= 0) > 0.9
    check_against_oracle(gpu_ctx, a["tris"], neg, EXACT, "negative-zero-directions", ref)
    ids_f, _, _, _ = gpu_closest(gpu_ctx, neg, 0)
    assert np.mean(ids_f != ref[0]) < 0.02  # plain FP32: these rays run along the quads' shared edges


def test_host_buffer_entry_point(gpu_ctx):
    tris = random_soup(2000, seed=31)
    rays = random_rays(20000, seed=32)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build_bvh()
    h = gpu_ctx.trace_closest_host(rays, EXACT)
    ids_o, _, _, _ = oracle.closest_hit(tris, rays)
    assert np.array_equal(h["tri"], ids_o)


def test_host_pipeline_many_chunks_equals_device_call(gpu_ctx):
    """prt_trace_closest_host streams 2^21-ray chunks through 4 staging slots on three streams
    (upload / trace / download).  11 chunks (ragged last one) reuse every slot at least twice; the
    result must be bit-identical to one device-resident call, and an odd-aligned device pointer
    (ray array starting 32 bytes + 16 into an allocation) must take the 128-bit load path."""
    torch = _torch()
    tris = random_soup(5000, seed=41)
    n = 10 * (1 << 21) + 12345
    rng = np.random.default_rng(42)
    rays = np.empty((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(0, 1, (n, 3))
    d = rng.normal(size=(n, 3))
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3] = 1e-5
    rays[:, 7] = 3.4e38
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build_bvh()
    h = gpu_ctx.trace_closest_host(rays, 0)
    rd = torch.from_numpy(rays).cuda()
    hd = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    gpu_ctx.trace_closest(rd, n, hd, 0)
    torch.cuda.synchronize()
    dev = hd.cpu().numpy().view(np.uint32)
    assert np.array_equal(h.view(np.uint32).reshape(n, 4), dev)
    # misaligned ray array: 16 bytes past a 32-byte boundary
    m = 100000
    buf = torch.empty((m * 8 + 4,), dtype=torch.float32, device="cuda")
    view = buf[4:].view(m, 8)
    view.copy_(rd[:m])
    assert view.data_ptr() % 32 == 16
    h2 = torch.empty((m, 4), dtype=torch.float32, device="cuda")
    gpu_ctx.trace_closest(view, m, h2, 0)
    torch.cuda.synchronize()
    assert np.array_equal(h2.cpu().numpy().view(np.uint32), dev[:m])


def test_million_triangle_soup_bvh_equals_exhaustive(gpu_ctx):
    """BASELINE config 4 size (1M triangles, the bench scene): the EXACT persistent kernel -- the
    one bench.py times as `exact` -- equals the CPU oracle on 4096 rays (4e9 reference triangle
    tests) and the exhaustive GPU answer on 2^14 rays; the plain FP32 throughput kernel (bench
    `value`) is compared with both at the same size."""
    tris = random_soup(1_000_000, seed=7)
    rays = random_rays(1 << 17, seed=11)
    gpu_ctx.set_triangles(tris)
    st = gpu_ctx.build_bvh()
    print("[soup1M] bvh", st)
    assert st["n_tris"] == 1_000_000 and 0 < st["n_nodes"] < 1_000_000 and st["morton_sorted"] == 1
    gpu_ctx.reset_counters()
    ids_b, t_b, _, _ = gpu_closest(gpu_ctx, rays, EXACT)
    flagged = gpu_ctx.counters()["flagged_rays"]
    ids_x, t_x, _, _ = gpu_closest(gpu_ctx, rays[: 1 << 14], EXACT | BRUTE)
    assert np.array_equal(ids_b[: 1 << 14], ids_x) and np.array_equal(t_b[: 1 << 14], t_x)
    n_o = 4096
    ids_o, t_o, _, _ = oracle.closest_hit(tris, rays[:n_o])
    assert np.array_equal(ids_b[:n_o], ids_o)
    hit = ids_o >= 0
    assert (np.abs(t_b[:n_o][hit] - t_o[hit]) / t_o[hit]).max() <= 4 * 2.0 ** -23
    assert np.mean(ids_b >= 0) > 0.8
    # throughput kernel (flags = 0) against the oracle and against the exact kernel
    ids_f, t_f, _, _ = gpu_closest(gpu_ctx, rays, 0)
    assert np.sum(ids_f[:n_o] != ids_o) <= 1
    n_bad = int(np.sum(ids_f != ids_b))
    print(f"[soup1M] exact: FP64 replays {flagged} of {rays.shape[0]}; plain FP32 vs exact: {n_bad} id mismatches")
    assert n_bad <= 3  # measured 7.7e-7 of 2^24 rays (profiles/prof_exact.py) -> 0.1 expected in 2^17
    gpu_ctx.reset_counters()
    gpu_closest(gpu_ctx, rays, COUNT)
    c = gpu_ctx.counters()
    gpu_ctx.reset_counters()
    gpu_closest(gpu_ctx, rays, COUNT | EXACT)
    cx = gpu_ctx.counters()
    print(f"[soup1M] node visits / tri tests per ray: fp32 {c['node_visits'] / rays.shape[0]:.2f} / {c['tri_tests'] / rays.shape[0]:.2f}, "
          f"exact {cx['node_visits'] / rays.shape[0]:.2f} / {cx['tri_tests'] / rays.shape[0]:.2f}, "
          f"in-place FP64 triangle decisions {cx['f64_decisions'] / rays.shape[0]:.2e} per ray")
    # the error-bound widening must not cost visits (a bound applied to the wrong axis once cost 3 %)
    assert cx["node_visits"] < 1.01 * c["node_visits"]


def test_fuzz_small_scenes_exact_ids(gpu_ctx):
    """Many small random scenes through the whole builder (Morton sort, Karras, SAH treelets,
    refit + rotations, 4-wide emit) and the exact closest-hit path: sizes around every builder
    boundary (1, 2, treelet size 128, its multiples), duplicated triangles, degenerate (zero-area)
    triangles and widely different scales in one scene."""
    rng = np.random.default_rng(2024)
    sizes = [1, 2, 3, 7, 31, 100, 128, 129, 255, 256, 257, 300, 511, 640, 1000]
    for i, nt in enumerate(sizes):
        kind = i % 3
        if kind == 0:
            tris = random_soup(nt, seed=500 + nt)
        elif kind == 1:  # duplicates + degenerate triangles
            base = random_soup(max(1, nt // 2), seed=600 + nt)
            tris = base[np.arange(nt) % base.shape[0]].copy()
            tris[::5, 2] = tris[::5, 1]  # zero-area
        else:  # mixed scales: a few huge triangles over many tiny ones
            tris = random_soup(nt, seed=700 + nt)
            big = rng.integers(0, nt, max(1, nt // 20))
            tris[big] = (rng.uniform(-3, 4, (big.size, 3, 3))).astype(np.float32)
        rays = random_rays(1500, seed=nt)
        gpu_ctx.set_triangles(tris)
        st = gpu_ctx.build_bvh(max_leaf_tris=(1, 4, 7)[i % 3])
        assert st["n_tris"] == nt and st["morton_sorted"] == 1
        check_against_oracle(gpu_ctx, tris, rays, EXACT, f"fuzz{nt}/{kind}")
        ids_o = oracle.closest_hit(tris, rays)[0]
        ids_f = gpu_closest(gpu_ctx, rays, 0)[0]
        # measured 0 mismatches on all of them; duplicated triangles tie exactly (lowest id wins in both modes)
        assert np.sum(ids_f != ids_o) <= 1, f"fuzz{nt}/{kind}: plain FP32 mismatches {np.sum(ids_f != ids_o)} of 1500"


def test_builder_paths_against_the_exhaustive_walk(gpu_ctx):
    """Every path of the builder -- treelet warps that refit their own subtree (default) and the unfused
    refit from the leaves (rotations = 2, treelets = 0), exact enumeration of 3..6-triangle ranges, the
    one-chunk and the 16-bin split, the warp-aggregated emit -- on soups, clusters, coplanar grids with
    duplicates and mixed scales with degenerate triangles: ids AND t of the EXACT BVH traversal equal the
    exhaustive EXACT walk over all triangles (same tests, no BVH), and a rebuild gives the same tree size."""
    torch = _torch()
    rng = np.random.default_rng(99)

    def scene(nt, kind):
        if kind == 0:
            return random_soup(nt, seed=nt)
        if kind == 1:  # clusters
            k = max(1, nt // 300)
            cen = rng.uniform(0, 1, (k, 1, 3))
            return (cen[rng.integers(0, k, nt)] + rng.normal(0, 0.003, (nt, 3, 3))).astype(np.float32)
        if kind == 2:  # coplanar quads split in two, an eighth of them duplicated
            m = int(np.ceil(np.sqrt(nt / 2))) + 1
            xs = np.linspace(0, 1, m + 1, dtype=np.float32)
            i, j = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
            a = np.stack([xs[i], xs[j], np.full_like(xs[i], 0.5)], -1).reshape(-1, 3)
            b = np.stack([xs[i + 1], xs[j], np.full_like(xs[i], 0.5)], -1).reshape(-1, 3)
            c = np.stack([xs[i + 1], xs[j + 1], np.full_like(xs[i], 0.5)], -1).reshape(-1, 3)
            d = np.stack([xs[i], xs[j + 1], np.full_like(xs[i], 0.5)], -1).reshape(-1, 3)
            t = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 0)[:nt].astype(np.float32)
            if nt > 8:
                t[rng.integers(0, nt, nt // 8)] = t[rng.integers(0, nt, nt // 8)]
            return t
        t = random_soup(nt, seed=3 * nt)  # mixed scales + zero-area triangles
        big = rng.integers(0, nt, max(1, nt // 50))
        t[big] = rng.uniform(-2, 3, (big.size, 3, 3)).astype(np.float32)
        deg = rng.integers(0, nt, max(1, nt // 40))
        t[deg, 2] = t[deg, 1]
        return t

    option_sets = [dict(), dict(max_leaf_tris=1), dict(max_leaf_tris=7, cost_tri=0.5), dict(rotations=0),
                   dict(rotations=2), dict(treelets=0)]
    sizes = [3, 4, 5, 6, 7, 9, 33, 127, 130, 700, 4097, 20000, 50000]
    for k, nt in enumerate(sizes):
        for kind in range(4):
            tris = scene(nt, kind)
            nr = 4096
            rays = random_rays(nr, seed=1000 + nt + kind)
            rays[: nr // 4, 6] = 0.0  # a quarter of the directions lie in a coordinate plane
            rays[:, 4:7] /= np.maximum(np.linalg.norm(rays[:, 4:7], axis=1, keepdims=True), 1e-20)
            r = torch.from_numpy(rays).cuda()
            opts = option_sets[(k + kind) % len(option_sets)]
            gpu_ctx.set_triangles(tris)
            st = gpu_ctx.build_bvh(**opts)
            assert st["n_tris"] == nt and 3 * st["depth"] + 1 <= 128 and st["morton_sorted"] == 1, (nt, kind, opts, st)
            hb = torch.empty((nr, 4), dtype=torch.float32, device="cuda")
            hx = torch.empty_like(hb)
            gpu_ctx.trace_closest(r, nr, hb, EXACT | BRUTE)
            gpu_ctx.trace_closest(r, nr, hx, EXACT)
            torch.cuda.synchronize()
            assert torch.equal(hb[:, 3].view(torch.int32), hx[:, 3].view(torch.int32)), (nt, kind, opts)
            assert torch.equal(hb[:, 0], hx[:, 0]), (nt, kind, opts)
            st2 = gpu_ctx.build_bvh(**opts)  # same input, same options: same tree size and cost
            assert st2["n_nodes"] == st["n_nodes"] and st2["sah_cost"] == st["sah_cost"], (nt, kind, opts)


def test_morton_63_bit_keys_and_grow_only_buffers(gpu_ctx):
    """63-bit Morton keys (21 bits per axis): picked automatically when the 30-bit grid cannot
    separate the triangles (a dense cluster inside a huge scene box), selectable by hand, and
    never a change of the hits.  Also: rebuilds reuse the context's buffers (prt_release_scratch
    gives the scratch back without touching the scene)."""
    rng = np.random.default_rng(63)
    n = 60000
    tris = (rng.uniform(0.4, 0.4002, (n, 1, 3)) + rng.uniform(-2e-6, 2e-6, (n, 3, 3))).astype(np.float32)
    tris[:8] = random_soup(8, seed=1) * 1000.0  # a few far-away triangles blow the scene box up to 1000 units
    rays = random_rays(20000, seed=64)
    d = tris[rng.integers(8, n, rays.shape[0])].mean(axis=1) - rays[:, 0:3]
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)  # aim at the cluster
    ref = oracle.closest_hit(tris, rays)
    assert np.mean(ref[0] >= 8) > 0.5
    gpu_ctx.set_triangles(tris)
    st = gpu_ctx.build_bvh()
    assert st["morton_bits"] == 63 and st["morton_sorted"] == 1 and st["ms_wall"] > 0
    check_against_oracle(gpu_ctx, tris, rays, EXACT, "cluster-auto63", ref)
    gpu_ctx.reset_counters()
    gpu_closest(gpu_ctx, rays, COUNT)
    v63 = gpu_ctx.counters()["node_visits"] / rays.shape[0]
    st30 = gpu_ctx.build_bvh(morton_bits=30)
    assert st30["morton_bits"] == 30
    check_against_oracle(gpu_ctx, tris, rays, EXACT, "cluster-30", ref)
    gpu_ctx.reset_counters()
    gpu_closest(gpu_ctx, rays, COUNT)
    v30 = gpu_ctx.counters()["node_visits"] / rays.shape[0]
    print(f"[morton] clustered scene: record visits per ray 30-bit {v30:.1f} -> 63-bit {v63:.1f}")
    assert v63 < v30
    # an ordinary soup stays on 30 bits in auto mode; forcing 63 changes no hit
    tris = random_soup(50000, seed=65)
    rays = random_rays(50000, seed=66)
    gpu_ctx.set_triangles(tris)
    assert gpu_ctx.build_bvh()["morton_bits"] == 30
    ids30, t30, _, _ = gpu_closest(gpu_ctx, rays, EXACT)
    assert gpu_ctx.build_bvh(morton_bits=63)["morton_bits"] == 63
    ids63, t63, _, _ = gpu_closest(gpu_ctx, rays, EXACT)
    assert np.array_equal(ids30, ids63) and np.array_equal(t30, t63)
    with pytest.raises(Exception):
        gpu_ctx.build_bvh(morton_bits=48)
    gpu_ctx.release_scratch()
    ids2, _, _, _ = gpu_closest(gpu_ctx, rays, EXACT)  # scene + BVH survive; scratch is re-grown on demand
    assert np.array_equal(ids2, ids63)
    st2 = gpu_ctx.build_bvh()
    assert st2["n_tris"] == 50000


def test_ray_binning_changes_the_order_not_the_answers(gpu_ctx):
    """PRT_TRACE_BIN: rays are counting-sorted by the cell of their origin before the traversal (indices
    only).  Every mode must return, bit for bit, what it returns without binning -- also for origins
    outside the scene box, a ragged batch, and the chunked host pipeline."""
    from pyrenderer_b200 import _abi
    torch = _torch()
    tris = random_soup(30000, seed=91)
    n = 300001
    rays = random_rays(n, seed=92)
    rays[:1000, 0:3] = rays[:1000, 0:3] * 8.0 - 3.0  # origins far outside the unit box
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build_bvh()
    r = torch.from_numpy(rays).cuda()
    for flags in (0, EXACT):
        h0 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        h1 = torch.full((n, 4), -7.0, dtype=torch.float32, device="cuda")
        gpu_ctx.trace_closest(r, n, h0, flags | _abi.TRACE_NO_BIN)
        gpu_ctx.trace_closest(r, n, h1, flags | _abi.TRACE_BIN)
        torch.cuda.synchronize()
        assert torch.equal(h0.view(torch.int32), h1.view(torch.int32)), flags
    o0 = torch.empty(n, dtype=torch.uint8, device="cuda")
    o1 = torch.full((n,), 9, dtype=torch.uint8, device="cuda")
    rs = rays.copy(); rs[:, 7] = 0.3
    rsd = torch.from_numpy(rs).cuda()
    gpu_ctx.trace_any(rsd, n, o0, _abi.TRACE_NO_BIN)
    gpu_ctx.trace_any(rsd, n, o1, _abi.TRACE_BIN)
    torch.cuda.synchronize()
    assert torch.equal(o0, o1) and 0.05 < o0.float().mean().item() < 0.99
    ids_o = oracle.closest_hit(tris, rays[:3000])[0]
    hb = gpu_ctx.trace_closest_host(rays, EXACT | _abi.TRACE_BIN)
    assert np.array_equal(hb["tri"][:3000], ids_o)
    hn = gpu_ctx.trace_closest_host(rays, EXACT | _abi.TRACE_NO_BIN)
    assert np.array_equal(hb.view(np.uint32), hn.view(np.uint32))
