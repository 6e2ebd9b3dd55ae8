"""The CPU oracle (oracle/pt_oracle.c) against fixtures produced by the
REFERENCE's own code (tests/golden/make_golden.py).  No GPU needed."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, make_rays

F32MAX = 3.4028234663852886e+38


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in oracle.philox4x32_10(ctr, key)) == want


@pytest.mark.parametrize("case", ["cornell", "soup64"])
def test_closest_hit_matches_reference_grouped_kernel(golden, case):
    """ids AND t bit-exact vs mathematics/intersection.py:68-118 (numba, f64)."""
    tris = golden[f"ch_{case}_tris"]
    rays = make_rays(golden[f"ch_{case}_o"], golden[f"ch_{case}_d"], tmin=1.1754943508222875e-38,
                     tmax=F32MAX)
    ids, t, _, _ = oracle.closest_hit(tris, rays)
    assert np.array_equal(ids, golden[f"ch_{case}_ids"])
    hit = ids >= 0
    assert hit.sum() > 100
    assert np.array_equal(t[hit], golden[f"ch_{case}_t"][hit])


def test_scalar_kernel_decisions(golden):
    """Decisions of intersection.py:7-39 (np.cross/np.dot) == oracle scalar port; t to 1e-12
    (BLAS dot vs explicit sums differ in the last bits -- SURVEY B.4)."""
    tris = golden["ch_soup64_tris"].astype(np.float64)
    o = golden["ch_soup64_o"][:300].astype(np.float64)
    d = golden["ch_soup64_d"][:300].astype(np.float64)
    dec, tt = golden["scalar_dec"], golden["scalar_t"]
    for i in range(0, 300, 3):
        for k in range(64):
            hit, t = oracle.mt_scalar(tris[k, 0], tris[k, 1], tris[k, 2], o[i], d[i])
            assert hit == bool(dec[i, k])
            if hit:
                assert abs(t - tt[i, k]) <= 1e-12 * max(1.0, abs(t))


def test_grouped_and_scalar_agree(golden):
    """the reference's two formulations accept the same set."""
    tris = golden["ch_soup64_tris"]
    rays = make_rays(golden["ch_soup64_o"][:300], golden["ch_soup64_d"][:300], tmin=1.1754943508222875e-38, tmax=F32MAX)
    cnt, _ = oracle.all_hits(tris, rays)
    assert np.array_equal(cnt, golden["scalar_dec"].sum(axis=1))


def test_slab(golden):
    g = golden
    for i in range(g["slab_o"].shape[0]):
        hit, t0 = oracle.slab(0.0, F32MAX, g["slab_o"][i], g["slab_inv"][i], g["slab_bmin"][i], g["slab_bmax"][i])
        assert hit == (g["slab_res"][i, 0] > 0)
        if hit:
            assert t0 == g["slab_res"][i, 1]


def test_concentric_disk(golden):
    for u, want in zip(golden["disk_u"], golden["disk_res"]):
        got = oracle.concentric_sample_disk(u[0], u[1])
        assert np.allclose(got, want, rtol=0, atol=1e-15)


def test_frame_and_cosine_sample(golden):
    for n, want in zip(golden["frame_n"], golden["frame_res"]):
        r1, r2, r3 = oracle.frame_z_to(n)
        assert np.allclose(np.stack([r1, r2, r3]), want, rtol=0, atol=1e-15)
    for i, n in enumerate(golden["frame_n"]):
        u = golden["disk_u"][i]
        got = oracle.cosine_sample_hemisphere(n, u[0], u[1])
        assert np.allclose(got, golden["cos_res"][i], rtol=0, atol=1e-14)
        assert np.dot(got, n / np.linalg.norm(n)) >= -1e-12


def test_camera(golden):
    from math import radians, tan
    specs = [(19.5, 1024, 1024), (35.0, 640, 480)]
    for ci, (fov, w, h) in enumerate(specs):
        sh = tan(radians(fov) / 2) * 1.0
        cam = oracle.make_camera(golden[f"cam{ci}_iview"], sh * (w / h * 1.0), sh, 1.0, w, h)
        for uv, want in zip(golden["cam_uv"], golden[f"cam{ci}_rays"]):
            o, d = oracle.generate_ray(cam, uv[0], uv[1])
            assert np.allclose(o, want[:3], rtol=0, atol=1e-15)
            if ci == 0:  # Cornell camera (pure translation): bit-exact in f64
                assert np.array_equal(d, want[3:])
            # rotated camera: numpy's BLAS matmul sums in another order -> <= 2 ulp(f64);
            # the f32 ray record that crosses the device boundary is identical
            assert np.allclose(d, want[3:], rtol=0, atol=1e-15)
            assert np.array_equal(d.astype(np.float32), want[3:].astype(np.float32))


def test_golden_path_of_reference_test_py(golden, cornell):
    """test.py:38-57: 9 recorded bounces.  Through the oracle on the restated loader's
    geometry: expected triangle ids, |dt| <= 5e-6, normals, albedo."""
    scene, _ = cornell
    a = scene.arrays()
    gp, rep = golden["golden_path"], golden["golden_path_replay"]
    want_ids = [4, 7, 5, 23, 4, 1, 7, 1, 7]
    rays = make_rays(gp[:, 1:4], gp[:, 4:7], tmin=1e-5, tmax=999.9)
    ids, t, _, _ = oracle.closest_hit(a["tris"], rays)
    assert list(ids) == want_ids
    assert np.all(np.abs(t - gp[:, 0]) <= 5e-6)
    # The reference's own replay (its NumPy Scene.hit, f64 geometry).  Row 3 starts exactly
    # on the back wall and the reference self-hits it at t ~ 2e-17 because its lower bound is
    # EPS = 1.18e-38 (SURVEY App. A.1 hazard); the recorded path and the oracle use t_min = 1e-5.
    ok = rep[:, 0] > 1e-5
    assert list(np.nonzero(~ok)[0]) == [3]
    assert np.all(np.abs(t - rep[:, 0])[ok] <= 1e-6)
    for k, tri in enumerate(ids):
        n = a["normals"][tri].astype(np.float64)
        if np.dot(n, -gp[k, 4:7]) < 0:
            n = -n
        assert np.allclose(n, gp[k, 13:16], atol=1.5e-6)
        if ok[k]:
            assert np.allclose(n, rep[k, 1:4], atol=1e-6)
        alb = a["materials"][a["tri_material"][tri]]["albedo"]
        assert np.allclose(alb, gp[k, 10:13], atol=1e-6)
        if ok[k]:
            assert np.allclose(alb, rep[k, 4:7], atol=1e-6)
    # row k+1 starts where row k ended
    for k in range(8):
        assert np.allclose(gp[k, 1:4] + t[k] * gp[k, 4:7], gp[k + 1, 1:4], atol=2e-5)


# ---------------------------------------------------------------------------------------------
# radiance + specular BSDFs: fixtures composed from the reference's own functions
# (tests/golden/make_radiance_golden.py)
RADIANCE = os.path.join(os.path.dirname(GOLDEN), "radiance_golden.npz")


@pytest.fixture(scope="module")
def rad_golden():
    return np.load(RADIANCE)


def test_radiance_matches_the_reference_composed_path_tracer(rad_golden):
    """Per-path radiance of a 32x32x4-spp Cornell render traced by a path tracer composed of the
    reference's imported functions (numba Moller-Trumbore, compute_pos, Camera.generate_ray,
    samplers_debug.cosine_sample_hemisphere, shapes2 sample_a_point) under the estimator of
    core/tracing.py:116-155: oracle.render reproduces every path to 1e-12 and every primary id."""
    g = rad_golden
    W, H, SPP, DEPTH, SEED = (int(x) for x in g["params"])
    mats = np.ascontiguousarray(g["materials"]).view(oracle.MATERIAL_DTYPE).reshape(-1)
    from pyrenderer_b200.core.camera import Camera
    cam = Camera([0, 1, 6.8], [0, 1, 0], [0, 1, 0], [W, H], fov=19.5)
    iview, sw, sh, focal, _, _ = cam.device_record()
    assert np.allclose(iview.reshape(4, 4), g["cam_iview"], rtol=0, atol=1e-15)
    ocam = oracle.make_camera(iview, sw, sh, focal, W, H)
    rays = oracle.generate_rays(ocam, seed=SEED, s0=0, s1=SPP, jitter=True)
    assert np.array_equal(rays[..., [0, 1, 2, 4, 5, 6]].view(np.uint32), g["rays"].view(np.uint32)), "camera rays differ"
    worst = 0.0
    for s in range(SPP):
        acc, ids, _ = oracle.render(g["tris"], g["normals"], g["tri_material"], mats, g["light_tris"], ocam,
                                    oracle.make_params(seed=SEED, spp_begin=s, spp_end=s + 1, max_depth=DEPTH), want_ids=True)
        assert np.array_equal(ids[..., 0], g["prim_ids"][..., s])
        ref = g["radiance"][:, :, s, :]
        err = np.abs(acc[..., :3] - ref).max() / np.abs(ref).max()
        worst = max(worst, float(err))
    print(f"[radiance] oracle vs reference-composed path tracer: max rel err {worst:.2e} over {W * H * SPP} paths")
    assert worst < 1e-12
    assert g["radiance"].mean() > 0.05 and np.mean(g["radiance"].sum(axis=-1) > 0) > 0.8


def test_specular_bsdfs_match_bsdf_taichi(rad_golden):
    """Schlick reflectance, reflect, refract, random_in_unit_sphere, Metal.scatter and
    Dielectric.scatter of core/bsdf_taichi.py / mathematics/vec3_taichi.py, run from the reference's
    source (taichi names bound to plain Python), against the oracle's functions."""
    g = rad_golden
    for i, c in enumerate(g["schlick_cos"]):
        for j, e in enumerate(g["schlick_idx"]):
            assert abs(oracle.schlick(c, e) - g["schlick_val"][i, j]) <= 1e-15
    # analytic anchors: normal incidence r0 = ((1-n)/(1+n))^2 (glass: 0.04), grazing incidence 1
    assert abs(oracle.schlick(1.0, 1.5) - 0.04) < 1e-15 and abs(oracle.schlick(0.0, 1.5) - 1.0) < 1e-15
    for k in range(g["spec_v"].shape[0]):
        v, n, eta = g["spec_v"][k], g["spec_n"][k], float(g["spec_eta"][k])
        assert np.abs(oracle.reflect(v, n) - g["reflect_res"][k]).max() <= 1e-15
        assert np.abs(oracle.refract(v, n, eta) - g["refract_res"][k]).max() <= 1e-14
        assert np.abs(oracle.in_unit_sphere(*g["sphere_u"][k]) - g["sphere_res"][k]).max() <= 1e-15
        d = v * g["metal_scale"][k]
        ok, wi = oracle.scatter_specular(4, d, n, roughness=float(g["metal_rough"][k]), u=g["sphere_u"][k])
        assert ok == bool(g["metal_ok"][k]) and np.abs(wi - g["metal_res"][k]).max() <= 1e-14
        ok, wi = oracle.scatter_specular(3, d, n, front=bool(g["diel_front"][k]), ior=float(g["diel_ior"][k]),
                                         u=(float(g["diel_u"][k]), 0.0, 0.0))
        assert ok and np.abs(wi - g["diel_res"][k]).max() <= 1e-14
        ok, wi = oracle.scatter_specular(2, d, n)  # mirror == Metal with roughness 0
        assert ok and np.abs(wi - g["reflect_res"][k]).max() <= 1e-14
    # the laws the functions stand for: reflection keeps the length and flips the normal component;
    # refraction obeys Snell (sin_t = eta sin_i) wherever it is not total internal reflection
    v, n, eta = g["spec_v"], g["spec_n"], g["spec_eta"]
    r = g["reflect_res"]
    assert np.allclose(np.linalg.norm(r, axis=1), 1.0, atol=1e-14) and np.allclose(np.sum(r * n, 1), -np.sum(v * n, 1), atol=1e-14)
    cos_i = -np.sum(v * n, 1)
    sin_i = np.sqrt(1 - cos_i ** 2)
    ok = eta * sin_i < 1.0
    t = g["refract_res"][ok]
    sin_t = np.linalg.norm(t - np.sum(t * n[ok], 1)[:, None] * n[ok], axis=1)
    assert np.allclose(sin_t, eta[ok] * sin_i[ok], atol=1e-13) and np.allclose(np.linalg.norm(t, axis=1), 1.0, atol=1e-13)
    assert np.all(np.sum(t * n[ok], 1) < 0)


def _rays(o, d, tmin=1e-5, tmax=3.4e38):
    o = np.asarray(o, np.float32).reshape(-1, 3)
    d = np.asarray(d, np.float32).reshape(-1, 3)
    r = np.empty((o.shape[0], 8), np.float32)
    r[:, 0:3], r[:, 3], r[:, 4:7], r[:, 7] = o, tmin, d, tmax
    return r


def test_oracle_edge_cases_and_accept_rule():
    """The accept rule of mathematics/intersection.py:42-65 and the tie rule of :106-116 /
    core/scene.py:66-73 as known answers (no GPU): inclusive edges and vertices, inclusive t range,
    lowest id on exact ties, no back-face culling, degenerate triangles never hit, empty inputs."""
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
    # empty scene / zero rays
    ids, t, _, _ = oracle.closest_hit(np.zeros((0, 3, 3), np.float32), _rays([[0, 0, 1]], [[0, 0, -1]]))
    assert ids.tolist() == [-1]
    ids, _, _, _ = oracle.closest_hit(tri, np.zeros((0, 8), np.float32))
    assert ids.shape == (0,)
    assert oracle.any_hit(np.zeros((0, 3, 3), np.float32), _rays([[0, 0, 1]], [[0, 0, -1]])).tolist() == [0]
    # interior, the three edges, the three vertices: all accepted (u, v >= 0, u + v <= 1 inclusive); from both sides
    pts = [[0.25, 0.25], [0.5, 0.0], [0.0, 0.5], [0.5, 0.5], [0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]
    for z, dz in ((1.0, -1.0), (-1.0, 1.0)):
        r = _rays([[x, y, z] for x, y in pts], [[0, 0, dz]] * len(pts))
        ids, t, u, v = oracle.closest_hit(tri, r)
        assert ids.tolist() == [0] * len(pts) and np.allclose(t, 1.0)
        assert np.allclose(u, [p[0] for p in pts]) and np.allclose(v, [p[1] for p in pts])
    # just outside every edge: miss
    r = _rays([[0.5, -1e-6, 1], [-1e-6, 0.5, 1], [0.5 + 1e-6, 0.5 + 1e-6, 1]], [[0, 0, -1]] * 3)
    assert oracle.closest_hit(tri, r)[0].tolist() == [-1, -1, -1]
    # t range is inclusive at both ends: tmin <= t <= tmax
    r = _rays([[0.25, 0.25, 1.0]] * 4, [[0, 0, -1]] * 4)
    r[:, 3] = [1.0, np.nextafter(np.float32(1.0), np.float32(2.0)), 0.0, 0.0]
    r[:, 7] = [9.0, 9.0, 1.0, np.nextafter(np.float32(1.0), np.float32(0.0))]
    assert oracle.closest_hit(tri, r)[0].tolist() == [0, -1, 0, -1]
    # a ray in the triangle's plane (det == 0 up to the reference's EPS) and a ray pointing away: miss
    r = _rays([[-1, 0.25, 0.0], [0.25, 0.25, 1.0]], [[1, 0, 0], [0, 0, 1]])
    assert oracle.closest_hit(tri, r)[0].tolist() == [-1, -1]
    # degenerate triangles (point, segment) are never hit
    deg = np.array([[[0.2, 0.2, 0]] * 3, [[0, 0, 0], [1, 0, 0], [1, 0, 0]]], np.float32)
    r = _rays([[0.2, 0.2, 1], [0.5, 0.0, 1]], [[0, 0, -1]] * 2)
    assert oracle.closest_hit(deg, r)[0].tolist() == [-1, -1]
    # exact ties: identical triangles -> the lowest id, whatever the order they are stored in; closest = min t
    stack = np.concatenate([tri + np.float32([0, 0, -1]), tri, tri, tri + np.float32([0, 0, 0.5]), tri + np.float32([0, 0, 0.5])])
    r = _rays([[0.25, 0.25, 2.0], [0.25, 0.25, 0.25], [0.25, 0.25, -3.0]], [[0, 0, -1], [0, 0, -1], [0, 0, 1]])
    ids, t, _, _ = oracle.closest_hit(stack, r)
    assert ids.tolist() == [3, 1, 0] and np.allclose(t, [1.5, 0.25, 2.0])
    cnt, sums = oracle.all_hits(stack, r)
    K = 0x9E3779B97F4A7C15  # the hit SET: every triangle on the line; checksum = sum((id + 1) * K) mod 2^64
    assert cnt.tolist() == [5, 3, 5]
    assert sums.tolist() == [sum((i + 1) * K for i in s) % 2 ** 64 for s in ((0, 1, 2, 3, 4), (0, 1, 2), (0, 1, 2, 3, 4))]
    assert oracle.any_hit(stack, r).tolist() == [1, 1, 1]
    # a shared edge between two triangles of a quad: both contain the point, the lower id wins
    quad = np.array([[[0, 0, 0], [1, 0, 0], [1, 1, 0]], [[0, 0, 0], [1, 1, 0], [0, 1, 0]]], np.float32)
    r = _rays([[0.5, 0.5, 1.0], [0.25, 0.25, 1.0]], [[0, 0, -1]] * 2)
    assert oracle.closest_hit(quad, r)[0].tolist() == [0, 0]
    assert oracle.closest_hit(quad[::-1].copy(), r)[0].tolist() == [0, 0]
    assert oracle.all_hits(quad, r)[0].tolist() == [2, 2]


def test_oracle_invariants_on_random_scenes():
    """Size-independent properties of the restated algorithm: the grouped kernel equals the scalar one on every
    (ray, triangle) pair it reports, a permutation of the triangles permutes the ids (no exact ties in a random
    soup), any-hit == (closest id >= 0), the all-hits count bounds both, and a hit never leaves [tmin, tmax]."""
    rng = np.random.default_rng(5)
    nt, nr = 400, 3000
    c = rng.uniform(0, 1, (nt, 1, 3))
    tris = np.concatenate([c, c + rng.uniform(-0.1, 0.1, (nt, 2, 3))], 1).astype(np.float32)
    d = rng.normal(size=(nr, 3))
    rays = _rays(rng.uniform(0, 1, (nr, 3)), d / np.linalg.norm(d, axis=1, keepdims=True), tmin=1e-5, tmax=0.8)
    ids, t, u, v = oracle.closest_hit(tris, rays)
    hit = ids >= 0
    assert 0.15 < hit.mean() < 1.0
    assert np.all((t[hit] >= 1e-5) & (t[hit] <= np.float64(np.float32(0.8))))
    assert np.all((u[hit] >= 0) & (v[hit] >= 0) & (u[hit] + v[hit] <= 1 + 1e-12))
    perm = rng.permutation(nt)
    ids_p, t_p, _, _ = oracle.closest_hit(tris[perm], rays)
    assert np.array_equal(np.where(ids_p >= 0, perm[np.maximum(ids_p, 0)], -1), ids) and np.array_equal(t_p, t)
    occ = oracle.any_hit(tris, rays)
    assert np.array_equal(occ.astype(bool), hit)
    cnt, sums = oracle.all_hits(tris, rays)
    K = np.uint64(0x9E3779B97F4A7C15)
    assert np.all((cnt > 0) == hit) and np.all(sums[cnt == 1] == (ids[cnt == 1].astype(np.uint64) + np.uint64(1)) * K)
    for k in np.nonzero(hit)[0][:200]:  # scalar kernel (intersection.py:7-39) on the reported pairs
        ok, ts = oracle.mt_scalar(*tris[ids[k]].astype(np.float64), rays[k, 0:3].astype(np.float64),
                                  rays[k, 4:7].astype(np.float64), float(rays[k, 7]))
        assert ok and abs(ts - t[k]) <= 1e-12 * max(1.0, abs(t[k]))
