#!/usr/bin/env python
"""Radiance and specular-BSDF golden vectors, composed from the REFERENCE's own functions.

Run in the build container only (needs /root/reference):  python tests/golden/make_radiance_golden.py
Writes tests/golden/radiance_golden.npz.

(1) RADIANCE.  The reference has no runnable integrator at HEAD (main.py imports a
    path_tracing() that no longer exists; the live estimator is a Taichi @ti.func).  This script
    composes one path tracer out of the reference's IMPORTED, unmodified functions:
      * ray / triangle: mathematics/intersection.py triangle_ray_intersection_grouping (the numba
        kernel), called once per triangle so that its literal lower bound EPS (intersection.py:49)
        can be replaced by the estimator's t_min = 1e-5 (core/tracing.py:127) without letting a
        self-hit at t ~ 1e-17 shrink the ray bound; closest = min t, lowest id
        (intersection.py:106-116, core/scene.py:66-73);
      * hit position: mathematics/fast_op.py compute_pos (o + d t);
      * camera rays: core/camera.py Camera.generate_ray under main.py:31-33's (i + rand) / W;
      * bounce direction: mathematics/samplers_debug.py cosine_sample_hemisphere (np.random.rand patched
        to return the Philox uniforms);
      * light point: mathematics/shapes2.py Quad.sample_a_point (random.randint / random.uniform
        patched the same way);
    and restates, line by line, only the glue that cannot be imported: the two-sided normal flip of
    shapes2.py:93-96 and the estimator of core/tracing.py:116-155 (+ sample_direct_lighting :92-108)
    with the ambiguities resolved as in SURVEY App. A.6 (Q7 shadow t_max = |p2-p| (1 - 1e-4); Q8 the
    estimator runs under main.py's pixel x sample loop).  Random numbers: Philox4x32-10 keyed as
    pyrenderer_b200 keys it (counter = pixel, sample, bounce, block), implemented here in Python and
    checked against the Random123 known-answer vectors.
    Inputs = the C-ABI scene arrays (f32 triangles / normals / materials of pyrenderer_b200's loader,
    themselves pinned to the reference's debug loader by reference_golden.npz).
    Stored: per-path radiance of a 32 x 32 x 4-spp Cornell render, depth 5 -- oracle.render must
    reproduce it to 1e-12.

(2) SPECULAR BSDFs.  core/bsdf_taichi.py:6-22 (reflectance = Schlick, reflect, refract) and :45-86
    (Metal.scatter, Dielectric.scatter) with mathematics/vec3_taichi.py:33-39 (random_in_unit_sphere) are
    Taichi functions whose bodies are plain arithmetic; with `ti.func` as the identity, ti.sqrt / cos /
    sin / acos as math functions, ti.random patched and ts.vec3 as a small numpy vector class, the
    reference's own source runs in Python.  Stored: inputs and outputs of those functions.
"""
import math
import os
import random as py_random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as mg  # noqa: E402  (stubs for the third-party modules the reference imports)

REF = mg.REF
W = H = 32
SPP, DEPTH, SEED = 4, 5, 20261018
T_MIN, T_MAX = float(np.float32(1e-5)), float(np.float32(99999.9))
LIGHT_COLOR = [float(np.float32(x)) for x in (0.9, 0.85, 0.7)]
INV_PI = 0.31830988618379067154


# ---- Philox4x32-10 (Salmon et al. 2011), counter = (pixel, sample, bounce, block), key = seed
def philox4x32_10(ctr, key):
    c = [int(x) for x in ctr]
    k = [int(x) for x in key]
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def rng4(pixel, sample, bounce, block):
    return philox4x32_10([pixel, sample, bounce, block], [SEED & 0xFFFFFFFF, SEED >> 32])


def u24(k):
    return (k >> 8) * (1.0 / 16777216.0)


def install_taichi_runtime():
    """Functional stand-ins for the taichi / taichi_glsl names bsdf_taichi.py and vec3_taichi.py use."""
    ti, ts = sys.modules["taichi"], sys.modules["taichi_glsl"]

    class Vec3(np.ndarray):
        def __new__(cls, x=0.0, y=0.0, z=0.0):
            return np.asarray([x, y, z], np.float64).view(cls)

        def dot(self, o):
            return float(np.asarray(self)[0] * np.asarray(o)[0] + np.asarray(self)[1] * np.asarray(o)[1] + np.asarray(self)[2] * np.asarray(o)[2])

        def norm_sqr(self):
            return self.dot(self)

        def norm(self):
            return math.sqrt(self.norm_sqr())

        def normalized(self):
            return (self / self.norm()).view(Vec3)

    def ident(f=None, **kw):
        return f

    ti.func = ident
    ti.kernel = ident
    ti.data_oriented = ident
    ti.sqrt, ti.cos, ti.sin, ti.acos = math.sqrt, math.cos, math.sin, math.acos
    ti.abs = abs
    ti.random = lambda *a: py_random.random()
    ts.vec3 = Vec3
    ts.vec4 = lambda *a: np.asarray(a, np.float64)
    ts.mat = lambda *a: np.asarray(a, np.float64)
    sys.modules["taichi_glsl.vector"].reflect = lambda v, n: v - 2.0 * v.dot(n) * n
    return Vec3


class Uniforms:
    """Hands a fixed list of numbers to the patched random sources, in order."""

    def __init__(self):
        self.q = []

    def set(self, *vals):
        self.q = list(vals)

    def pop(self):
        return self.q.pop(0)


def main():
    mg._install_stubs()
    Vec3 = install_taichi_runtime()  # before ANY reference import: vec3_taichi.py binds Vector = ts.vec3 at import time
    sys.path.insert(0, REF)
    mount = tempfile.mkdtemp()
    os.symlink(REF, os.path.join(mount, "pyr"))
    sys.path.insert(0, mount)
    cwd = os.getcwd()
    os.chdir(REF)
    import importlib
    from core.ray import Ray
    from mathematics import intersection as ref_int
    from mathematics import fast_op as ref_fast
    from mathematics import samplers_debug as ref_samp
    from mathematics import shapes2 as ref_shapes2
    cam_mod = importlib.import_module("pyr.core.camera")

    import oracle
    for c, k in (([0, 0, 0, 0], [0, 0]), ([0xffffffff] * 4, [0xffffffff] * 2),
                 ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])):
        assert list(oracle.philox4x32_10(c, k)) == philox4x32_10(c, k), "Python Philox != oracle Philox"

    out = {}
    # ------------------------------------------------------------------ scene = the C-ABI arrays
    os.chdir(cwd)
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    scene, _ = read_file(os.path.join(ROOT, "pyrenderer_b200", "media", "cornell_box.json"))
    os.chdir(REF)
    a = scene.arrays()
    tris = a["tris"].astype(np.float32).astype(np.float64)        # [36,3,3]
    normals = a["normals"].astype(np.float32).astype(np.float64)  # [36,3]
    tri_mat, mats, light_tris = a["tri_material"], a["materials"], a["light_tris"]
    nt = tris.shape[0]
    out.update(tris=a["tris"].astype(np.float32), normals=a["normals"].astype(np.float32), tri_material=tri_mat,
               materials=mats.view(np.uint8).reshape(mats.shape[0], -1), light_tris=light_tris)
    assert all(int(mats[m]["type"]) in (0, 1) for m in tri_mat), "the Cornell box is Lambert + emitter only"

    # per-triangle flat arrays exactly as shapes2.py:43-61 lays them out (p0, e1 = v1 - v0, e2 = v2 - v0)
    P0 = [tris[i, 0].copy() for i in range(nt)]
    E1 = [tris[i, 1] - tris[i, 0] for i in range(nt)]
    E2 = [tris[i, 2] - tris[i, 0] for i in range(nt)]
    scratch = [np.zeros(3) for _ in range(3)] + [np.zeros(1) for _ in range(4)]

    def tri_hit(i, o, d):
        """(hit, t, position) of triangle i through the reference's grouped numba kernel (n = 1)."""
        res = np.array([-1.0, 0.0])
        ray = Ray(o, d)
        r = ref_int.triangle_ray_intersection_grouping(ray, 1, scratch[0], scratch[1], scratch[2], P0[i], E1[i], E2[i],
                                                       scratch[3], scratch[4], scratch[5], scratch[6], res)
        if not r:
            return False, 0.0, None
        return True, r[0][0]["t"], r[0][0]["position"]

    def closest(o, d):
        best, bt, bp = -1, float(np.finfo(np.float32).max), None
        for i in range(nt):
            hit, t, pos = tri_hit(i, o, d)
            if hit and T_MIN <= t <= T_MAX and t < bt:  # strict <: lowest id wins exact ties (intersection.py:109)
                best, bt, bp = i, t, pos
        return best, bt, bp

    def occluded(o, d, tl):
        for i in range(nt):
            hit, t, _ = tri_hit(i, o, d)
            if hit and T_MIN <= t <= tl:
                return True
        return False

    # the reference's light primitive: a shapes2.Quad over the light's two triangles, so that ITS
    # sample_a_point draws the light point (vertices / faces replaced by the C-ABI triangles)
    U = Uniforms()
    light_prim = ref_shapes2.Quad.__new__(ref_shapes2.Quad)
    lv = tris[light_tris].reshape(-1, 3)
    light_prim.vertices = lv
    light_prim.faces = np.arange(lv.shape[0]).reshape(-1, 3)
    saved = (np.random.rand, ref_shapes2.random.randint, ref_shapes2.random.uniform)
    np.random.rand = lambda k: np.array([U.pop() for _ in range(k)])
    ref_shapes2.random.randint = lambda lo, hi: int(U.pop())
    ref_shapes2.random.uniform = lambda lo, hi: U.pop()

    cam = cam_mod.Camera(position=[0, 1, 6.8], looking_at=[0, 1, 0], up=[0, 1, 0], resolution=[W, H], fov=19.5)
    out["cam_iview"] = np.asarray(cam.iview, np.float64)

    def trace(pixel, sample, o, d):
        """core/tracing.py:116-155 (trace) + :92-108 (sample_direct_lighting), one path."""
        L = np.zeros(3)
        beta = np.ones(3)
        prim = -1
        for bounce in range(DEPTH):
            idx, t, p = closest(o, d)
            if bounce == 0:
                prim = idx
            if idx < 0:
                break                                                   # :141-142
            m = mats[tri_mat[idx]]
            n = normals[idx].copy()
            if int(m["type"]) == 1:                                     # :129-139
                d1 = float(np.dot(-d, n))
                if d1 > 0.0:
                    L += np.array(LIGHT_COLOR) * beta * (1.0 if bounce == 0 else d1)
                break
            if int(m["two_sided"]) and float(np.dot(n, -d)) < 0.0:      # shapes2.py:93-96
                n = -n
            r1 = rng4(pixel, sample, bounce, 1)
            r2 = rng4(pixel, sample, bounce, 2)
            U.set(u24(r1[0]), u24(r1[1]))
            wi = ref_samp.cosine_sample_hemisphere(n.copy())            # shapes2.py:98
            c = float(np.dot(n, wi))
            pdf = abs(c) * INV_PI                                       # shapes.py:108
            alb = np.array([float(x) for x in m["albedo"]])
            with np.errstate(divide="ignore", invalid="ignore"):
                nb = alb * max(c, 0.0) / pdf * INV_PI                   # :145
            if np.any(np.isnan(nb)):
                nb = alb * max(c, 0.0) / 1e-4 * INV_PI                  # :146-148
            beta = beta * nb                                            # :149
            # sample_direct_lighting(hit_pos, normal): :92-108
            face = (r1[2] * light_tris.shape[0]) >> 32
            U.set(face, u24(r2[0]), u24(r2[1]))
            p2 = light_prim.sample_a_point()                            # shapes2.py:72-79
            lt = int(light_tris[face])
            n2 = normals[lt]
            wv = p2 - p
            dist2 = float(np.dot(wv, wv))                               # sqrLength(p - p2)
            dist = math.sqrt(dist2)
            w = wv / dist
            w2 = -wv / dist
            if not occluded(p, w, dist * (1.0 - 1e-4)):                 # Q7
                dot1, dot2 = float(np.dot(n, w)), float(np.dot(n2, w2))
                if dot1 > 0.0 and dot2 > 0.0:
                    emissive = np.array([float(x) for x in mats[tri_mat[lt]]["albedo"]])  # BSDFLight.evaluate, bsdf.py:52-53
                    L += beta * emissive * dot1 * dot2 / dist2
            o, d = p, wi                                                # :153-154
        return L, prim

    rad = np.zeros((H, W, SPP, 3))
    prim = np.zeros((H, W, SPP), np.int32)
    rays = np.zeros((H, W, SPP, 6), np.float32)
    for j in range(H):
        for i in range(W):
            pixel = j * W + i
            for s in range(SPP):
                r0 = rng4(pixel, s, 0, 0)
                uv = np.array([(i + u24(r0[0])) / float(W), (j + u24(r0[1])) / float(H)])   # main.py:31-32
                ray = cam.generate_ray(uv)
                # rays cross the C ABI as f32 records
                o = np.asarray(ray.position, np.float64).astype(np.float32).astype(np.float64)
                d = np.asarray(ray.direction, np.float64).astype(np.float32).astype(np.float64)
                rays[j, i, s, :3], rays[j, i, s, 3:] = o, d
                rad[j, i, s], prim[j, i, s] = trace(pixel, s, o, d)
        print(f"row {j + 1}/{H}", flush=True)
    np.random.rand, ref_shapes2.random.randint, ref_shapes2.random.uniform = saved
    out.update(radiance=rad, prim_ids=prim, rays=rays,
               params=np.array([W, H, SPP, DEPTH, SEED], np.int64))

    # ------------------------------------------------------------------ (2) specular BSDFs from bsdf_taichi.py
    bt = importlib.import_module("pyr.core.bsdf_taichi")
    vt = importlib.import_module("pyr.mathematics.vec3_taichi")
    rng = np.random.default_rng(77)
    cos_g = np.concatenate([np.linspace(0.0, 1.0, 41), rng.uniform(0, 1, 23)])
    idx_g = np.array([1.0 / 1.5, 1.5, 1.0 / 1.33, 1.33, 1.0, 2.4])
    out["schlick_cos"], out["schlick_idx"] = cos_g, idx_g
    out["schlick_val"] = np.array([[bt.reflectance(float(c), float(e)) for e in idx_g] for c in cos_g])
    nvec = 200
    vs = rng.normal(size=(nvec, 3)); vs /= np.linalg.norm(vs, axis=1, keepdims=True)
    ns = rng.normal(size=(nvec, 3)); ns /= np.linalg.norm(ns, axis=1, keepdims=True)
    ns = np.where((np.sum(vs * ns, axis=1) > 0)[:, None], -ns, ns)  # normal faces the incoming side
    etas = rng.choice([1.0 / 1.5, 1.5, 1.0 / 1.33, 1.33], nvec)
    out["spec_v"], out["spec_n"], out["spec_eta"] = vs, ns, etas
    out["reflect_res"] = np.array([np.asarray(bt.reflect(Vec3(*v), Vec3(*n))) for v, n in zip(vs, ns)])
    out["refract_res"] = np.array([np.asarray(bt.refract(Vec3(*v), Vec3(*n), float(e))) for v, n, e in zip(vs, ns, etas)])
    # random_in_unit_sphere: three ti.random() draws in the order theta, v, r (vec3_taichi.py:33-39)
    us = rng.uniform(0, 1, (nvec, 3))
    sph = []
    for u in us:
        U.set(*u)
        py_random.random, keep = (lambda: U.pop()), py_random.random
        sph.append(np.asarray(vt.random_in_unit_sphere()))
        py_random.random = keep
    out["sphere_u"], out["sphere_res"] = us, np.array(sph)
    # Metal.scatter (bsdf_taichi.py:54-60): in_direction NOT normalised on purpose, roughness in [0, 1]
    rough = rng.choice([0.0, 0.05, 0.3, 1.0], nvec)
    scale = rng.uniform(0.5, 2.0, nvec)
    mres, mok = [], []
    for v, n, r, u, sc in zip(vs, ns, rough, us, scale):
        U.set(*u)
        py_random.random, keep = (lambda: U.pop()), py_random.random
        ok, _, wo, _ = bt.Metal.scatter(Vec3(*(v * sc)), Vec3(0, 0, 0), Vec3(*n), Vec3(1, 1, 1), float(r))
        py_random.random = keep
        mres.append(np.asarray(wo)); mok.append(bool(ok))
    out["metal_rough"], out["metal_scale"], out["metal_res"], out["metal_ok"] = rough, scale, np.array(mres), np.array(mok)
    # Dielectric.scatter (:71-86): one ti.random() for the Fresnel choice; front_facing picks 1/ior or ior
    iors = rng.choice([1.5, 1.33, 2.4], nvec)
    front = rng.integers(0, 2, nvec).astype(bool)
    uf = rng.uniform(0, 1, nvec)
    dres = []
    for v, n, ior, f, u, sc in zip(vs, ns, iors, front, uf, scale):
        U.set(u)
        py_random.random, keep = (lambda: U.pop()), py_random.random
        _, _, wo, _ = bt.Dielectric.scatter(Vec3(*(v * sc)), Vec3(0, 0, 0), Vec3(*n), Vec3(1, 1, 1), float(ior), bool(f))
        py_random.random = keep
        dres.append(np.asarray(wo))
    out["diel_ior"], out["diel_front"], out["diel_u"], out["diel_res"] = iors, front, uf, np.array(dres)

    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "radiance_golden.npz"), **out)
    print("wrote radiance_golden.npz", {k: np.asarray(v).shape for k, v in out.items()})
    print("mean radiance", rad.mean(axis=(0, 1, 2)), "hit fraction", float(np.mean(prim >= 0)))


if __name__ == "__main__":
    main()
