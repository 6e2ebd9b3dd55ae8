#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box does not
have it):  python tests/golden/make_golden.py

What runs is the reference's code, imported unmodified from /root/reference:
  mathematics/intersection.py   (scalar + grouped numba Moller-Trumbore)
  mathematics/fast_op.py, bbox.py (compute), samplers_debug.py, vec3.py
  mathematics/affine_transformation.py, mathematics/shapes2.py
  io_utils/read_utils_debug.py  (the loader test.py drives), core/ray.py
  core/camera.py                (Camera.generate_ray)

Third-party packages the reference imports but this image lacks are replaced
by *stubs defined below* -- they restate only the documented call semantics:
  trimesh.Trimesh(vertices, faces, process=False).apply_transform(M)
  pyrr.matrix44.create_look_at(eye, target, up)
  open3d / taichi / taichi_glsl  (import-only on the code paths used here)
Those stubs are themselves pinned by the reference's golden path
(test.py:38-57, data copied into GOLDEN_PATH below) and by the light mask of
media/cornell-box/TungstenRender.exr (exported here as light_mask rows/cols).
"""
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------
# stubs for absent third-party modules
# --------------------------------------------------------------------------
def _permissive(*a, **k):
    if len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0])):
        return a[0]  # acts as a decorator
    return _permissive


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _permissive


def _install_stubs():
    for name in ["open3d", "taichi", "taichi_glsl", "taichi_glsl.vector", "taichi_glsl.randgen",
                 "taichi_glsl.scalar", "skimage", "skimage.io"]:
        sys.modules[name] = _StubModule(name)

    tm = types.ModuleType("trimesh")

    class Trimesh:
        def __init__(self, vertices=None, faces=None, process=False):
            self.vertices = np.array(vertices, dtype=np.float64)
            self.faces = np.array(faces, dtype=np.int64)

        def apply_transform(self, matrix):
            m = np.asarray(matrix, dtype=np.float64)
            stack = np.column_stack((self.vertices, np.ones(len(self.vertices))))
            self.vertices = np.dot(m, stack.T).T[:, :3]
            if np.linalg.det(m[:3, :3]) < 0:
                self.faces = np.ascontiguousarray(np.fliplr(self.faces))
            return self

    tm.Trimesh = Trimesh
    sys.modules["trimesh"] = tm

    pyrr = types.ModuleType("pyrr")
    m44 = types.ModuleType("pyrr.matrix44")

    def create_look_at(eye, target, up, dtype=None):
        eye = np.asarray(eye, dtype=np.float64)
        target = np.asarray(target, dtype=np.float64)
        up = np.asarray(up, dtype=np.float64)
        f = target - eye
        f = f / np.linalg.norm(f)
        s = np.cross(f, up)
        s = s / np.linalg.norm(s)
        u = np.cross(s, f)
        u = u / np.linalg.norm(u)
        return np.array([[s[0], u[0], -f[0], 0.0],
                         [s[1], u[1], -f[1], 0.0],
                         [s[2], u[2], -f[2], 0.0],
                         [-np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye), 1.0]])

    m44.create_look_at = create_look_at
    m44.create_from_eulers = _permissive
    pyrr.matrix44 = m44
    sys.modules["pyrr"] = pyrr
    sys.modules["pyrr.matrix44"] = m44


# test.py:38-57 (hit, t, ro, rd, wi, albedo, n)
GOLDEN_PATH = [
    [1, 7.830270, [0.000000, 1.000000, 6.800000], [0.034281, 0.080880, -0.996134], [0.799974, -0.512694, 0.311747], [0.725000, 0.710000, 0.680000], [-0.000000, 0.000000, 1.000000]],
    [1, 0.914496, [0.268427, 1.633309, -1.000000], [0.799974, -0.512694, 0.311747], [-0.944430, -0.165854, -0.283803], [0.140000, 0.450000, 0.091000], [-1.000000, -0.000000, -0.000000]],
    [1, 1.004540, [1.000000, 1.164452, -0.714908], [-0.944430, -0.165854, -0.283803], [-0.539883, -0.658663, 0.524109], [0.725000, 0.710000, 0.680000], [-0.000000, 0.000000, 1.000000]],
    [1, 0.766014, [0.051282, 0.997845, -1.000000], [-0.539883, -0.658663, 0.524109], [-0.039377, 0.827128, -0.560632], [0.725000, 0.710000, 0.680000], [-0.328669, 0.000000, -0.944445]],
    [1, 0.716111, [-0.362276, 0.493300, -0.598526], [-0.039377, 0.827128, -0.560632], [0.688579, -0.690442, 0.221697], [0.725000, 0.710000, 0.680000], [-0.000000, 0.000000, 1.000000]],
    [1, 1.572349, [-0.390474, 1.085616, -1.000000], [0.688579, -0.690442, 0.221697], [0.631949, 0.704407, -0.323189], [0.725000, 0.710000, 0.680000], [-0.000000, 1.000000, -0.000000]],
    [1, 0.487045, [0.692212, 0.000000, -0.651414], [0.631949, 0.704407, -0.323189], [-0.074693, -0.669229, 0.739292], [0.140000, 0.450000, 0.091000], [-1.000000, -0.000000, -0.000000]],
    [1, 0.512646, [1.000000, 0.343078, -0.808822], [-0.074693, -0.669229, 0.739292], [0.324842, 0.923928, -0.202077], [0.725000, 0.710000, 0.680000], [-0.000000, 1.000000, -0.000000]],
    [1, 0.117876, [0.961709, -0.000000, -0.429826], [0.324842, 0.923928, -0.202077], [-0.052710, -0.073551, 0.995898], [0.140000, 0.450000, 0.091000], [-1.000000, -0.000000, -0.000000]],
]


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    # namespace mount so that modules with beyond-top-level relative imports load
    mount = tempfile.mkdtemp()
    os.symlink(REF, os.path.join(mount, "pyr"))
    sys.path.insert(0, mount)
    os.chdir(REF)

    import json
    from core.ray import Ray
    from mathematics import bbox as ref_bbox
    from mathematics import intersection as ref_int
    from mathematics import samplers_debug as ref_samp
    from mathematics.affine_transformation import make_transformation_matrix
    from io_utils.read_utils_debug import read_scene

    rng = np.random.default_rng(20261018)
    out = {}

    # ---- (1) loader: transforms + geometry through the reference's debug loader
    with open("media/cornell-box/scene.json") as f:
        data = json.load(f)
    out["transforms"] = np.stack([make_transformation_matrix(p["transform"])
                                  for p in data["primitives"]]).astype(np.float64)
    sc = read_scene("media/cornell-box/scene.json")
    out["scene_vertices"] = np.asarray(sc.vertices, np.float64)
    out["scene_faces"] = np.asarray(sc.faces, np.int64)
    out["scene_normals"] = np.vstack([p.normal_vectors for p in sc.primitives]).astype(np.float64)
    out["scene_albedo"] = np.vstack([np.tile(np.asarray(p.bsdf, np.float64).reshape(-1)[:3]
                                             if np.ndim(p.bsdf) else np.full(3, float(p.bsdf)),
                                             (p.faces.shape[0], 1)) for p in sc.primitives])
    out["prim_bounds"] = np.stack([np.stack([p.bounds.min_coord, p.bounds.max_coord])
                                   for p in sc.primitives]).astype(np.float64)

    # ---- (2) golden path of test.py replayed through the reference Scene.hit
    gp = np.array([[r[1]] + r[2] + r[3] + r[4] + r[5] + r[6] for r in GOLDEN_PATH], np.float64)
    out["golden_path"] = gp  # t, ro3, rd3, wi3, albedo3, n3
    rep = []
    for r in GOLDEN_PATH:
        np.random.seed(2)
        res = sc.hit(np.array(r[2]), np.array(r[3]))
        rep.append([res["t"]] + list(res["normal"]) + list(np.asarray(res["bsdf"], np.float64)))
    out["golden_path_replay"] = np.array(rep, np.float64)  # t, n3, albedo3 from the reference

    # ---- (3) closest hit through the reference's grouped numba kernel
    def flat_arrays(tri):  # tri f64[n,3,3] -> p0,e1,e2 flattened like shapes2.py:59-61
        p0 = tri[:, 0, :].reshape(-1).copy()
        e1 = (tri[:, 1, :] - tri[:, 0, :]).reshape(-1).copy()
        e2 = (tri[:, 2, :] - tri[:, 0, :]).reshape(-1).copy()
        return p0, e1, e2

    def ref_grouped(tri, o, d):
        n = tri.shape[0]
        p0, e1, e2 = flat_arrays(tri)
        s = np.zeros(3 * n); q = np.zeros(3 * n); r = np.zeros(3 * n)
        a = np.zeros(n); e2r = np.zeros(n); sq = np.zeros(n); rdr = np.zeros(n)
        res = np.zeros(2 * n)
        res[0::2] = -1.0
        ray = Ray(np.asarray(o, np.float64), np.asarray(d, np.float64))
        results = ref_int.triangle_ray_intersection_grouping(ray, n, s, q, r, p0, e1, e2, a, e2r,
                                                             sq, rdr, res)
        if not results:
            return -1, 0.0
        ret, idx = min(results, key=lambda du: du[0]["t"])
        return idx, ret["t"]

    def f32(a):
        return np.asarray(a, np.float32)

    cornell_tri = f32(sc.vertices[sc.faces]).astype(np.float64)  # f32-quantised, [36,3,3]
    cases = {}
    # Cornell: camera-ish rays + interior rays
    n1 = 1500
    o = np.tile(np.array([0.0, 1.0, 6.8]), (n1, 1))
    tgt = np.stack([rng.uniform(-1.1, 1.1, n1), rng.uniform(-0.1, 2.1, n1), np.zeros(n1)], 1)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o2 = np.stack([rng.uniform(-0.99, 0.99, n1), rng.uniform(0.01, 1.97, n1),
                   rng.uniform(-0.99, 0.99, n1)], 1)
    d2 = rng.normal(size=(n1, 3))
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    cases["cornell"] = (cornell_tri, np.vstack([o, o2]), np.vstack([d, d2]))
    # random 64-triangle soup (B.4)
    c = rng.uniform(0, 1, (64, 1, 3))
    e = rng.uniform(-0.3, 0.3, (64, 2, 3))
    soup = f32(np.concatenate([c, c + e[:, :1], c + e[:, 1:]], 1)).astype(np.float64)
    o3 = rng.uniform(0, 1, (2500, 3))
    d3 = rng.normal(size=(2500, 3))
    d3 /= np.linalg.norm(d3, axis=1, keepdims=True)
    cases["soup64"] = (soup, o3, d3)
    for name, (tri, oo, dd) in cases.items():
        oo = f32(oo).astype(np.float64)
        dd = f32(dd).astype(np.float64)
        ids = np.empty(len(oo), np.int32)
        ts = np.empty(len(oo), np.float64)
        for i in range(len(oo)):
            ids[i], ts[i] = ref_grouped(tri, oo[i], dd[i])
        out[f"ch_{name}_tris"] = f32(tri)
        out[f"ch_{name}_o"] = f32(oo)
        out[f"ch_{name}_d"] = f32(dd)
        out[f"ch_{name}_ids"] = ids
        out[f"ch_{name}_t"] = ts

    # ---- (3b) scalar kernel decisions (intersection.py:7-39), 64 tris x 300 rays
    dec = np.zeros((300, 64), np.uint8)
    tt = np.zeros((300, 64), np.float64)
    oo = f32(o3[:300]).astype(np.float64)
    dd = f32(d3[:300]).astype(np.float64)
    for i in range(300):
        for k in range(64):
            ray = Ray(oo[i], dd[i])
            res = ref_int.triangle_ray_intersection(soup[k], ray)
            dec[i, k] = 1 if res["hit"] else 0
            tt[i, k] = res["t"]
    out["scalar_dec"] = dec
    out["scalar_t"] = tt

    # ---- (4) slab test bbox.compute
    nb = 1000
    bmin = rng.uniform(-1, 1, (nb, 3))
    bmax = bmin + rng.uniform(0, 1, (nb, 3))
    bo = rng.uniform(-2, 2, (nb, 3))
    bd = rng.normal(size=(nb, 3))
    bd /= np.linalg.norm(bd, axis=1, keepdims=True)
    bd[::17, 0] = 0.0  # axis-aligned rays -> inf inverse direction (core/ray.py:11)
    with np.errstate(divide="ignore"):
        binv = 1.0 / bd
    hold = np.zeros(2, np.float64)
    bres = np.zeros((nb, 2))
    for i in range(nb):
        hold[:] = 0
        ref_bbox.compute(0.0, float(np.finfo(np.float32).max), bo[i], binv[i], bmin[i], bmax[i], hold)
        bres[i] = hold
    out.update(slab_bmin=bmin, slab_bmax=bmax, slab_o=bo, slab_inv=binv, slab_res=bres)

    # ---- (5) samplers
    us = rng.uniform(0, 1, (1000, 2))
    us[0] = (0.5, 0.5)
    us[1] = (0.25, 0.75)
    out["disk_u"] = us
    out["disk_res"] = np.stack([ref_samp.concentric_sample_disk(u) for u in us])
    ns = rng.normal(size=(200, 3))
    ns /= np.linalg.norm(ns, axis=1, keepdims=True)
    ns[0] = (0, 1, 0); ns[1] = (0, -1, 0); ns[2] = (1, 0, 0); ns[3] = (0, 0, -1)
    fr = []
    for n in ns:
        r1, r2, r3, r4 = ref_samp.rotate_z_to(n.copy())
        fr.append(np.stack([r1[:3], r2[:3], r3[:3]]))
    out["frame_n"] = ns
    out["frame_res"] = np.stack(fr)
    cs = []
    saved = np.random.rand
    for i, n in enumerate(ns):
        u = us[i]
        np.random.rand = lambda k, _u=u: _u.copy()
        cs.append(ref_samp.cosine_sample_hemisphere(n.copy()))
    np.random.rand = saved
    out["cos_res"] = np.stack(cs)

    # ---- (6) camera (core/camera.py) via the namespace mount
    import importlib
    cam_mod = importlib.import_module("pyr.core.camera")
    cams = [dict(position=[0, 1, 6.8], looking_at=[0, 1, 0], up=[0, 1, 0], resolution=[1024, 1024], fov=19.5),
            dict(position=[2.6, 2.1, 3.4], looking_at=[0.5, 0.5, 0.5], up=[0, 1, 0], resolution=[640, 480], fov=35.0)]
    uv = rng.uniform(0, 1, (256, 2))
    uv[0] = (0.5, 0.5); uv[1] = (0, 0); uv[2] = (1, 1)
    out["cam_uv"] = uv
    for ci, kw in enumerate(cams):
        cam = cam_mod.Camera(**kw)
        out[f"cam{ci}_iview"] = np.asarray(cam.iview, np.float64)
        rr = []
        for p in uv:
            ray = cam.generate_ray(p.copy())
            rr.append(np.concatenate([ray.position, ray.direction]))
        out[f"cam{ci}_rays"] = np.array(rr, np.float64)

    # ---- (7) light mask of the Tungsten EXR (geometric fixture)
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    exr = cv2.imread(os.path.join(REF, "media/cornell-box/TungstenRender.exr"), cv2.IMREAD_UNCHANGED)
    rgb = exr[:, :, ::-1]
    mask = np.all(np.abs(rgb - np.array([17, 12, 4], np.float32)) < 1e-3, axis=2)
    ys, xs = np.nonzero(mask)
    out["exr_light_rows"] = ys.astype(np.int16)
    out["exr_light_cols"] = xs.astype(np.int16)

    np.savez_compressed(os.path.join(OUT, "reference_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "reference_golden.npz"),
          {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
