"""ctypes front-end of the CPU parity oracle (oracle/pt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs, never by the product
package ``pyrenderer_b200``.  See the header of pt_oracle.c for the
reference file:line each function restates and for how parity is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpt_oracle.so")
_SRC = os.path.join(_HERE, "pt_oracle.c")

MATERIAL_DTYPE = np.dtype(
    [("albedo", "<f4", 3), ("type", "<u4"), ("ior", "<f4"), ("roughness", "<f4"),
     ("two_sided", "<u4"), ("pad", "<u4"), ("emission", "<f4", 3), ("pad2", "<u4")])
assert MATERIAL_DTYPE.itemsize == 48


class Camera(C.Structure):
    _fields_ = [("iview", C.c_double * 16), ("sensor_w", C.c_double), ("sensor_h", C.c_double),
                ("focal", C.c_double), ("width", C.c_uint32), ("height", C.c_uint32),
                ("aperture", C.c_double)]


class RenderParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("spp_begin", C.c_uint32), ("spp_end", C.c_uint32),
                ("max_depth", C.c_uint32), ("rr_start", C.c_uint32),
                ("light_color", C.c_float * 3), ("tmin", C.c_float), ("tmax", C.c_float),
                ("flags", C.c_uint32)]


def build(force=False):
    """Compile pt_oracle.c with gcc (recipe == oracle/Makefile)."""
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    cmd = [cc, "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def num_threads():
    return int(lib().orc_num_threads())


def closest_hit(tris, rays, nthreads=0):
    """tris f32[nt,3,3]; rays f32[n,8] (o,tmin,d,tmax) -> ids i32[n], t,u,v f64[n]."""
    tris = _f32(tris, (-1, 9))
    rays = _f32(rays, (-1, 8))
    n = rays.shape[0]
    ids = np.empty(n, np.int32)
    t = np.empty(n, np.float64)
    u = np.empty(n, np.float64)
    v = np.empty(n, np.float64)
    rc = lib().orc_closest_hit(_p(tris, C.c_float), C.c_uint32(tris.shape[0]), _p(rays, C.c_float),
                               C.c_uint64(n), _p(ids, C.c_int32), _p(t, C.c_double),
                               _p(u, C.c_double), _p(v, C.c_double), C.c_int(nthreads))
    assert rc == 0
    return ids, t, u, v


def any_hit(tris, rays, nthreads=0):
    tris = _f32(tris, (-1, 9))
    rays = _f32(rays, (-1, 8))
    n = rays.shape[0]
    occ = np.empty(n, np.uint8)
    rc = lib().orc_any_hit(_p(tris, C.c_float), C.c_uint32(tris.shape[0]), _p(rays, C.c_float),
                           C.c_uint64(n), _p(occ, C.c_uint8), C.c_int(nthreads))
    assert rc == 0
    return occ


def all_hits(tris, rays, nthreads=0):
    tris = _f32(tris, (-1, 9))
    rays = _f32(rays, (-1, 8))
    n = rays.shape[0]
    cnt = np.empty(n, np.uint32)
    sums = np.empty(n, np.uint64)
    rc = lib().orc_all_hits(_p(tris, C.c_float), C.c_uint32(tris.shape[0]), _p(rays, C.c_float),
                            C.c_uint64(n), _p(cnt, C.c_uint32), _p(sums, C.c_uint64),
                            C.c_int(nthreads))
    assert rc == 0
    return cnt, sums


def mt_scalar(v0, v1, v2, o, d, bound_hi=3.4028234663852886e+38):
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (v0, v1, v2, o, d)]
    t = C.c_double(0.0)
    lib().orc_mt_scalar.restype = C.c_int
    hit = lib().orc_mt_scalar(*[_p(x, C.c_double) for x in a], C.c_double(bound_hi), C.byref(t))
    return bool(hit), t.value


def slab(t0, t1, pos, inv_dir, bmin, bmax):
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (pos, inv_dir, bmin, bmax)]
    out = C.c_double(0.0)
    hit = lib().orc_slab(C.c_double(t0), C.c_double(t1), *[_p(x, C.c_double) for x in a],
                         C.byref(out))
    return bool(hit), out.value


def concentric_sample_disk(u1, u2):
    out = np.zeros(2)
    lib().orc_concentric_sample_disk(C.c_double(u1), C.c_double(u2), _p(out, C.c_double))
    return out


def frame_z_to(n):
    n = np.ascontiguousarray(n, dtype=np.float64)
    r = [np.zeros(3) for _ in range(3)]
    lib().orc_frame_z_to(_p(n, C.c_double), *[_p(x, C.c_double) for x in r])
    return r


def cosine_sample_hemisphere(n, u1, u2):
    n = np.ascontiguousarray(n, dtype=np.float64)
    out = np.zeros(3)
    lib().orc_cosine_sample_hemisphere(_p(n, C.c_double), C.c_double(u1), C.c_double(u2),
                                       _p(out, C.c_double))
    return out


def schlick(cosine, idx):
    lib().orc_schlick.restype = C.c_double
    return float(lib().orc_schlick(C.c_double(cosine), C.c_double(idx)))


def reflect(v, n):
    v, n = (np.ascontiguousarray(x, dtype=np.float64) for x in (v, n))
    out = np.zeros(3)
    lib().orc_reflect(_p(v, C.c_double), _p(n, C.c_double), _p(out, C.c_double))
    return out


def refract(v, n, eta):
    v, n = (np.ascontiguousarray(x, dtype=np.float64) for x in (v, n))
    out = np.zeros(3)
    lib().orc_refract(_p(v, C.c_double), _p(n, C.c_double), C.c_double(eta), _p(out, C.c_double))
    return out


def in_unit_sphere(ua, ub, uc):
    out = np.zeros(3)
    lib().orc_in_unit_sphere(C.c_double(ua), C.c_double(ub), C.c_double(uc), _p(out, C.c_double))
    return out


def scatter_specular(kind, d, ns, front=True, ior=1.5, roughness=0.0, u=(0.5, 0.5, 0.5)):
    """(valid, wi un-normalised) of the oracle's specular sampler (kind: 2 mirror, 3 dielectric, 4 conductor)."""
    d, ns = (np.ascontiguousarray(x, dtype=np.float64) for x in (d, ns))
    out = np.zeros(3)
    ok = lib().orc_scatter_specular(C.c_uint32(kind), _p(d, C.c_double), _p(ns, C.c_double), C.c_int(1 if front else 0),
                                    C.c_double(ior), C.c_double(roughness), C.c_double(u[0]), C.c_double(u[1]),
                                    C.c_double(u[2]), _p(out, C.c_double))
    return bool(ok), out


def make_camera(iview, sensor_w, sensor_h, focal, width, height, aperture=0.0):
    cam = Camera()
    iv = np.ascontiguousarray(iview, dtype=np.float64).reshape(16)
    for i in range(16):
        cam.iview[i] = float(iv[i])
    cam.sensor_w, cam.sensor_h, cam.focal = float(sensor_w), float(sensor_h), float(focal)
    cam.width, cam.height = int(width), int(height)
    cam.aperture = float(aperture)
    return cam


def generate_ray(cam, u, v):
    o = np.zeros(3)
    d = np.zeros(3)
    lib().orc_generate_ray(C.byref(cam), C.c_double(u), C.c_double(v), _p(o, C.c_double),
                           _p(d, C.c_double))
    return o, d


def generate_rays(cam, seed=0, s0=0, s1=1, jitter=False, tmin=1e-5, tmax=99999.9):
    ns = s1 - s0
    rays = np.empty((cam.height, cam.width, ns, 8), np.float32)
    lib().orc_generate_rays(C.byref(cam), C.c_uint64(seed), C.c_uint32(s0), C.c_uint32(s1),
                            C.c_int(1 if jitter else 0), C.c_float(tmin), C.c_float(tmax),
                            _p(rays, C.c_float))
    return rays


def make_params(seed=1, spp_begin=0, spp_end=1, max_depth=5, rr_start=0xFFFFFFFF,
                light_color=(0.9, 0.85, 0.7), tmin=1e-5, tmax=99999.9, flags=0):
    P = RenderParams()
    P.seed, P.spp_begin, P.spp_end = int(seed), int(spp_begin), int(spp_end)
    P.max_depth, P.rr_start = int(max_depth), int(rr_start)
    for k in range(3):
        P.light_color[k] = float(light_color[k])
    P.tmin, P.tmax, P.flags = float(tmin), float(tmax), int(flags)
    return P


def render(tris, normals, tri_mat, mats, light_tris, cam, params, rows=None, want_ids=False,
           nthreads=0):
    """Returns accum f64[h,w,4] (rgb sums + count), prim ids (or None), (n_closest, n_shadow)."""
    tris = _f32(tris, (-1, 9))
    normals = _f32(normals, (-1, 3))
    tri_mat = np.ascontiguousarray(tri_mat, dtype=np.uint32)
    mats = np.ascontiguousarray(mats, dtype=MATERIAL_DTYPE)
    light_tris = np.ascontiguousarray(light_tris, dtype=np.uint32)
    H, W = cam.height, cam.width
    ns = params.spp_end - params.spp_begin
    accum = np.zeros((H, W, 4), np.float64)
    ids = np.full((H, W, ns), -2, np.int32) if want_ids else None
    stats = np.zeros(2, np.uint64)
    row0, row1 = (0, H) if rows is None else rows
    rc = lib().orc_render(_p(tris, C.c_float), _p(normals, C.c_float), C.c_uint32(tris.shape[0]),
                          _p(tri_mat, C.c_uint32), C.c_void_p(mats.ctypes.data),
                          C.c_uint32(mats.shape[0]), _p(light_tris, C.c_uint32),
                          C.c_uint32(light_tris.shape[0]), C.byref(cam), C.byref(params),
                          C.c_uint32(row0), C.c_uint32(row1), _p(accum, C.c_double),
                          _p(ids, C.c_int32), _p(stats, C.c_uint64), C.c_int(nthreads))
    assert rc == 0
    return accum, ids, (int(stats[0]), int(stats[1]))
