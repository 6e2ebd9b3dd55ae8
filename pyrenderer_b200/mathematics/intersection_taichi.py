"""``World`` -- the container main_taichi.py:41-44 builds (reference:
mathematics/intersection_taichi.py:189-291).

    world = World()
    for p in a_scene.primitives:
        world.add(p)
    world.commit()
    path_tracer = PathTracer(world, max_depth, image_width, image_height)

Same three calls here.  ``World`` is a ``Scene`` (core/scene.py) under the Taichi path's names:
``add`` = add_primitive, ``commit`` = upload + BVH build (and the reference's "There is no
lights!!!" assertion), ``hit_all`` = closest hit returning the reference's 8-tuple.  In the
reference ``hit_all`` is a ``@ti.func`` whose static unroll over the primitives makes code size
grow with the scene; here it is one closest-hit query on the device BVH, batched when
given arrays of rays, with the BSDF sample drawn on the host from ``rng``.
"""
import numpy as np

from .. import _abi
from ..core.scene import Scene
from .constants import MAX_F


def _cosine_sample_hemisphere(n, u1, u2):
    """mathematics/samplers.py: concentric disk -> cosine hemisphere around n (frame of
    mat4_taichi.py:9-60: x = normalize(n x Y), z = normalize(x x n))."""
    a, b = 2.0 * u1 - 1.0, 2.0 * u2 - 1.0
    if a == 0.0 and b == 0.0:
        dx = dy = 0.0
    elif abs(a) > abs(b):
        r, th = a, (np.pi / 4.0) * (b / a)
        dx, dy = r * np.cos(th), r * np.sin(th)
    else:
        r, th = b, np.pi / 2.0 - (np.pi / 4.0) * (a / b)
        dx, dy = r * np.cos(th), r * np.sin(th)
    dz = np.sqrt(max(0.0, 1.0 - dx * dx - dy * dy))
    if abs(abs(n[1]) - 1.0) == 0.0:
        x, z = np.array([1.0, 0.0, 0.0]), np.array([0.0, 0.0, 1.0]) * np.sign(n[1])
    else:
        x = np.cross(n, np.array([0.0, 1.0, 0.0]))
        x /= np.linalg.norm(x)
        z = np.cross(x, n)
        z /= np.linalg.norm(z)
    w = dx * x + dy * z + dz * n
    return w / np.linalg.norm(w)


class World(Scene):
    def __init__(self, seed=0):
        super().__init__()
        self.rng = np.random.default_rng(seed)

    def add(self, prim):
        self.add_primitive(prim)

    def commit(self, device=0, **bvh_options):
        """Commit should be called after all objects added (intersection_taichi.py:226-233)."""
        assert len(self.lights) > 0, "There is no lights!!!"
        return super().commit(device, **bvh_options)

    def sample_a_light(self):
        """(point, normal, emissive) of a random light (intersection_taichi.py:194-207,
        shapes.py:62-71)."""
        prim = self.lights[int(self.rng.integers(0, len(self.lights)))]
        k = int(self.rng.integers(0, prim.faces.shape[0]))
        tri = prim.triangles()[k]
        su, v = np.sqrt(self.rng.random()), self.rng.random()
        a, b = su * (1.0 - v), su * v
        return a * tri[0] + b * tri[1] + (1.0 - a - b) * tri[2], prim.normal_vectors[k], prim.bsdf.evaluate()

    def hit_all(self, ray_origin, ray_direction, t_min=1e-5, closest_so_far=99999.9):
        """-> (hit_anything, t, p, normal, emissive, attenuation, scattered_dir, pdf), the tuple of
        intersection_taichi.py:238-291.  One ray (3-vectors) or n rays ([n,3] arrays: every field
        gains a leading n)."""
        o = np.asarray(ray_origin, np.float64)
        single = o.ndim == 1
        o = np.atleast_2d(o)
        d = np.atleast_2d(np.asarray(ray_direction, np.float64))
        n = o.shape[0]
        rec = np.empty((n, 8), np.float32)
        rec[:, 0:3], rec[:, 3] = o, t_min
        rec[:, 4:7], rec[:, 7] = d, min(closest_so_far, MAX_F)
        hits = self.commit().trace_closest_host(rec, _abi.TRACE_EXACT)
        arrays = self.arrays()
        out = []
        for i in range(n):
            tri = int(hits["tri"][i])
            if tri < 0:
                z = np.zeros(3)
                out.append((False, float(closest_so_far), z, z, 0, z, z, 0.0))
                continue
            prim = self.primitives[int(arrays["tri_prim"][tri])]
            normal = arrays["normals"][tri].astype(np.float64)
            if prim.bsdf.sided == 0 and np.dot(normal, -d[i]) < 0.0:  # shapes.py:99-102
                normal = -normal
            t = float(hits["t"][i])
            wi = _cosine_sample_hemisphere(normal, self.rng.random(), self.rng.random())
            pdf = abs(float(np.dot(normal, wi))) / np.pi
            out.append((True, t, o[i] + t * d[i], normal, int(prim.bsdf.emitting_light),
                        np.asarray(prim.bsdf.evaluate(), np.float64), wi, pdf))
        if single:
            return out[0]
        return tuple(np.array([r[k] for r in out]) for k in range(8))
