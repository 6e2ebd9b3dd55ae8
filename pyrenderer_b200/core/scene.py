"""Scene container (reference: core/scene.py:11-73).

Keeps the reference's protocol -- ``primitives``, merged ``vertices`` / ``faces``
whose face order is the GLOBAL triangle id, ``lights``, ``add_primitive``,
``build_bvh_tree``, ``hit`` / ``hit_faster``, ``sample_light`` -- but the
acceleration structure and every intersection live on the GPU behind the C ABI
(``commit`` uploads the triangle arrays and runs the LBVH build).
"""
import random

import numpy as np

from .. import _abi
from ..mathematics.constants import MAX_F


class Scene:
    def __init__(self):
        self.primitives = []
        self.vertices = None
        self.faces = None
        self.lights = []
        self._ctx = None
        self._arrays = None

    @classmethod
    def from_arrays(cls, arrays):
        """Scene over ready-made flat arrays (the dict :meth:`arrays` returns), e.g. a subdivided or
        procedurally generated mesh; the primitive list stays empty."""
        s = cls()
        s._arrays = dict(arrays)
        return s

    # -- reference protocol -------------------------------------------------
    def add_primitive(self, prim):
        prim.id = len(self.primitives)
        self.primitives.append(prim)
        if prim.bsdf.emitting_light:
            self.lights.append(prim)
        if self.vertices is None:
            self.vertices = prim.vertices
            self.faces = prim.faces
        else:
            offset = self.vertices.shape[0]
            self.vertices = np.vstack([self.vertices, prim.vertices])
            self.faces = np.vstack([self.faces, prim.faces + offset])
        self._ctx = None
        self._arrays = None

    def sample_light(self):
        if not self.lights:
            print("[WARNING] no lights found")
            return None
        prim = random.choice(self.lights)
        tri = prim.triangles()[random.randint(0, prim.faces.shape[0] - 1)]
        u = random.uniform(0, 1) ** 0.5
        v = random.uniform(0, 1)
        a, b = u * (1 - v), u * v
        return a * tri[0] + b * tri[1] + (1.0 - a - b) * tri[2]

    def build_bvh_tree(self, device=0):
        self.commit(device)

    def hit(self, ray):
        """Closest hit of one host-side Ray -> the reference's result dict."""
        ctx = self.commit()
        rec = np.array([[*ray.position, max(ray.bounds[0], 0.0), *ray.direction,
                         min(ray.bounds[1], MAX_F)]], np.float32)
        h = ctx.trace_closest_host(rec, _abi.TRACE_EXACT)[0]
        if h["tri"] < 0:
            return {"origin": ray.position, "hit": False, "t": MAX_F}
        tri = int(h["tri"])
        arrays = self.arrays()
        prim = self.primitives[int(arrays["tri_prim"][tri])]
        normal = arrays["normals"][tri].astype(np.float64)
        if prim.bsdf.sided == 0 and np.dot(normal, -ray.direction) < 0.0:
            normal = -normal
        t = float(h["t"])
        ray.bounds[1] = t
        return {"origin": ray.position, "hit": True, "t": t, "position": ray.position + t * ray.direction,
                "bsdf": prim.bsdf, "normal": normal, "triangle": tri}

    hit_faster = hit

    # -- device hand-off ----------------------------------------------------
    def arrays(self):
        """Flat arrays in global-triangle-id order, as the C ABI takes them."""
        if self._arrays is not None:
            return self._arrays
        tris, normals, tri_mat, tri_prim, mats, lights = [], [], [], [], [], []
        mat_index = {}
        nt = 0
        for pi, prim in enumerate(self.primitives):
            emission = tuple(float(x) for x in getattr(prim, "emission", (0.0, 0.0, 0.0)))
            key = (id(prim.bsdf), emission)  # Tungsten puts `emission` on the primitive, not the bsdf
            if key not in mat_index:
                mat_index[key] = len(mats)
                mats.append(prim.bsdf.record() + (emission,))
            k = prim.faces.shape[0]
            tris.append(prim.triangles())
            normals.append(prim.normal_vectors)
            tri_mat += [mat_index[key]] * k
            tri_prim += [pi] * k
            if prim.bsdf.emitting_light:
                lights += list(range(nt, nt + k))
            nt += k
        m = np.zeros(len(mats), _abi.MATERIAL_DTYPE)
        for i, (alb, kind, ior, rough, two, emission) in enumerate(mats):
            m[i] = (alb, kind, ior, rough, two, 0, emission, 0)
        self._arrays = {
            "tris": np.concatenate(tris).astype(np.float32) if tris else np.zeros((0, 3, 3), np.float32),
            "normals": np.concatenate(normals).astype(np.float32) if normals else np.zeros((0, 3), np.float32),
            "tri_material": np.asarray(tri_mat, np.uint32),
            "tri_prim": np.asarray(tri_prim, np.uint32),
            "materials": m,
            "light_tris": np.asarray(lights, np.uint32),
        }
        return self._arrays

    def commit(self, device=0, **bvh_options):
        """Upload + LBVH build; returns the device context (cached)."""
        if self._ctx is not None and self._ctx.device == device and not bvh_options:
            return self._ctx
        a = self.arrays()
        ctx = _abi.Context(device)
        ctx.set_triangles(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"])
        self.bvh_stats = ctx.build_bvh(**bvh_options)
        self._ctx = ctx
        return ctx
