"""Scene primitives as plain triangle lists.

Reference: mathematics/shapes.py:15-243 (Quad, Cube) and shapes2.py.  There a
primitive owns Taichi fields and a per-primitive ``hit``; here a primitive is
only geometry + a BSDF handle -- all intersection happens on the GPU over the
scene's merged triangle array, so each primitive just has to produce

    vertices        f64[nv,3]  world space
    faces           int[nf,3]
    normal_vectors  f64[nf,3]  geometric normals with the reference's sign
                               (quad: -normalize(e1 x e2), shapes.py:43-47;
                                cube: +normalize(e1 x e2), shapes.py:172-176)

The transform is applied the way ``trimesh.Trimesh.apply_transform`` does:
``v' = (M [v,1])[:3]`` and the winding is reversed when det(M[:3,:3]) < 0.
"""
import numpy as np

from .bbox import BBox

_QUAD_V = np.array([[-0.5, 0, -0.5], [0.5, 0, -0.5], [0.5, 0, 0.5], [-0.5, 0, 0.5]], np.float64)
_QUAD_F = np.array([[0, 1, 2], [2, 3, 0]], np.int64)

# 24 vertices (4 per side) / 12 faces, listed side by side: -y, +y, -z, +z, -x, +x
_CUBE_SIDES = [
    [(-1, -1, -1), (-1, -1, 1), (1, -1, 1), (1, -1, -1)],
    [(-1, 1, 1), (-1, 1, -1), (1, 1, -1), (1, 1, 1)],
    [(-1, 1, -1), (-1, -1, -1), (1, -1, -1), (1, 1, -1)],
    [(1, 1, 1), (1, -1, 1), (-1, -1, 1), (-1, 1, 1)],
    [(-1, 1, 1), (-1, -1, 1), (-1, -1, -1), (-1, 1, -1)],
    [(1, 1, -1), (1, -1, -1), (1, -1, 1), (1, 1, 1)],
]
_CUBE_V = 0.5 * np.array([c for side in _CUBE_SIDES for c in side], np.float64)
_CUBE_F = np.array([f for s in range(6) for f in ((4 * s + 2, 4 * s + 1, 4 * s),
                                                   (4 * s, 4 * s + 3, 4 * s + 2))], np.int64)


def apply_transform(vertices, faces, matrix):
    m = np.asarray(matrix, np.float64)
    homo = np.column_stack((vertices, np.ones(len(vertices))))
    out = np.dot(m, homo.T).T[:, :3]
    if np.linalg.det(m[:3, :3]) < 0:
        faces = np.ascontiguousarray(np.fliplr(faces))
    return out, faces


def _unit(v):
    return v / np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


class _TrianglePrimitive:
    normal_sign = 1.0

    def __init__(self, vertices, faces, trans_mat, bsdf):
        self.id = -1
        self.trans_mat = trans_mat
        self.vertices, self.faces = apply_transform(vertices, faces, trans_mat)
        self.bsdf = bsdf
        self.bounds = BBox()
        self.bounds.from_vertices(self.vertices)
        self.center = self.bounds.center()
        tri = self.vertices[self.faces]
        e1 = tri[:, 1] - tri[:, 0]
        e2 = tri[:, 2] - tri[:, 0]
        self.normal_vectors = np.stack([self.normal_sign * _unit(np.cross(a, b))
                                        for a, b in zip(e1, e2)])

    @property
    def bounding_box(self):
        return self.bounds.min_coord, self.bounds.max_coord

    def triangles(self):
        """f64[nf,3,3] world-space corner positions."""
        return self.vertices[self.faces]


class Quad(_TrianglePrimitive):
    normal_sign = -1.0

    def __init__(self, trans_mat, bsdf):
        super().__init__(_QUAD_V, _QUAD_F, trans_mat, bsdf)


class Cube(_TrianglePrimitive):
    normal_sign = 1.0

    def __init__(self, trans_mat, bsdf):
        super().__init__(_CUBE_V, _CUBE_F, trans_mat, bsdf)


class TriangleMesh(_TrianglePrimitive):
    """Arbitrary indexed mesh (OBJ / Tungsten "mesh")."""
    normal_sign = 1.0

    def __init__(self, vertices, faces, trans_mat, bsdf):
        super().__init__(np.asarray(vertices, np.float64), np.asarray(faces, np.int64),
                         trans_mat, bsdf)
