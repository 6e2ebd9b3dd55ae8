"""Tungsten scene.json -> (Scene, Camera)   (reference: io_utils/read_tungsten.py:15-46).

Same behaviour as the reference loader: ``quad`` and ``cube`` primitives,
``lambert`` and ``null`` BSDFs, unknown primitive types are skipped with a
``[WARNING]`` line, unknown BSDF types raise NotImplementedError.  Additions
(SURVEY 8f rank 1): ``mesh`` primitives referencing a Wavefront OBJ file, and
the mirror / dielectric / conductor BSDFs.
"""
import json
import os

import numpy as np

from ..core.bsdf import BSDF
from ..core.camera import Camera
from ..core.scene import Scene
from ..mathematics.affine_transformation import make_transformation_matrix
from ..mathematics.shapes import Cube, Quad, TriangleMesh

PRIM_TYPES = {
    "quad": Quad,
    "cube": Cube,
}


def read_obj(path):
    """Minimal OBJ reader: ``v`` and ``f`` records (``f a//n b//n c//n`` accepted),
    polygons are fanned.  Face order in the file is the triangle order."""
    verts, faces = [], []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(x) for x in tok[1:4]])
            elif tok[0] == "f":
                idx = [int(t.split("/")[0]) for t in tok[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(verts, np.float64), np.asarray(faces, np.int64)


def process_primitives(data, base_dir="."):
    a_scene = Scene()
    cam = data["camera"]
    a_camera = Camera(cam["transform"]["position"], cam["transform"]["look_at"],
                      cam["transform"]["up"], cam["resolution"], fov=cam["fov"])
    name2bsdf = {}
    for info_bsdf in data["bsdfs"]:
        name2bsdf[info_bsdf["name"]] = BSDF(info_bsdf).get_distribution()
    for info in data["primitives"]:
        trans_mat = make_transformation_matrix(info.get("transform", {}))
        kind = info["type"]
        if kind == "mesh" and str(info.get("file", "")).endswith(".obj"):
            v, f = read_obj(os.path.join(base_dir, info["file"]))
            prim = TriangleMesh(v, f, trans_mat, name2bsdf[info["bsdf"]])
        elif kind in PRIM_TYPES:
            prim = PRIM_TYPES[kind](trans_mat, name2bsdf[info["bsdf"]])
        else:
            print(f"[WARNING] {kind} not implemented")
            continue
        # radiance of an emitter (scene.json:234-238; ignored by the reference, used by the
        # physically-based mode)
        e = np.asarray(info.get("emission", 0.0), np.float64).reshape(-1)
        prim.emission = np.full(3, e[0]) if e.size == 1 else e[:3].copy()
        a_scene.add_primitive(prim)
    return a_scene, a_camera


def read_file(filename):
    with open(filename) as json_file:
        data = json.load(json_file)
    return process_primitives(data, os.path.dirname(os.path.abspath(filename)))
