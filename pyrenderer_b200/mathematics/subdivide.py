"""Uniform 4-way midpoint subdivision of triangle arrays.

Used to construct BASELINE config 5 ("subdivided Cornell box, ~5M triangles") from the 36
Cornell triangles.  Midpoints are computed as (a + b) * 0.5, which is symmetric in a and b, so
the two triangles sharing an edge get bit-identical new vertices (no cracks).
"""
import numpy as np


def subdivide(tris, levels):
    """tris [n,3,3] -> ([n*4**levels,3,3], parent index [n*4**levels]); winding preserved."""
    tris = np.asarray(tris)
    parent = np.arange(tris.shape[0])
    for _ in range(int(levels)):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab, bc, ca = (a + b) * 0.5, (b + c) * 0.5, (c + a) * 0.5
        tris = np.stack([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1),
                         np.stack([ca, bc, c], 1), np.stack([ab, bc, ca], 1)], 1).reshape(-1, 3, 3)
        parent = np.repeat(parent, 4)
    return tris, parent


def subdivide_scene_arrays(arrays, levels_per_triangle):
    """Scene.arrays() dict -> same dict with triangle i split ``levels_per_triangle[i]`` times.
    Per-triangle attributes (normal, material) are inherited; light_tris is rebuilt."""
    lv = np.asarray(levels_per_triangle, int)
    out_t, out_p = [], []
    for k in np.unique(lv):
        idx = np.nonzero(lv == k)[0]
        t, p = subdivide(arrays["tris"][idx].astype(np.float32), k)
        out_t.append(t)
        out_p.append(idx[p])
    tris = np.concatenate(out_t)
    par = np.concatenate(out_p)
    order = np.argsort(par, kind="stable")  # keep the original global-id order between groups
    tris, par = tris[order], par[order]
    lights = np.nonzero(np.isin(par, arrays["light_tris"]))[0].astype(np.uint32)
    out = dict(arrays)
    out.update(tris=np.ascontiguousarray(tris, np.float32), normals=arrays["normals"][par],
               tri_material=arrays["tri_material"][par], light_tris=lights)
    if "tri_prim" in arrays:
        out["tri_prim"] = arrays["tri_prim"][par]
    return out
