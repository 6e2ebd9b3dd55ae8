"""Tungsten `transform` block -> 4x4 matrix.

Mirrors reference mathematics/affine_transformation.py:39-55: the result is
``T(position) . Rx . Ry . Rz . S(scale)`` in the column-vector convention,
angles in degrees, zero angles skipped.  Two details of the reference are
kept because they move vertices by ~1e-8: translation and scale entries pass
through float32 (the reference builds them in ``np.identity(4, float32)``),
and single-axis rotations come from the unit quaternion of the half angle
(what scipy's ``Rotation.from_euler(axis, deg).as_matrix()`` evaluates), not
from cos/sin of the full angle.  Checked against the reference's own output
in tests/test_oracle_golden.py::test_transforms.
"""
from math import cos, radians, sin

import numpy as np


def _axis_rotation(axis, degrees):
    half = radians(degrees) / 2.0
    q = [0.0, 0.0, 0.0, cos(half)]
    q[axis] = sin(half)
    x, y, z, w = q
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    return np.array([
        [x2 - y2 - z2 + w2, 2.0 * (xy - zw), 2.0 * (xz + yw)],
        [2.0 * (xy + zw), -x2 + y2 - z2 + w2, 2.0 * (yz - xw)],
        [2.0 * (xz - yw), 2.0 * (yz + xw), -x2 - y2 + z2 + w2]])


def make_rotation_matrix(degrees, homo=True):
    rot = np.identity(3)
    for axis, angle in enumerate(degrees):
        if angle != 0:
            rot = rot @ _axis_rotation(axis, angle)
    if not homo:
        return rot
    out = np.identity(4)
    out[:3, :3] = rot
    return out


def make_translation_matrix(moves):
    out = np.identity(4)
    out[:3, 3] = np.asarray(moves, np.float32)
    return out


def make_scale_matrix(scales):
    return np.diag(np.append(np.asarray(scales, np.float32).astype(np.float64), 1.0))


def make_transformation_matrix(transforms):
    """``{'position': [...], 'rotation': [...], 'scale': [...]}`` -> f64[4,4]."""
    out = np.identity(4)
    if "position" in transforms:
        out = out @ make_translation_matrix(transforms["position"])
    if "rotation" in transforms:
        out = out @ make_rotation_matrix(transforms["rotation"])
    if "scale" in transforms:
        out = out @ make_scale_matrix(transforms["scale"])
    return out
