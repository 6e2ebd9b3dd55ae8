"""Host-side ray record (reference: core/ray.py:5-17)."""
import numpy as np

from ..mathematics.constants import MAX_F


class Ray:
    def __init__(self, position, direction, depth=0):
        self.position = np.asarray(position, np.float64)
        self.direction = np.asarray(direction, np.float64)
        self.depth = depth
        self.bounds = np.array([0.0, MAX_F])
        with np.errstate(divide="ignore"):
            self.inv_direction = 1.0 / self.direction

    def reset_bounds(self):
        self.bounds = np.array([0.0, MAX_F])

    def __str__(self):
        return f"Ray: pos={self.position} dir={self.direction}"
