"""Pinhole camera (reference: core/camera.py:14-72).

``generate_ray`` is the host-side, one-ray-at-a-time entry the reference's
main.py uses; the render path generates rays on the device
(csrc/wavefront.cu::raygen_kernel) from :meth:`Camera.device_record`, in f64
with the same operation order, so both produce the same f32 ray records.
``look_at`` restates pyrr.matrix44.create_look_at (row-vector convention).
"""
from math import radians, tan

from random import random

import numpy as np

from .ray import Ray


def look_at(eye, target, up):
    eye, target, up = (np.asarray(a, np.float64) for a in (eye, target, up))
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


class Camera:
    def __init__(self, position, looking_at, up, resolution, fov=90, aperture=0, focal_dist=1.0):
        self.position = np.array(position, np.float64)
        self.looking_at = np.array(looking_at, np.float64)
        self.up = np.array(up, np.float64)
        self.view = look_at(self.position, self.looking_at, self.up)
        self.iview = np.linalg.inv(self.view)
        self.resolution = list(resolution)
        self.aperture = aperture
        self.focal_dist = focal_dist
        self.fov = fov

    @property
    def aspect_ratio(self):
        return self.resolution[0] / self.resolution[1] * 1.0

    def get_resolution(self):
        return self.resolution

    def sensor(self):
        h = tan(radians(self.fov) / 2) * self.focal_dist
        return h * self.aspect_ratio, h

    def device_record(self):
        """(iview f64[16] row-major, sensor_w, sensor_h, focal, width, height)."""
        sw, sh = self.sensor()
        return (np.ascontiguousarray(self.iview, np.float64).reshape(16), sw, sh,
                float(self.focal_dist), int(self.resolution[0]), int(self.resolution[1]))

    def generate_ray(self, screen_coordinates):
        sw, sh = self.sensor()
        c = np.asarray(screen_coordinates, np.float64) - 0.5
        d_cam = np.ones(4, np.float32)  # the reference rounds this vector to f32
        d_cam[:3] = [c[0] * sw / 0.5, c[1] * sh / 0.5, -self.focal_dist]
        h = d_cam.astype(np.float64)
        m = self.iview
        d_w = ((h[0] * m[0] + h[1] * m[1]) + h[2] * m[2]) + h[3] * m[3]
        o_w = m[3].copy()
        if self.aperture > 0:  # thin lens: origin on a square lens in camera space (camera.py:63-65)
            o_cam = np.zeros(4, np.float32)
            o_cam[0] = self.aperture * random() - self.aperture / 2.0
            o_cam[1] = self.aperture * random() - self.aperture / 2.0
            o_cam[3] = 1.0
            g = o_cam.astype(np.float64)
            o_w = (g[0] * m[0] + g[1] * m[1]) + g[3] * m[3]
        d = (d_w - o_w)[:3]
        d = d / np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        return Ray(o_w[:3], d, 8)
