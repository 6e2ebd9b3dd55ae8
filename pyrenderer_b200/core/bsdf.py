"""Tungsten BSDF blocks -> material records for the device.

Reference: core/bsdf.py:18-91 knows ``lambert`` (BSDFLambertian, two-sided,
``sided = 0``) and ``null`` (BSDFLight: the emitter, one-sided, ``sided = 1``)
and raises NotImplementedError otherwise.  Mirror / dielectric / conductor
follow the semantics of core/bsdf_taichi.py:45-86 (Metal with roughness 0,
Dielectric, Metal).  Sampling and evaluation are device functions
(csrc/bsdf.cuh); these classes only carry parameters across the C ABI.
"""
import numpy as np

LAMBERT, EMITTER, MIRROR, DIELECTRIC, CONDUCTOR = 0, 1, 2, 3, 4


def _rgb(value):
    a = np.asarray(value, np.float64).reshape(-1)
    return np.full(3, a[0]) if a.size == 1 else a[:3].copy()


class _Distribution:
    kind = LAMBERT
    emitting_light = 0
    sided = 0
    ior = 1.0
    roughness = 0.0

    def __init__(self, data):
        self.rho = _rgb(data.get("albedo", 1.0))

    def evaluate(self):
        return self.rho

    def record(self):
        """(albedo3, type, ior, roughness, two_sided) as the C ABI wants it."""
        return (tuple(float(x) for x in self.rho), self.kind, float(self.ior),
                float(self.roughness), 1 if self.sided == 0 else 0)


class BSDFLambertian(_Distribution):
    kind = LAMBERT


class BSDFLight(_Distribution):
    kind = EMITTER
    emitting_light = 1
    sided = 1


class BSDFMirror(_Distribution):
    kind = MIRROR


class BSDFDielectric(_Distribution):
    kind = DIELECTRIC
    sided = 1

    def __init__(self, data):
        super().__init__(data)
        self.ior = float(data.get("ior", 1.5))


class BSDFConductor(_Distribution):
    kind = CONDUCTOR

    def __init__(self, data):
        super().__init__(data)
        self.roughness = min(float(data.get("roughness", 0.0)), 1.0)


_TYPES = {"lambert": BSDFLambertian, "null": BSDFLight, "mirror": BSDFMirror,
          "dielectric": BSDFDielectric, "conductor": BSDFConductor}


class BSDF:
    def __init__(self, data):
        self._type = data["type"]
        if self._type not in _TYPES:
            print(f"[WARNING] bsdf of type {self._type} not implemented")
            raise NotImplementedError
        self.distribution = _TYPES[self._type](data)
        self.emitting_light = self.distribution.emitting_light
        self.sided = self.distribution.sided

    def get_distribution(self):
        return self.distribution

    def bsdf_info(self):
        return self.emitting_light, self.sided
