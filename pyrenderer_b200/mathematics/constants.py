"""Numeric constants of the path (reference: mathematics/constants.py:3-16)."""
import numpy as np

Pi = 3.14159265358979323846
InvPi = 0.31830988618379067154
_F32 = np.finfo(np.float32)
MAX_F = float(_F32.max)
EPS = float(_F32.tiny)
MACHINE_EPS = float(_F32.eps) * 0.5
GAMMA2_3 = (3 * MACHINE_EPS) / (1 - 3 * MACHINE_EPS)

# values the integrator uses (reference: core/tracing.py:120,127)
T_MIN = 1e-5
T_MAX = 99999.9
LIGHT_COLOR = (0.9, 0.85, 0.7)
