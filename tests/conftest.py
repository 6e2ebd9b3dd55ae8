import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
SCENE_JSON = os.path.join(ROOT, "pyrenderer_b200", "media", "cornell_box.json")
CUBE_OBJ = os.path.join(ROOT, "pyrenderer_b200", "media", "unit_cube.obj")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def cornell():
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    return read_file(SCENE_JSON)


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyrenderer_b200 import _abi
    return _abi.Context(0)


def make_rays(o, d, tmin=1e-5, tmax=3.4028234663852886e+38):
    o = np.asarray(o, np.float32)
    d = np.asarray(d, np.float32)
    r = np.empty((o.shape[0], 8), np.float32)
    r[:, 0:3] = o
    r[:, 3] = tmin
    r[:, 4:7] = d
    r[:, 7] = tmax
    return r


def random_soup(n, seed=7):
    """centres U[0,1]^3, two edge vectors U[-h,h]^3, h = 0.75 n^(-1/3)."""
    rng = np.random.default_rng(seed)
    h = 0.75 * n ** (-1.0 / 3.0)
    c = rng.uniform(0, 1, (n, 1, 3))
    e = rng.uniform(-h, h, (n, 2, 3))
    return np.concatenate([c, c + e[:, :1], c + e[:, 1:]], 1).astype(np.float32)


def random_rays(n, seed=11, tmin=1e-5, tmax=3.4e38):
    rng = np.random.default_rng(seed)
    o = rng.uniform(0, 1, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return make_rays(o, d, tmin, tmax)
