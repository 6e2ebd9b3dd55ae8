"""Fixture for the physically-based mode: 8x8-pixel block means of the
reference repository's only externally produced image, media/cornell-box/TungstenRender.exr
(1024x1024 linear radiance rendered by Tungsten from the same scene.json).

    python tests/golden/make_tungsten_fixture.py      # needs /root/reference (this container only)

Writes tests/golden/tungsten_cornell_128.npz: `mean` f32[128,128,3], row 0 = TOP image row."""
import os

import numpy as np

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402

REF = "/root/reference/media/cornell-box/TungstenRender.exr"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tungsten_cornell_128.npz")
exr = cv2.imread(REF, cv2.IMREAD_UNCHANGED)[:, :, ::-1].astype(np.float64)  # BGR -> RGB
assert exr.shape == (1024, 1024, 3)
mean = exr.reshape(128, 8, 128, 8, 3).mean(axis=(1, 3)).astype(np.float32)
np.savez_compressed(OUT, mean=mean)
print("wrote", OUT, mean.shape, float(mean.mean()))
