"""Render driver (reference: main.py:40-126 and main_taichi.py:102-127).

    python -m pyrenderer_b200.main [--scene media/cornell_box.json] [--samples 8]
        [--max-depth 5] [--width W --height H] [--seed 1] [--out test.png]

Loads a Tungsten scene, renders it on the GPU and writes the image the way
main.py does (row flip, x255 -> uint8; clamped instead of wrapping).
"""
import argparse
import os
import time

import numpy as np

from .core import tracing
from .io_utils.read_tungsten import read_file

DEFAULT_SCENE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "media",
                             "cornell_box.json")


def write_png(path, rgb8):
    """Tiny PNG writer (the reference uses skimage.io.imsave, absent here)."""
    import struct
    import zlib
    h, w, _ = rgb8.shape
    raw = b"".join(b"\x00" + rgb8[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def write_hdr(path, img):
    """Linear radiance, top row first, to ``.pfm`` (Portable Float Map, written here) or ``.exr``
    (through OpenCV when it is built with OpenEXR) -- the HDR half of SURVEY 8f rank 2; the
    reference only ever writes the 8-bit image (main.py:57-59)."""
    img = np.ascontiguousarray(img, np.float32)
    if path.lower().endswith(".pfm"):
        h, w, _ = img.shape
        with open(path, "wb") as f:
            f.write(f"PF\n{w} {h}\n-1.0\n".encode())  # negative scale = little endian
            f.write(img[::-1].tobytes())                 # PFM stores the bottom row first
        return
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    if not cv2.imwrite(path, img[:, :, ::-1]):
        raise OSError(f"could not write {path}")


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = (int(x) for x in f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(), "<f4" if scale < 0 else ">f4").reshape(h, w, 3)
    return np.ascontiguousarray(data[::-1])


def main(scene_file=DEFAULT_SCENE, samples=8, max_depth=5, width=None, height=None, seed=1,
         out="test.png", tonemap=None, device=0, physical=False):
    a_scene, a_camera = read_file(scene_file)
    if width and height:
        a_camera.resolution = [width, height]
    import torch
    t0 = time.time()
    accum = tracing.render(a_scene, a_camera, spp=samples, max_depth=max_depth, seed=seed,
                           device=device, physical=physical)
    torch.cuda.synchronize()
    dt = time.time() - t0
    image = tracing.to_uint8(tracing.to_image(accum, tonemap))
    if out and out.lower().endswith((".pfm", ".exr")):
        write_hdr(out, tracing.to_image(accum, None))
    elif out:
        write_png(out, image)
    c = a_scene.commit(device).counters()
    rays = c["rays_closest"] + c["rays_shadow"]
    print(f"{samples} spp in {dt:.3f} s  ({samples / dt:.2f} samples/s, {rays / dt / 1e6:.1f} Mrays/s)")
    return image


def main_progressive(scene_file=DEFAULT_SCENE, iterations=1001, max_depth=16, seed=1, out="out.png",
                     interval=10, save_every=100, device=0, width=None, height=None):
    """Progressive driver in the shape of the reference's Taichi loop (main_taichi.py:102-127):
    one sample per pixel per iteration into the same accumulation buffer, a "samples/s" line
    every `interval` iterations, sqrt-tonemapped image every `save_every`, stop after `iterations`.
    Exactly resumable: iteration k is Philox sample index k."""
    import torch
    a_scene, a_camera = read_file(scene_file)
    if width and height:
        a_camera.resolution = [width, height]
    accum = tracing.new_accum(a_camera, device)
    last_t = time.time()
    for iteration in range(iterations):
        tracing.render(a_scene, a_camera, spp=1, max_depth=max_depth, seed=seed, spp_begin=iteration,
                       accum=accum, device=device)
        if iteration % interval == 0:
            torch.cuda.synchronize()
            print("{:.2f} samples/s ({} iterations)".format(interval / max(time.time() - last_t, 1e-9), iteration))
            last_t = time.time()
        if iteration % save_every == 0 and iteration > 0 and out:
            write_png(out, tracing.to_uint8(tracing.to_image(accum, "sqrt")))
    torch.cuda.synchronize()
    if out:
        write_png(out, tracing.to_uint8(tracing.to_image(accum, "sqrt")))
    return accum


def main_debug(scene_file=DEFAULT_SCENE, step=10, samples=6, max_depth=5, seed=1, out="rays.obj", device=0):
    """The reference's main_debug (main.py:66-85): 6 jittered camera rays through every 10th
    pixel, traced with a RayLogger; the logged segments go to a Wavefront OBJ of line elements
    instead of an Open3D window."""
    import random
    from .debug.ray_logger import RayLogger
    a_scene, a_camera = read_file(scene_file)
    x_dim, y_dim = a_camera.get_resolution()
    rng = random.Random(seed)
    rays = []
    for i in range(0, x_dim, step):
        for j in range(0, y_dim, step):
            for _ in range(samples):
                x = (i + rng.random()) / float(x_dim)
                y = (j + rng.random()) / float(y_dim)
                rays.append(a_camera.generate_ray(np.array([x, y])))
    ray_logger = RayLogger()
    tracing.path_tracing(rays, a_scene, ray_logger, spp=1, max_depth=max_depth, seed=seed, device=device)
    if out:
        ray_logger.write_obj(out)
    print(f"{len(rays)} paths, {len(ray_logger.lines)} segments"
          f" ({sum(1 for k in ray_logger.kinds if k < 0)} light connections)")
    return ray_logger


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default=DEFAULT_SCENE)
    ap.add_argument("--samples", type=int, default=8, help="number of spp")
    ap.add_argument("--max-depth", type=int, default=5)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default="test.png")
    ap.add_argument("--tonemap", default=None, choices=[None, "sqrt", "reinhard"])
    ap.add_argument("--progressive", type=int, default=0, metavar="N",
                    help="main_taichi.py-style loop: N iterations of 1 spp into one buffer")
    ap.add_argument("--physical", action="store_true",
                    help="physically-based estimator (scene emission + MIS) instead of the reference's")
    ap.add_argument("--debug-rays", metavar="OBJ", default=None,
                    help="main_debug: log the segments of a sparse set of camera paths to a Wavefront OBJ")
    a = ap.parse_args()
    if a.debug_rays:
        main_debug(a.scene, max_depth=a.max_depth, seed=a.seed, out=a.debug_rays)
        raise SystemExit(0)
    if a.progressive:
        main_progressive(a.scene, a.progressive, a.max_depth, a.seed, a.out)
        raise SystemExit(0)
    main(a.scene, a.samples, a.max_depth, a.width, a.height, a.seed, a.out, a.tonemap, physical=a.physical)
