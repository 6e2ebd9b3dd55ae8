"""Host-side logic and the C-ABI surface; no GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from conftest import CUBE_OBJ, ROOT, SCENE_JSON


def test_transforms_match_reference(golden):
    import json
    from pyrenderer_b200.mathematics.affine_transformation import make_transformation_matrix
    data = json.load(open(SCENE_JSON))
    for k, prim in enumerate(data["primitives"]):
        m = make_transformation_matrix(prim["transform"])
        assert np.allclose(m, golden["transforms"][k], rtol=0, atol=1e-15)
        assert np.linalg.det(m[:3, :3]) > 0


def test_loader_counts_and_geometry(golden, cornell):
    scene, cam = cornell
    assert scene.vertices.shape == (72, 3) and scene.faces.shape == (36, 3)
    assert np.allclose(scene.vertices, golden["scene_vertices"], rtol=0, atol=1e-15)
    assert np.array_equal(scene.faces, golden["scene_faces"])
    a = scene.arrays()
    assert np.allclose(a["normals"], golden["scene_normals"], atol=1e-7)
    assert list(a["light_tris"]) == [34, 35]
    alb = a["materials"][a["tri_material"]]["albedo"]
    assert np.allclose(alb, golden["scene_albedo"], atol=1e-7)
    for k, prim in enumerate(scene.primitives):
        assert np.allclose(prim.bounds.min_coord, golden["prim_bounds"][k, 0], atol=1e-15)
        assert np.allclose(prim.bounds.max_coord, golden["prim_bounds"][k, 1], atol=1e-15)
    # zero-extent ("empty") boxes exactly where the reference's BBox.update_empty says so
    want = [bool(np.any(np.abs(b[0] - b[1]) <= 1.1754943508222875e-38)) for b in golden["prim_bounds"]]
    assert [p.bounds.is_empty() for p in scene.primitives] == want
    assert want[0] and not want[5] and not want[6]
    assert cam.get_resolution() == [1024, 1024]


def test_loader_errors(tmp_path, capsys):
    import json
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    base = json.load(open(SCENE_JSON))
    bad = dict(base)
    bad["primitives"] = base["primitives"] + [{"type": "sphere", "bsdf": "Floor", "transform": {}}]
    p = tmp_path / "s.json"
    p.write_text(json.dumps(bad))
    scene, _ = read_file(str(p))
    assert "[WARNING] sphere not implemented" in capsys.readouterr().out
    assert len(scene.primitives) == 8
    bad2 = dict(base)
    bad2["bsdfs"] = base["bsdfs"] + [{"name": "x", "type": "plastic", "albedo": 1}]
    p.write_text(json.dumps(bad2))
    with pytest.raises(NotImplementedError):
        read_file(str(p))


def test_camera_matches_reference(golden):
    from pyrenderer_b200.core.camera import Camera
    cams = [dict(position=[0, 1, 6.8], looking_at=[0, 1, 0], up=[0, 1, 0], resolution=[1024, 1024], fov=19.5),
            dict(position=[2.6, 2.1, 3.4], looking_at=[0.5, 0.5, 0.5], up=[0, 1, 0], resolution=[640, 480], fov=35.0)]
    for ci, kw in enumerate(cams):
        cam = Camera(**kw)
        assert np.allclose(cam.iview, golden[f"cam{ci}_iview"], rtol=0, atol=1e-15)
        for uv, want in zip(golden["cam_uv"][:64], golden[f"cam{ci}_rays"][:64]):
            r = cam.generate_ray(uv)
            assert np.allclose(r.position, want[:3], atol=1e-15)
            assert np.allclose(r.direction, want[3:], atol=1e-15)


def test_exr_light_mask(golden, cornell):
    """the pixels of TungstenRender.exr that equal the emission (17,12,4)
    must lie inside the light-quad mask of the restated camera (pins look_at / fov /
    aspect / row flip)."""
    scene, cam = cornell
    a = scene.arrays()
    W = H = 1024
    sw, sh = cam.sensor()
    ocam = oracle.make_camera(cam.iview, sw, sh, 1.0, W, H)
    rows, cols = golden["exr_light_rows"].astype(int), golden["exr_light_cols"].astype(int)
    assert rows.size == 4464
    r0, r1, c0, c1 = rows.min() - 3, rows.max() + 4, cols.min() - 3, cols.max() + 4
    jj, ii = np.meshgrid(np.arange(r0, r1), np.arange(c0, c1), indexing="ij")
    rays = np.empty((jj.size, 8), np.float32)
    for k, (row, col) in enumerate(zip(jj.ravel(), ii.ravel())):
        # image row 0 is the top: v = (H-1-row + 0.5)/H   (main.py:55 row flip)
        o, d = oracle.generate_ray(ocam, (col + 0.5) / W, (H - 1 - row + 0.5) / H)
        rays[k] = [*o, 1e-5, *d, 99999.9]
    ids, _, _, _ = oracle.closest_hit(a["tris"], rays)
    mask = np.isin(ids, [34, 35]).reshape(jj.shape)
    exr = np.zeros_like(mask)
    exr[rows - r0, cols - c0] = True
    assert not np.any(exr & ~mask), "EXR light pixels outside the restated light mask"
    iou = (exr & mask).sum() / (exr | mask).sum()
    assert iou > 0.88
    ys, xs = np.nonzero(mask)
    assert abs((ys.min() + r0) - rows.min()) <= 2 and abs((xs.max() + c0) - cols.max()) <= 2


def test_obj_loader():
    from pyrenderer_b200.io_utils.read_tungsten import read_obj
    v, f = read_obj(CUBE_OBJ)
    assert v.shape == (8, 3) and f.shape == (12, 3)
    assert v.min() == 0.0 and v.max() == 1.0
    assert list(f[0]) == [0, 6, 4]


def test_sample_sharding():
    from pyrenderer_b200.core.tracing import shard_samples
    for spp in (1, 7, 16, 1024):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_samples(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def test_abi_library_exports_every_declared_symbol():
    """libprt.so loads without a GPU and exports exactly what include/prt.h declares."""
    from pyrenderer_b200 import _abi, build
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "prt.h")).read()
    declared = sorted(set(re.findall(r"\b(prt_[a-z_]+)\s*\(", header)))
    assert declared == sorted(_abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.prt_abi_version() == 3


def test_abi_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyrenderer_b200 import _abi
    with pytest.raises(_abi.PrtError) as e:
        _abi.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_struct_layouts_match_oracle():
    from pyrenderer_b200 import _abi
    assert _abi.MATERIAL_DTYPE == oracle.MATERIAL_DTYPE
    assert ctypes.sizeof(_abi.PrtCamera) == ctypes.sizeof(oracle.Camera) == 168
    assert ctypes.sizeof(_abi.PrtRenderParams) == ctypes.sizeof(oracle.RenderParams) == 48
    assert _abi.HIT_DTYPE.itemsize == 16


def test_oracle_light_hit_known_answer(cornell):
    """a primary ray that hits the light returns exactly light_color."""
    scene, cam = cornell
    a = scene.arrays()
    sw, sh = cam.sensor()
    W = H = 64
    ocam = oracle.make_camera(cam.iview, sw, sh, 1.0, W, H)
    P = oracle.make_params(seed=3, spp_begin=0, spp_end=4, max_depth=5)
    acc, ids, stats = oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"],
                                    a["light_tris"], ocam, P, want_ids=True)
    on_light = np.all(np.isin(ids, [34, 35]), axis=2)
    assert on_light.sum() >= 10
    mean = acc[..., :3] / acc[..., 3:]
    assert np.allclose(mean[on_light], np.array([0.9, 0.85, 0.7], np.float32).astype(np.float64), atol=1e-12)
    assert stats[0] > W * H * 4 and stats[1] > 0
    assert np.all(mean >= 0) and np.isfinite(mean).all()


def test_subdivision_utility(cornell):
    from pyrenderer_b200.mathematics.subdivide import subdivide, subdivide_scene_arrays
    scene, _ = cornell
    a = scene.arrays()
    t, parent = subdivide(a["tris"], 3)
    assert t.shape == (36 * 64, 3, 3) and np.array_equal(np.bincount(parent), np.full(36, 64))
    area = lambda x: 0.5 * np.linalg.norm(np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]), axis=1)
    assert np.allclose(np.bincount(parent, weights=area(t)), area(a["tris"]), rtol=1e-5)
    # children keep the parent's orientation (normals) and shared edges get identical midpoints
    n_child = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    n_par = np.cross(a["tris"][:, 1] - a["tris"][:, 0], a["tris"][:, 2] - a["tris"][:, 0])[parent]
    assert np.all(np.einsum("ij,ij->i", n_child, n_par) > 0)
    verts = np.unique(t.reshape(-1, 3), axis=0)
    assert verts.shape[0] < t.shape[0] * 3 / 3.5  # welded: ~6 triangles share a vertex
    b = subdivide_scene_arrays(a, np.where(np.isin(a["tri_prim"], [5, 6]), 1, 2))
    assert b["tris"].shape[0] == 12 * 16 + 24 * 4
    assert np.array_equal(np.unique(b["tri_material"][b["light_tris"]]), [7])
    assert b["light_tris"].shape[0] == 2 * 16


def test_output_stage(tmp_path):
    from pyrenderer_b200.core import tracing
    from pyrenderer_b200.main import write_png
    acc = np.zeros((4, 6, 4), np.float32)
    acc[..., 3] = 2.0
    acc[0, :, :3] = 2.0    # bottom row (v = 0) -> mean 1.0
    acc[3, 0, :3] = 8.0    # top-left -> mean 4.0 (would wrap in the reference's uint8 cast)
    img = tracing.to_image(acc)
    assert img.shape == (4, 6, 3) and img[3, 2, 0] == 1.0 and img[0, 0, 0] == 4.0  # row flip, main.py:55
    assert np.allclose(tracing.to_image(acc, "sqrt")[0, 0], 2.0)
    r = tracing.to_image(acc, "reinhard")
    assert np.allclose(r[0, 0], 1.0, atol=1e-6) and 0 < r[3, 2, 0] < 1.0   # white point maps to 1
    u8 = tracing.to_uint8(img)
    assert u8[0, 0, 0] == 255 and u8[3, 2, 0] == 255 and u8[1, 1, 0] == 0   # clamp, not wrap-around
    p = tmp_path / "o.png"
    write_png(str(p), u8)
    data = p.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and b"IHDR" in data and b"IEND" in data
    try:
        import cv2
        back = cv2.imread(str(p))[..., ::-1]
        assert np.array_equal(back, u8)
    except ImportError:
        pass


def test_tungsten_mesh_primitive(tmp_path):
    """a Tungsten scene that references an OBJ mesh loads into the same
    Scene protocol (global triangle ids = file face order, transform applied, material bound)."""
    import json
    import shutil
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    shutil.copy(CUBE_OBJ, tmp_path / "cube.obj")
    scene_json = {
        "bsdfs": [{"name": "m", "albedo": [0.2, 0.4, 0.6], "type": "lambert"},
                  {"name": "glass", "type": "dielectric", "ior": 1.33},
                  {"name": "L", "albedo": 1, "type": "null"}],
        "primitives": [
            {"type": "mesh", "file": "cube.obj", "bsdf": "m", "transform": {"position": [1, 2, 3], "scale": [2, 2, 2]}},
            {"type": "cube", "bsdf": "glass", "transform": {}},
            {"type": "quad", "bsdf": "L", "transform": {"position": [0, 5, 0]}}],
        "camera": {"resolution": [32, 16], "fov": 40,
                   "transform": {"position": [0, 0, 5], "look_at": [0, 0, 0], "up": [0, 1, 0]}}}
    p = tmp_path / "scene.json"
    p.write_text(json.dumps(scene_json))
    scene, cam = read_file(str(p))
    a = scene.arrays()
    assert a["tris"].shape == (12 + 12 + 2, 3, 3)
    assert np.allclose(a["tris"][:12].reshape(-1, 3).min(0), [1, 2, 3]) and np.allclose(a["tris"][:12].reshape(-1, 3).max(0), [3, 4, 5])
    assert list(a["light_tris"]) == [24, 25]
    m = a["materials"]
    assert m[1]["type"] == 3 and abs(m[1]["ior"] - 1.33) < 1e-6 and m[1]["two_sided"] == 0
    assert m[2]["type"] == 1 and m[0]["two_sided"] == 1
    # outward normals of the OBJ cube (+normalize(e1 x e2)): they point away from the centre
    c = a["tris"][:12].mean(axis=1) - np.array([2, 3, 4])
    assert np.all(np.einsum("ij,ij->i", a["normals"][:12], c) > 0)
    assert cam.aspect_ratio == 2.0


def test_ray_logger_container(tmp_path):
    """debug/ray_logger.py:1-16: points / lines / colors grow as in the reference; device records
    are appended in (path, bounce, light connection last) order; OBJ export has one `l` per line."""
    from pyrenderer_b200.core.ray import Ray
    from pyrenderer_b200.debug.ray_logger import BOUNCE_COLORS, LIGHT_COLOR, RayLogger
    lg = RayLogger()
    lg.add(Ray(np.array([0.0, 1.0, 2.0]), np.array([0.0, 0.0, -1.0])), t=5, color=[0, 1, 0])
    lg.add_line([0, 0, 0], [1, 1, 1])
    assert len(lg.points) == 4 and lg.lines == [[0, 1], [2, 3]] and lg.colors == [[0, 1, 0], [1, 0, 0]]
    assert np.allclose(lg.points[1], [0, 1, -3])
    rec = np.zeros((4, 8), np.float32)
    kinds = np.array([1, -1, 0, 0], np.int32)
    paths = np.array([7, 7, 7, 3], np.uint32)
    rec[:, 3] = kinds.view(np.float32)
    rec[:, 7] = paths.view(np.float32)
    rec[:, 0] = np.arange(4)
    lg.add_device_segments(rec)
    assert lg.paths[2:] == [3, 7, 7, 7] and lg.kinds[2:] == [0, 0, 1, -1]
    assert lg.colors[-1] == LIGHT_COLOR and lg.colors[-2] == BOUNCE_COLORS[1]
    out = tmp_path / "rays.obj"
    lg.write_obj(str(out))
    text = out.read_text().splitlines()
    assert sum(l.startswith("v ") for l in text) == 12 and sum(l.startswith("l ") for l in text) == 6


def test_hdr_output_roundtrip(tmp_path):
    """SURVEY 8f rank 2, HDR half: linear radiance survives a .pfm round trip bit for bit (values
    above 1 included), top row first on both sides."""
    from pyrenderer_b200.main import read_pfm, write_hdr
    rng = np.random.default_rng(2)
    img = (rng.uniform(0, 20, (7, 5, 3))).astype(np.float32)
    p = str(tmp_path / "x.pfm")
    write_hdr(p, img)
    assert np.array_equal(read_pfm(p), img)
    head = open(p, "rb").read(12)
    assert head.startswith(b"PF\n5 7\n")


def test_thin_lens_camera_host_and_oracle(cornell):
    """core/camera.py:63-65: with aperture > 0 the ray starts on a square lens (camera-space x, y in
    [-a/2, a/2)) and still aims at the pinhole ray's point on the plane z = -focal_dist: all lens rays
    of one screen coordinate meet there.  aperture == 0 (the loader's value) stays bit-identical."""
    import oracle
    from pyrenderer_b200.core.camera import Camera
    pin = Camera([0, 1, 6.8], [0, 1, 0], [0, 1, 0], [64, 64], fov=19.5, focal_dist=6.0)
    lens = Camera([0, 1, 6.8], [0, 1, 0], [0, 1, 0], [64, 64], fov=19.5, aperture=0.4, focal_dist=6.0)
    uv = np.array([0.3, 0.8])
    r0 = pin.generate_ray(uv)
    focus = r0.position + r0.direction * (6.0 / -r0.direction[2])  # plane z_cam = -6  <=>  world z = 0.8
    origins = []
    for _ in range(50):
        r = lens.generate_ray(uv)
        t = (focus[2] - r.position[2]) / r.direction[2]
        assert np.allclose(r.position + t * r.direction, focus, atol=1e-6)
        assert abs(np.linalg.norm(r.direction) - 1) < 1e-12 and abs(r.position[2] - 6.8) < 1e-12
        origins.append(r.position[:2] - np.array([0.0, 1.0]))
    origins = np.array(origins)
    assert np.all(np.abs(origins) <= 0.2 + 1e-7) and origins.std(axis=0).min() > 0.05
    # oracle: same construction with explicit lens samples; lens centre == pinhole, bit for bit
    iview, sw, sh, focal, W, H = lens.device_record()
    oc_pin = oracle.make_camera(iview, sw, sh, focal, W, H)
    oc_lens = oracle.make_camera(iview, sw, sh, focal, W, H, aperture=0.4)
    o0, d0 = oracle.generate_ray(oc_pin, 0.3, 0.8)
    assert np.array_equal(o0, r0.position) and np.allclose(d0, r0.direction, atol=1e-15)
    rays = oracle.generate_rays(oc_lens, seed=4, s0=0, s1=8, jitter=True)
    o = rays[..., 0:3].reshape(-1, 3).astype(np.float64) - np.array([0.0, 1.0, 6.8])
    assert np.abs(o[:, 2]).max() < 1e-6 and np.abs(o[:, :2]).max() <= 0.2 + 1e-6 and o[:, :2].std(axis=0).min() > 0.08
    rays0 = oracle.generate_rays(oc_pin, seed=4, s0=0, s1=8, jitter=True)
    assert np.all(rays0[..., 0:3] == np.array([0.0, 1.0, 6.8], np.float32))


def test_ncu_profiles_carry_a_source_fingerprint_and_bench_refuses_stale_ones(tmp_path, monkeypatch, capsys):
    """profiles/ncu_*.json hold hardware counters captured under ncu; each is stamped with the fingerprint
    of the CUDA sources it belongs to.  bench.py quotes a counter only while the fingerprint matches and
    says so loudly (stderr + "stale") when it does not."""
    import glob
    import json
    import bench
    from pyrenderer_b200.kernel_fingerprint import fingerprint
    fp = fingerprint()
    assert len(fp) == 16 and fp == fingerprint()
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_*.json")))
    assert files, "no committed ncu profile"
    for f in files:
        rec = json.load(open(f))
        for key in ("kernel", "duration_ns", "dram_bytes", "source_fingerprint", "l1_data_pipe_pct", "lanes_per_instruction"):
            assert key in rec, (f, key)
        assert rec["duration_ns"] > 0 and rec["dram_bytes"] > 0
    name = os.path.basename(files[0])[4:-5]
    rec = bench.ncu_profile(name)
    assert rec is not None and (rec.get("stale") or rec["source_fingerprint"] == fp)
    # a profile captured from other sources is refused, loudly
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "profiles")
    json.dump({"source_fingerprint": "0" * 16, "dram_bytes": 1.0}, open(tmp_path / "profiles" / "ncu_x.json", "w"))
    out = bench.ncu_profile("x")
    assert out["stale"] is True and "dram_bytes" not in out
    assert "STALE" in capsys.readouterr().err
    assert bench.ncu_profile("missing") is None
