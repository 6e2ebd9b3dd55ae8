"""ctypes binding of libprt.so (include/prt.h).

This is the only place the Python side touches the native library.  There is
no CPU fallback: if the library is missing it is built with nvcc, and if that
fails -- or no B200 is present when a context is created -- an error is raised.
PyTorch appears only as the owner of device buffers (``tensor.data_ptr()``).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

MATERIAL_DTYPE = np.dtype(
    [("albedo", "<f4", 3), ("type", "<u4"), ("ior", "<f4"), ("roughness", "<f4"),
     ("two_sided", "<u4"), ("pad", "<u4"), ("emission", "<f4", 3), ("pad2", "<u4")])
assert MATERIAL_DTYPE.itemsize == 48
BSDF_QUERY_DTYPE = np.dtype([("d", "<f4", 3), ("type", "<u4"), ("ns", "<f4", 3), ("front", "<u4"), ("ior", "<f4"),
                             ("roughness", "<f4"), ("u", "<f4", 3), ("pad", "<u4", 3)])
assert BSDF_QUERY_DTYPE.itemsize == 64
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("tri", "<i4")])

TRACE_EXACT, TRACE_COUNT, TRACE_BRUTE, TRACE_BIN, TRACE_NO_BIN = 1, 2, 4, 8, 16
RENDER_EXACT_PRIMARY = 1
RENDER_PHYSICAL = 2
RENDER_COUNT = 4


class PrtCamera(C.Structure):
    _fields_ = [("iview", C.c_double * 16), ("sensor_w", C.c_double), ("sensor_h", C.c_double),
                ("focal", C.c_double), ("width", C.c_uint32), ("height", C.c_uint32),
                ("aperture", C.c_double)]


class PrtRenderParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("spp_begin", C.c_uint32), ("spp_end", C.c_uint32),
                ("max_depth", C.c_uint32), ("rr_start", C.c_uint32),
                ("light_color", C.c_float * 3), ("tmin", C.c_float), ("tmax", C.c_float),
                ("flags", C.c_uint32)]


class PrtBvhStats(C.Structure):
    _fields_ = [("n_tris", C.c_uint32), ("n_nodes", C.c_uint32), ("depth", C.c_uint32),
                ("max_leaf_tris", C.c_uint32), ("morton_sorted", C.c_uint32), ("sah_cost", C.c_float), ("ms_total", C.c_float),
                ("ms_morton", C.c_float), ("ms_sort", C.c_float), ("ms_hierarchy", C.c_float),
                ("ms_refit", C.c_float), ("ms_emit", C.c_float), ("ms_wall", C.c_float), ("morton_bits", C.c_uint32)]


class PrtBvhOptions(C.Structure):
    _fields_ = [("max_leaf_tris", C.c_uint32), ("cost_node", C.c_float), ("cost_tri", C.c_float),
                ("rotations", C.c_uint32), ("treelets", C.c_uint32), ("morton_bits", C.c_uint32)]


class PrtCounters(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64),
                ("flagged_rays", C.c_uint64), ("paths", C.c_uint64), ("warp_iters", C.c_uint64),
                ("node_lane_iters", C.c_uint64), ("leaf_phases", C.c_uint64),
                ("leaf_lane_phases", C.c_uint64), ("f64_decisions", C.c_uint64)]


class PrtKernelTimes(C.Structure):
    _fields_ = [("ms", C.c_float * 8), ("launches", C.c_uint32 * 8)]


PROF_CLASSES = ("raygen", "closest", "shade", "shadow", "exact_fixup", "other", "allreduce")
COMM_ID_BYTES = 128

EXPORTS = [
    "prt_abi_version", "prt_create", "prt_destroy", "prt_last_error", "prt_scene_set_triangles",
    "prt_scene_set_triangles_dev", "prt_bvh_build", "prt_camera_set", "prt_generate_rays",
    "prt_trace_closest", "prt_trace_any", "prt_trace_all", "prt_trace_closest_host", "prt_render", "prt_trace_paths",
    "prt_render_host", "prt_set_wave_paths", "prt_set_path_log", "prt_get_counters", "prt_reset_counters",
    "prt_synchronize", "prt_comm_unique_id", "prt_comm_init", "prt_comm_attach", "prt_comm_destroy", "prt_comm_info",
    "prt_allreduce_sum", "prt_render_sharded", "prt_profile_begin", "prt_profile_end", "prt_release_scratch",
    "prt_eval_specular",
]


class PrtError(RuntimeError):
    pass


def library_path():
    return _build.LIB


def load():
    """Load (building if needed) libprt.so.  Raises if it cannot be produced."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("PRT_LIB") or _build.build()  # PRT_LIB: a tuning variant built by profiles/
    lib = C.CDLL(path)
    vp, u32, u64, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float
    lib.prt_abi_version.restype = C.c_int
    lib.prt_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.prt_destroy.argtypes = [vp]
    lib.prt_destroy.restype = None
    lib.prt_last_error.argtypes = [vp]
    lib.prt_last_error.restype = C.c_char_p
    lib.prt_scene_set_triangles.argtypes = [vp, vp, vp, u32, vp, vp, u32, vp, u32]
    lib.prt_scene_set_triangles_dev.argtypes = [vp, vp, u32, vp]
    lib.prt_bvh_build.argtypes = [vp, C.POINTER(PrtBvhOptions), C.POINTER(PrtBvhStats)]
    lib.prt_camera_set.argtypes = [vp, C.POINTER(PrtCamera)]
    lib.prt_generate_rays.argtypes = [vp, u64, u32, u32, C.c_int, f32, f32, vp, vp]
    lib.prt_trace_closest.argtypes = [vp, vp, u64, vp, u32, vp]
    lib.prt_trace_any.argtypes = [vp, vp, u64, vp, u32, vp]
    lib.prt_trace_all.argtypes = [vp, vp, u64, vp, vp, u32, vp]
    lib.prt_trace_closest_host.argtypes = [vp, vp, u64, vp, u32]
    lib.prt_render.argtypes = [vp, C.POINTER(PrtRenderParams), vp, vp, vp]
    lib.prt_render_host.argtypes = [vp, C.POINTER(PrtRenderParams), vp]
    lib.prt_trace_paths.argtypes = [vp, vp, u64, C.POINTER(PrtRenderParams), vp, vp, vp]
    lib.prt_set_wave_paths.argtypes = [vp, u64]
    lib.prt_set_path_log.argtypes = [vp, vp, u64, vp]
    lib.prt_get_counters.argtypes = [vp, C.POINTER(PrtCounters)]
    lib.prt_reset_counters.argtypes = [vp]
    lib.prt_synchronize.argtypes = [vp]
    lib.prt_comm_unique_id.argtypes = [vp]
    lib.prt_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.prt_comm_attach.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.prt_comm_destroy.argtypes = [vp]
    lib.prt_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.prt_allreduce_sum.argtypes = [vp, vp, u64, vp]
    lib.prt_render_sharded.argtypes = [vp, C.POINTER(PrtRenderParams), vp, vp]
    lib.prt_release_scratch.argtypes = [vp]
    lib.prt_eval_specular.argtypes = [vp, vp, u64, vp, vp]
    lib.prt_profile_begin.argtypes = [vp]
    lib.prt_profile_end.argtypes = [vp, C.POINTER(PrtKernelTimes)]
    for name in EXPORTS:
        if name not in ("prt_destroy", "prt_last_error"):
            getattr(lib, name).restype = C.c_int
    if lib.prt_abi_version() != 3:
        raise PrtError("libprt.so ABI version mismatch")
    _LIB = lib
    return lib


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _dev_ptr(t):
    """torch CUDA tensor (or int address) -> void*"""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    if not t.is_cuda or not t.is_contiguous():
        raise PrtError("device buffers must be contiguous CUDA tensors")
    return C.c_void_p(t.data_ptr())


def _stream_ptr(stream, device=None):
    """cudaStream_t of ``stream`` (torch stream or raw handle); None = torch's current stream ON THE
    CONTEXT'S DEVICE (a stream of another device would be an invalid handle there)."""
    if stream is None:
        import torch
        stream = torch.cuda.current_stream(device)
    return C.c_void_p(getattr(stream, "cuda_stream", stream))


class Context:
    """One device context (prt_create .. prt_destroy)."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.prt_create(int(device), C.byref(h))
        if rc != 0:
            raise PrtError(self.lib.prt_last_error(None).decode())
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.prt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PrtError(f"[{rc}] " + self.lib.prt_last_error(self.h).decode())

    # -- scene ---------------------------------------------------------------
    def set_triangles(self, verts, normals=None, tri_material=None, materials=None, light_tris=None):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
        nt = verts.shape[0]
        normals = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(nt, 3)
        tri_material = None if tri_material is None else np.ascontiguousarray(tri_material, np.uint32)
        materials = None if materials is None else np.ascontiguousarray(materials, MATERIAL_DTYPE)
        light_tris = np.zeros(0, np.uint32) if light_tris is None else np.ascontiguousarray(light_tris, np.uint32)
        self._check(self.lib.prt_scene_set_triangles(
            self.h, _np_ptr(verts), _np_ptr(normals), nt, _np_ptr(tri_material), _np_ptr(materials),
            0 if materials is None else materials.shape[0], _np_ptr(light_tris), light_tris.shape[0]))

    def set_triangles_dev(self, verts_dev, nt, stream=None):
        self._check(self.lib.prt_scene_set_triangles_dev(self.h, _dev_ptr(verts_dev), int(nt),
                                                         _stream_ptr(stream, self.device)))

    def build_bvh(self, max_leaf_tris=4, cost_node=1.0, cost_tri=2.0, rotations=1, treelets=1, morton_bits=0):
        opts = PrtBvhOptions(int(max_leaf_tris), float(cost_node), float(cost_tri), int(rotations), int(treelets),
                             int(morton_bits))
        st = PrtBvhStats()
        self._check(self.lib.prt_bvh_build(self.h, C.byref(opts), C.byref(st)))
        return {k: getattr(st, k) for k, _ in PrtBvhStats._fields_}

    def set_camera(self, iview, sensor_w, sensor_h, focal, width, height, aperture=0.0):
        cam = PrtCamera()
        iv = np.ascontiguousarray(iview, np.float64).reshape(16)
        for i in range(16):
            cam.iview[i] = float(iv[i])
        cam.sensor_w, cam.sensor_h, cam.focal = float(sensor_w), float(sensor_h), float(focal)
        cam.width, cam.height = int(width), int(height)
        cam.aperture = float(aperture)
        self._check(self.lib.prt_camera_set(self.h, C.byref(cam)))
        self.resolution = (int(width), int(height))

    # -- tracing ---------------------------------------------------------------
    def generate_rays(self, rays_dev, seed=0, s0=0, s1=1, jitter=False, tmin=1e-5, tmax=99999.9,
                      stream=None):
        self._check(self.lib.prt_generate_rays(self.h, int(seed), int(s0), int(s1), 1 if jitter else 0,
                                               float(tmin), float(tmax), _dev_ptr(rays_dev),
                                               _stream_ptr(stream, self.device)))

    def trace_closest(self, rays_dev, n, hits_dev, flags=0, stream=None):
        self._check(self.lib.prt_trace_closest(self.h, _dev_ptr(rays_dev), int(n), _dev_ptr(hits_dev),
                                               int(flags), _stream_ptr(stream, self.device)))

    def trace_any(self, rays_dev, n, occluded_dev, flags=0, stream=None):
        self._check(self.lib.prt_trace_any(self.h, _dev_ptr(rays_dev), int(n), _dev_ptr(occluded_dev),
                                           int(flags), _stream_ptr(stream, self.device)))

    def trace_all(self, rays_dev, n, counts_dev, sums_dev, flags=0, stream=None):
        self._check(self.lib.prt_trace_all(self.h, _dev_ptr(rays_dev), int(n), _dev_ptr(counts_dev),
                                           _dev_ptr(sums_dev), int(flags), _stream_ptr(stream, self.device)))

    def trace_closest_host(self, rays, flags=0, out=None):
        """``out``: optional HIT_DTYPE array to receive the hits (page-locked for speed)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        if out is not None:
            hits = out
            assert hits.dtype == HIT_DTYPE and hits.shape[0] == rays.shape[0] and hits.flags.c_contiguous
        elif rays.shape[0] >= (1 << 16):  # big batches: page-locked result buffer (async D2H)
            import torch
            hits = torch.empty((rays.shape[0], 4), dtype=torch.float32, pin_memory=True).numpy()
            hits = hits.view(HIT_DTYPE).reshape(rays.shape[0])
        else:
            hits = np.empty(rays.shape[0], HIT_DTYPE)
        self._check(self.lib.prt_trace_closest_host(self.h, _np_ptr(rays), rays.shape[0], _np_ptr(hits),
                                                    int(flags)))
        return hits

    # -- render ---------------------------------------------------------------
    @staticmethod
    def render_params(seed=1, spp_begin=0, spp_end=1, max_depth=5, rr_start=0xFFFFFFFF,
                      light_color=(0.9, 0.85, 0.7), tmin=1e-5, tmax=99999.9, flags=0):
        p = PrtRenderParams()
        p.seed, p.spp_begin, p.spp_end = int(seed), int(spp_begin), int(spp_end)
        p.max_depth, p.rr_start = int(max_depth), int(rr_start)
        for k in range(3):
            p.light_color[k] = float(light_color[k])
        p.tmin, p.tmax, p.flags = float(tmin), float(tmax), int(flags)
        return p

    def render(self, params, accum_dev, prim_ids_dev=None, stream=None):
        self._check(self.lib.prt_render(self.h, C.byref(params), _dev_ptr(accum_dev),
                                        _dev_ptr(prim_ids_dev), _stream_ptr(stream, self.device)))

    def trace_paths(self, rays_dev, n, params, radiance_dev, prim_ids_dev=None, stream=None):
        self._check(self.lib.prt_trace_paths(self.h, _dev_ptr(rays_dev), int(n), C.byref(params),
                                             _dev_ptr(radiance_dev), _dev_ptr(prim_ids_dev), _stream_ptr(stream, self.device)))

    def set_path_log(self, segments_dev=None, count_dev=None):
        """segments_dev: torch f32 [capacity, 8] (prt_segment records), count_dev: torch i32/u32 [1];
        None switches the log off."""
        cap = 0 if segments_dev is None else int(segments_dev.shape[0])
        self._check(self.lib.prt_set_path_log(self.h, _dev_ptr(segments_dev), cap, _dev_ptr(count_dev)))

    def render_host(self, params, accum):
        assert accum.dtype == np.float32 and accum.flags.c_contiguous
        self._check(self.lib.prt_render_host(self.h, C.byref(params), _np_ptr(accum)))

    def eval_specular(self, queries_dev, n, out_dev, stream=None):
        """queries_dev: device buffer of n BSDF_QUERY_DTYPE records; out_dev: torch f32 [n, 4] (wi, valid)."""
        self._check(self.lib.prt_eval_specular(self.h, _dev_ptr(queries_dev), int(n), _dev_ptr(out_dev),
                                               _stream_ptr(stream, self.device)))

    def release_scratch(self):
        self._check(self.lib.prt_release_scratch(self.h))

    def set_wave_paths(self, paths):
        self._check(self.lib.prt_set_wave_paths(self.h, int(paths)))

    # -- multi-GPU ------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id (rank 0 makes it; hand it to every rank)."""
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = load().prt_comm_unique_id(buf)
        if rc != 0:
            raise PrtError(f"[{rc}] prt_comm_unique_id failed (is libnccl.so.2 loadable?)")
        return buf.raw

    def comm_init(self, uid, world, rank):
        assert len(uid) == COMM_ID_BYTES
        self._check(self.lib.prt_comm_init(self.h, C.c_char_p(uid), int(world), int(rank)))

    def comm_destroy(self):
        self._check(self.lib.prt_comm_destroy(self.h))

    def comm_info(self):
        w, r, v = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.prt_comm_info(self.h, C.byref(w), C.byref(r), C.byref(v)))
        return {"world": w.value, "rank": r.value, "nccl_version": v.value}

    def allreduce_sum(self, buf_dev, n_floats=None, stream=None):
        n = int(buf_dev.numel()) if n_floats is None else int(n_floats)
        self._check(self.lib.prt_allreduce_sum(self.h, _dev_ptr(buf_dev), n, _stream_ptr(stream, self.device)))

    def render_sharded(self, params, accum_dev, stream=None):
        """One frame of the sample-sharded render: this rank's shard, one all-reduce, sum added to accum."""
        self._check(self.lib.prt_render_sharded(self.h, C.byref(params), _dev_ptr(accum_dev),
                                                _stream_ptr(stream, self.device)))

    # -- measurement ----------------------------------------------------------
    def profile_begin(self):
        self._check(self.lib.prt_profile_begin(self.h))

    def profile_end(self):
        """{class: (ms, launches)} of every kernel launched since profile_begin (synchronises)."""
        t = PrtKernelTimes()
        self._check(self.lib.prt_profile_end(self.h, C.byref(t)))
        return {name: (float(t.ms[i]), int(t.launches[i])) for i, name in enumerate(PROF_CLASSES)}

    def counters(self):
        c = PrtCounters()
        self._check(self.lib.prt_get_counters(self.h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in PrtCounters._fields_}

    def reset_counters(self):
        self._check(self.lib.prt_reset_counters(self.h))

    def synchronize(self):
        self._check(self.lib.prt_synchronize(self.h))
