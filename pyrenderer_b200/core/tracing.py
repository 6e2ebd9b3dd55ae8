"""Render entry points (reference: core/tracing.py:47-155 PathTracer.trace, and
the pixel x sample loops of main.py:28-59 / main_taichi.py:80-99).

``render`` replaces the whole loop with ONE C-ABI call per frame: the wavefront
integrator (csrc/wavefront.cu) adds samples [spp_begin, spp_end) of every
pixel to an fp32 accumulation buffer owned by a torch tensor.
``render_distributed`` shards the sample range over the ranks of the default
torch.distributed group (replicated scene + BVH) and sums the buffers with one
all-reduce (NCCL over NVLink on GPUs).
"""
import numpy as np

from ..mathematics.constants import LIGHT_COLOR, T_MAX, T_MIN

RR_OFF = 0xFFFFFFFF


def _torch():
    import torch
    return torch


def new_accum(camera, device=0):
    torch = _torch()
    w, h = camera.get_resolution()
    return torch.zeros((h, w, 4), dtype=torch.float32, device=f"cuda:{device}")


def render(scene, camera, spp=8, max_depth=5, seed=1, spp_begin=0, rr_start=RR_OFF, device=0,
           accum=None, want_prim_ids=False, light_color=LIGHT_COLOR, stream=None,
           exact_primary=True, physical=False):
    """Add samples [spp_begin, spp_begin+spp) to ``accum`` (created if None).

    ``exact_primary`` (default): bounce 0 runs in PRT_TRACE_EXACT mode, so primary-hit triangle
    ids are bit-exact against the reference's intersection code (mathematics/intersection.py:42-65,
    106-116); it costs ~2 % of a depth-8 Cornell render (bench.py ``render.exact_primary_cost``).

    ``physical`` selects the physically-based estimator (PRT_RENDER_PHYSICAL: scene emission,
    MIS of light and BSDF sampling -- comparable with Tungsten's render of the same scene)
    instead of the reference's own estimator (core/tracing.py:116-155).

    Returns ``accum`` (torch f32 [h, w, 4]: rgb sums + sample count, row 0 = bottom
    image row), or ``(accum, prim_ids)`` with ``want_prim_ids``.
    """
    torch = _torch()
    ctx = scene.commit(device)
    ctx.set_camera(*camera.device_record(), aperture=float(getattr(camera, "aperture", 0.0)))
    if accum is None:
        accum = new_accum(camera, device)
    w, h = camera.get_resolution()
    ids = None
    if want_prim_ids:
        ids = torch.full((h, w, spp), -2, dtype=torch.int32, device=accum.device)
    params = ctx.render_params(seed=seed, spp_begin=spp_begin, spp_end=spp_begin + spp,
                               max_depth=max_depth, rr_start=rr_start, light_color=light_color,
                               tmin=T_MIN, tmax=T_MAX,
                               flags=(1 if exact_primary else 0) | (2 if physical else 0))
    ctx.render(params, accum, ids, stream)
    return (accum, ids) if want_prim_ids else accum


def path_tracing(ray, a_scene, ray_logger=None, spp=1, max_depth=5, seed=1, sample_index=0, device=0):
    """Drop-in for the reference's ``e, r = path_tracing(ray, a_scene[, ray_logger])`` (call
    sites main.py:22,34,78; the function itself is missing at the reference's HEAD).
    ``ray`` is one host-side Ray or a sequence of them.  Returns ``(e, r)`` with ``e + r`` the
    mean radiance of ``spp`` paths started on the ray: ``e`` = what a directly visible emitter
    contributes, ``r`` = everything gathered after the first bounce.  With a ``ray_logger``
    (debug/ray_logger.py RayLogger) every traced segment -- path segments per bounce and
    unoccluded light connections -- is appended to it (main.py:66-85)."""
    torch = _torch()
    rays = ray if isinstance(ray, (list, tuple)) else [ray]
    rec = np.array([[*r.position, T_MIN, *r.direction, T_MAX] for r in rays], np.float32)
    ctx = a_scene.commit(device)
    d = torch.from_numpy(rec).to(f"cuda:{device}")
    rad = torch.zeros((len(rays), 4), dtype=torch.float32, device=d.device)
    ids = torch.empty((len(rays), spp), dtype=torch.int32, device=d.device)
    params = ctx.render_params(seed=seed, spp_begin=sample_index, spp_end=sample_index + spp,
                               max_depth=max_depth, tmin=T_MIN, tmax=T_MAX)
    if ray_logger is not None:
        cap = len(rays) * spp * max_depth * 2
        seg = torch.zeros((cap, 8), dtype=torch.float32, device=d.device)
        cnt = torch.zeros((1,), dtype=torch.int32, device=d.device)
        ctx.set_path_log(seg, cnt)
    try:
        ctx.trace_paths(d, len(rays), params, rad, ids)
        torch.cuda.synchronize(d.device)
    finally:
        if ray_logger is not None:
            ctx.set_path_log(None, None)
    if ray_logger is not None:
        ray_logger.add_device_segments(seg[: min(int(cnt.item()), cap)].cpu().numpy())
    out = rad.cpu().numpy().astype(np.float64)
    mean = out[:, :3] / np.maximum(out[:, 3:4], 1.0)
    lights = set(int(t) for t in a_scene.arrays()["light_tris"])
    direct = np.array([[int(t) in lights for t in row] for row in ids.cpu().numpy()]).all(axis=1)
    e = np.where(direct[:, None], mean, 0.0)
    r = np.where(direct[:, None], 0.0, mean)
    return (e[0], r[0]) if not isinstance(ray, (list, tuple)) else (e, r)


def shard_samples(spp, rank, world):
    """Contiguous sample range of ``rank``: [r*S/G, (r+1)*S/G)."""
    return (spp * rank) // world, (spp * (rank + 1)) // world


def default_device():
    """Device of this rank: LOCAL_RANK under torchrun, else torch's current CUDA device."""
    import os
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    torch = _torch()
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def init_distributed(scene, device=None, group=None):
    """Give the scene's device context an NCCL communicator over the ranks of the torch.distributed
    ``group`` (C ABI: prt_comm_unique_id on rank 0, the 128 bytes broadcast through the group,
    prt_comm_init on every rank).  After this, :func:`render_distributed` runs the whole frame --
    shard render, all-reduce, accumulate -- inside libprt.so (prt_render_sharded)."""
    import torch.distributed as dist
    if device is None:
        device = default_device()
    ctx = scene.commit(device)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if ctx.comm_info()["world"] == world and world > 1:
        return ctx
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(box[0], world, rank)
    return ctx


def render_distributed(scene, camera, spp, max_depth=5, seed=1, rr_start=RR_OFF, device=None,
                       accum=None, group=None, render_fn=None, spp_begin=0, **render_kw):
    """Samples [spp_begin, spp_begin+spp) of every pixel, sharded over the ranks of ``group``
    (SURVEY 8e): every rank renders its contiguous sample range into a FRESH buffer, ONE all-reduce
    sums the shards, and the sum is added to ``accum`` (created if None) -- so a progressive or
    resumed call with a non-empty ``accum`` adds exactly ``spp`` samples on every rank.

    ``device`` defaults to this rank's device (LOCAL_RANK).  ``render_fn`` (default
    :func:`render`) exists so the sharding + reduction logic can be exercised without a GPU
    (tests/test_distributed_cpu.py plugs the CPU oracle in over gloo).
    """
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    if device is None:
        device = default_device()
    if render_fn is None and world == 1:  # one rank: the plain frame, still ONE C-ABI call
        return render(scene, camera, spp=spp, max_depth=max_depth, seed=seed, spp_begin=spp_begin,
                      rr_start=rr_start, device=device, accum=accum, **render_kw)
    if render_fn is None and world > 1:
        ctx = scene.commit(device)
        if ctx.comm_info()["world"] == world:  # init_distributed was called: the C-ABI frame
            ctx.set_camera(*camera.device_record(), aperture=float(getattr(camera, "aperture", 0.0)))
            if accum is None:
                accum = new_accum(camera, device)
            flags = (1 if render_kw.get("exact_primary", True) else 0) | (2 if render_kw.get("physical", False) else 0)
            params = ctx.render_params(seed=seed, spp_begin=spp_begin, spp_end=spp_begin + spp, max_depth=max_depth,
                                       rr_start=rr_start, light_color=render_kw.get("light_color", LIGHT_COLOR),
                                       tmin=T_MIN, tmax=T_MAX, flags=flags)
            ctx.render_sharded(params, accum, render_kw.get("stream"))
            return accum
    s0, s1 = shard_samples(spp, rank, world)
    shard = (render_fn or render)(scene, camera, spp=s1 - s0, max_depth=max_depth, seed=seed,
                                  spp_begin=spp_begin + s0, rr_start=rr_start, device=device, accum=None,
                                  **render_kw)
    if world > 1:
        dist.all_reduce(shard, op=dist.ReduceOp.SUM, group=group)
    if accum is None:
        return shard
    accum += shard
    return accum


def resolve(accum):
    """accum [h,w,4] -> mean radiance [h,w,3] (main.py:37 ``total/SAMPLES``)."""
    a = accum.detach().cpu().numpy() if hasattr(accum, "detach") else np.asarray(accum)
    n = np.maximum(a[..., 3:4], 1.0)
    return a[..., :3] / n


def reinhard_extended(img):
    """Extended Reinhard on luminance, white point = max luminance
    (reference: main_taichi.py:53-59,67-78 == tone_map.py:17-33)."""
    lum = img[..., 0] * 0.2126 + img[..., 1] * 0.7152 + img[..., 2] * 0.0722
    white = float(np.max(lum)) if lum.size else 1.0
    if not white > 0.0:
        return img.copy()
    l_new = lum * (1.0 + lum / (white * white)) / (1.0 + lum)
    scale = np.divide(l_new, lum, out=np.zeros_like(lum), where=lum > 0)
    return img * scale[..., None]


def to_image(accum, tonemap=None):
    """Mean radiance as the reference stores it: ``image[W-1-j, i]`` (main.py:55), i.e. top row
    first.  tonemap: None (main.py), "sqrt" (main_taichi.py:61-64), "reinhard"."""
    img = resolve(accum)[::-1]
    if tonemap == "sqrt":
        img = np.sqrt(np.maximum(img, 0.0))
    elif tonemap == "reinhard":
        img = reinhard_extended(np.maximum(img, 0.0))
    return np.ascontiguousarray(img)


def to_uint8(img):
    """``image*255 -> uint8`` of main.py:57-58, with a clamp instead of wrap-around."""
    return (np.clip(img, 0.0, 1.0) * 255.0).astype(np.uint8)


class PathTracer:
    """Shape of the reference class (core/tracing.py:47-50): bound to a scene, a
    depth and an image size; ``trace_image`` is the batched form of ``trace``."""

    def __init__(self, world, depth, img_w=None, img_h=None):
        self.world = world
        self.depth = depth
        self.img_w, self.img_h = img_w, img_h

    def trace(self, ro, rd, depth=None, x=0, y=0, spp=1, seed=1):
        """Radiance along one ray (or [n,3] arrays of rays): the host form of the reference's
        ``@ti.func trace(ro, rd, depth, x, y)`` (core/tracing.py:116-155).  (x, y) only seed the
        Philox stream there is no per-pixel state to index."""
        from .ray import Ray
        o = np.atleast_2d(np.asarray(ro, np.float64))
        d = np.atleast_2d(np.asarray(rd, np.float64))
        rays = [Ray(o[i], d[i]) for i in range(o.shape[0])]
        e, r = path_tracing(rays, self.world, spp=spp, max_depth=self.depth if depth is None else depth,
                            seed=seed + 7919 * int(x) + 104729 * int(y))
        out = e + r
        return out[0] if np.asarray(ro).ndim == 1 else out

    def trace_image(self, camera, spp=1, seed=1, spp_begin=0, accum=None, device=0):
        return render(self.world, camera, spp=spp, max_depth=self.depth, seed=seed,
                      spp_begin=spp_begin, accum=accum, device=device)
