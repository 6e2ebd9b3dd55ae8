"""N>1 host logic on CPU: world_size-2 gloo run of core.tracing.render_distributed with the
CPU oracle plugged in as the per-rank renderer.  Sample sharding + one all-reduce must give
the single-rank image (Philox streams are keyed by the sample index, not by the rank)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import SCENE_JSON

W = H = 24
SPP, DEPTH, SEED = 6, 4, 5


def _oracle_render_fn(scene, camera, spp, max_depth, seed, spp_begin, rr_start, device, accum):
    a = scene.arrays()
    iview, sw, sh, focal, _, _ = camera.device_record()
    ocam = oracle.make_camera(iview, sh * (W / H), sh, focal, W, H)
    P = oracle.make_params(seed=seed, spp_begin=spp_begin, spp_end=spp_begin + spp, max_depth=max_depth,
                           rr_start=rr_start)
    acc, _, _ = oracle.render(a["tris"], a["normals"], a["tri_material"], a["materials"], a["light_tris"],
                              ocam, P, nthreads=2)
    out = torch.from_numpy(acc.astype(np.float32))
    return out if accum is None else accum.add_(out)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyrenderer_b200.core import tracing
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    scene, cam = read_file(SCENE_JSON)
    acc = tracing.render_distributed(scene, cam, SPP, max_depth=DEPTH, seed=SEED, render_fn=_oracle_render_fn)
    np.save(os.path.join(out_dir, f"acc{rank}.npy"), acc.numpy())
    # progressive / resumed call: a PRE-FILLED accumulation buffer gets exactly the new samples added
    # (an in-place all-reduce of the caller's buffer would multiply what is already there by `world`)
    more = tracing.render_distributed(scene, cam, 4, max_depth=DEPTH, seed=SEED, render_fn=_oracle_render_fn,
                                      spp_begin=SPP, accum=acc.clone())
    np.save(os.path.join(out_dir, f"more{rank}.npy"), more.numpy())
    dist.destroy_process_group()


def test_two_rank_sample_sharding_matches_single_rank(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a0 = np.load(tmp_path / "acc0.npy")
    a1 = np.load(tmp_path / "acc1.npy")
    assert np.array_equal(a0, a1)  # all-reduce leaves the same buffer on every rank
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    scene, cam = read_file(SCENE_JSON)
    full = _oracle_render_fn(scene, cam, SPP, DEPTH, SEED, 0, 0xFFFFFFFF, 0, None).numpy()
    assert np.all(a0[..., 3] == SPP)
    assert np.allclose(a0, full, rtol=1e-6, atol=1e-6)
    m0 = np.load(tmp_path / "more0.npy")
    assert np.array_equal(m0, np.load(tmp_path / "more1.npy"))
    full10 = _oracle_render_fn(scene, cam, SPP + 4, DEPTH, SEED, 0, 0xFFFFFFFF, 0, None).numpy()
    assert np.all(m0[..., 3] == SPP + 4)
    assert np.allclose(m0, full10, rtol=1e-6, atol=1e-6)
