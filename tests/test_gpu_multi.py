"""Multi-GPU path through the C ABI: sample-sharded render with ONE NCCL all-reduce per
frame (prt_comm_* / prt_render_sharded behind core.tracing.init_distributed / render_distributed).
Needs >= 2 GPUs (skipped otherwise): the N-GPU image must equal the 1-GPU image up to fp32
summation order, a pre-filled accumulation buffer must get exactly the new samples added, and
prt_allreduce_sum must sum."""
import os
import socket

import numpy as np
import pytest

from conftest import SCENE_JSON

pytestmark = pytest.mark.gpu

W = H = 96
SPP, DEPTH, SEED = 12, 6, 3


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # only carries the 128-byte NCCL id
    from pyrenderer_b200.core import tracing
    from pyrenderer_b200.core.camera import Camera
    from pyrenderer_b200.io_utils.read_tungsten import read_file
    scene, cam0 = read_file(SCENE_JSON)
    cam = Camera(cam0.position, cam0.looking_at, cam0.up, [W, H], fov=cam0.fov)
    ctx = tracing.init_distributed(scene)
    info = ctx.comm_info()
    assert info["world"] == world and info["rank"] == rank and info["nccl_version"] >= 22000
    # prt_allreduce_sum: in-place fp32 sum over the ranks
    t = torch.full((1000,), float(rank + 1), dtype=torch.float32, device=f"cuda:{rank}")
    ctx.allreduce_sum(t)
    torch.cuda.synchronize()
    assert torch.all(t == sum(range(1, world + 1)))
    acc = tracing.render_distributed(scene, cam, SPP, max_depth=DEPTH, seed=SEED)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"acc{rank}.npy"), acc.cpu().numpy())
    more = tracing.render_distributed(scene, cam, 5, max_depth=DEPTH, seed=SEED, spp_begin=SPP, accum=acc.clone())
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"more{rank}.npy"), more.cpu().numpy())
    if rank == 0:  # the single-GPU image of the same samples, same context
        one = tracing.render(scene, cam, spp=SPP + 5, max_depth=DEPTH, seed=SEED, device=0)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "one.npy"), one.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_n_gpu_sharded_render_equals_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = [np.load(tmp_path / f"acc{r}.npy") for r in range(world)]
    m = [np.load(tmp_path / f"more{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(a[0], a[r]) and np.array_equal(m[0], m[r])  # all-reduce: same buffer on every rank
    assert np.all(a[0][..., 3] == SPP) and np.all(m[0][..., 3] == SPP + 5)
    one = np.load(tmp_path / "one.npy")
    scale = np.abs(one[..., :3]).mean()
    err = np.abs(m[0][..., :3] - one[..., :3]).max() / scale
    print(f"[multi-gpu] world {world}: max |sharded - single| / mean = {err:.2e}")
    assert err < 1e-6 * (SPP + 5)  # same paths, fp32 sums in a different order
