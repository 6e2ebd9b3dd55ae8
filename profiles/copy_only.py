"""Host<->device copy ceiling of the e2e closest-hit path, with NO kernel: every rank moves what
prt_trace_closest_host moves per step (512 MiB of rays up, 256 MiB of hits down, 2^21-ray chunks, an
upload stream and a download stream) from / to its own pinned buffers.  Run under torchrun at
N = 1, 2, 4, 8 on one node: if the aggregate stops growing with N, the bound is the host (DRAM /
PCIe root complexes shared by the GPUs), not this library.
usage: python -m torch.distributed.run --nproc-per-node N profiles/copy_only.py"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, CH = 1 << 24, 1 << 21
rays_h = torch.empty((N, 8), dtype=torch.float32).pin_memory()
hits_h = torch.empty((N, 4), dtype=torch.float32).pin_memory()
rays_h.fill_(1.0)
rays_d = torch.empty((4 * CH, 8), dtype=torch.float32, device=dev)
hits_d = torch.zeros((4 * CH, 4), dtype=torch.float32, device=dev)
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()


def step(up=True, down=True):
    for i in range(N // CH):
        slot = (i % 4) * CH
        if up:
            with torch.cuda.stream(s_up):
                rays_d[slot:slot + CH].copy_(rays_h[i * CH:(i + 1) * CH], non_blocking=True)
        if down:
            with torch.cuda.stream(s_down):
                hits_h[i * CH:(i + 1) * CH].copy_(hits_d[slot:slot + CH], non_blocking=True)
    torch.cuda.synchronize()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


out = {"n_gpus": world}
for name, kw, nbytes in (("both", {}, N * 48), ("h2d_only", {"down": False}, N * 32), ("d2h_only", {"up": False}, N * 16)):
    step(**kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        step(**kw)
    barrier()
    dt = (time.perf_counter() - t0) / 5
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    out[name] = {"ms_per_step": dt * 1e3, "GBps_per_gpu": nbytes / dt / 1e9, "GBps_total": world * nbytes / dt / 1e9,
                 "equivalent_Mrays_per_s_total": world * N / dt / 1e6}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
