"""Pinned host <-> device copy bandwidth of the box, per GPU and aggregate.

    python tools/micro/pcie_bw.py                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/micro/pcie_bw.py                                     # all GPUs concurrently

Copies the byte counts of one 24 MP frame of the render path: 288 MB host -> device (float32 XYZ frame) with a
72 MB device -> host copy (uint8 result) in flight, which is what the pipelined batch export moves per frame.
Under torchrun every rank drives its own GPU at the same time (barrier before the timed loop), so the aggregate
is the ceiling for the end-to-end batch numbers of bench.py at that N.
"""
import json
import os
import time

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

H2D, D2H = 288_000_000, 72_000_000
x = torch.empty(H2D, dtype=torch.uint8, pin_memory=True)
d = torch.empty(H2D, dtype=torch.uint8, device="cuda")
o = torch.empty(D2H, dtype=torch.uint8, device="cuda")
oh = torch.empty(D2H, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(n, both):
    barrier()
    t = time.perf_counter()
    for _ in range(n):
        with torch.cuda.stream(s1):
            d.copy_(x, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                oh.copy_(o, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n


run(3, True)
alone, both = run(20, False), run(20, True)
vals = torch.tensor([alone, both], dtype=torch.float64, device="cuda")
if world > 1:
    worst = vals.clone()
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
else:
    worst = vals
if rank == 0:
    a, b = float(worst[0]), float(worst[1])
    print(json.dumps({
        "gpus": world, "h2d_bytes": H2D, "d2h_bytes": D2H,
        "h2d_alone_ms_max_rank": a * 1e3, "h2d_alone_gbs_per_gpu": H2D / a / 1e9, "h2d_alone_gbs_aggregate": world * H2D / a / 1e9,
        "h2d_with_d2h_ms_max_rank": b * 1e3, "h2d_with_d2h_gbs_per_gpu": H2D / b / 1e9,
        "h2d_with_d2h_gbs_aggregate": world * H2D / b / 1e9,
        "frames_per_s_ceiling_f32": world / b,
        "note": "ceiling of the float32 end-to-end batch: one 288 MB H2D + one 72 MB D2H per 24 MP frame and GPU"}))
if world > 1:
    dist.destroy_process_group()
