"""Multi-rank batch export host logic on CPU: world_size-2 gloo, a fake processor standing in for
the GPU (sharding, 1-deep producer/consumer hand-off, skip-on-producer-failure, max-over-ranks)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from raw2film_b200.batch import BatchExporter


class FakeProcessor:
    """Same two-phase surface as B200Processor; 'renders' by quantising the mean."""

    def extract_image_data_cpu(self, src, **settings):
        if src is None:
            raise ValueError("unreadable frame")
        return {"image_array": src, "pipeline_resolution": src.shape[1::-1]}

    def process_preloaded(self, payload, negative_film, grain_size, grain_sigma, **settings):
        v = int(payload["image_array"].mean()) + settings.get("offset", 0)
        return np.full((2, 2, 3), v, np.uint8)


def _tasks(n):
    out = []
    for i in range(n):
        src = None if i == 5 else np.full((4, 4, 3), i, np.float32)
        out.append({"src": src, "negative_film": f"stock{i % 4}", "grain_size": 6.0, "grain_sigma": 0.4,
                    "settings": {"offset": 100}})
    return out


def test_single_rank_order_and_skip():
    got = {}
    rep = BatchExporter(FakeProcessor()).run(_tasks(9), lambda i, im: got.__setitem__(i, int(im[0, 0, 0])))
    assert rep["frames"] == [0, 1, 2, 3, 4, 6, 7, 8] and rep["skipped"] == [5]
    assert got == {i: 100 + i for i in rep["frames"]}


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = {}
    ex = BatchExporter(FakeProcessor(), world_size=world, rank=rank)
    rep = ex.run(_tasks(n), lambda i, im: got.__setitem__(i, int(im[0, 0, 0])))
    total = BatchExporter.reduce_report(rep)
    dist.barrier()
    ret[rank] = (rep["frames"], rep["skipped"], got, total)
    dist.destroy_process_group()


def test_two_ranks_gloo_partition_and_reduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n, world = 11, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    f0, s0, g0, t0 = ret[0]
    f1, s1, g1, t1 = ret[1]
    assert f0 == [0, 2, 4, 6, 8, 10] and f1 == [1, 3, 7, 9] and s1 == [5] and s0 == []
    assert sorted(list(g0) + list(g1)) == [i for i in range(n) if i != 5]
    assert t0 == t1 and t0["frames"] == 10 and t0["skipped"] == 1 and t0["seconds"] > 0
