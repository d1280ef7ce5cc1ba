"""Batch export across the GPUs of one box: whole frames are sharded one-per-GPU, no data-path
collective (frames are independent).

Shape taken from the reference's batch workers (src/raw2film/gui_objects.py:65-115 `GpuWorker`:
a producer thread runs the CPU phase `extract_image_data_cpu` into a 1-deep queue, the consumer
runs `process_preloaded` and saves): here every rank (one process per GPU, torch.distributed for
rendezvous / barrier / max-over-ranks timing only) runs that producer/consumer pair over its own
shard `frame i -> rank i mod world_size` (settings.shard_frames).  A producer failure becomes a
`None` payload that the consumer skips, like gui_objects.py:86-87, 101-103.

The consumer does not wait for a frame before it takes the next: it submits to a `PipelinedRenderer`
(host -> device copy, render and device -> host copy of consecutive frames overlap) and hands finished
frames to the sink in submission order.  Tasks may name different film stocks and settings (BASELINE
config 4): each stock owns a device table slot, so alternating stocks uploads nothing.
"""
from __future__ import annotations

import queue
import threading
import time
from typing import Callable, Sequence

from .settings import shard_frames


class BatchExporter:
    def __init__(self, processor, world_size: int = 1, rank: int = 0, queue_depth: int = 1, pipeline_depth: int = 3):
        self.processor = processor
        self.world_size, self.rank = int(world_size), int(rank)
        self.queue_depth = max(1, int(queue_depth))
        self.pipeline_depth = max(2, int(pipeline_depth))
        self._pipe = None

    def _pipeline(self):
        """The overlapped renderer; processors without a CUDA pipeline (tests with a stand-in processor) fall
        back to their synchronous `process_preloaded`."""
        if self._pipe is None and hasattr(self.processor, "device") and hasattr(self.processor, "render_device"):
            from .pipeline import PipelinedRenderer

            self._pipe = PipelinedRenderer(self.processor, depth=self.pipeline_depth)
        return self._pipe

    # -- one rank's shard ------------------------------------------------------------------
    def run(self, tasks: Sequence[dict], sink: Callable[[int, object], None]) -> dict:
        """tasks[i] = {"src": array | path, "negative_film": stock, "grain_size": .., "grain_sigma": ..,
        "settings": {...}}.  `sink(frame_index, uint8 image)` receives every finished frame of this rank, in
        submission order; the image is a view of a pinned buffer that is reused `pipeline_depth` frames later.
        Returns {"frames": [...], "skipped": [...], "seconds": wall time of this rank,
        "phase1_seconds": time the producer spent in extract_image_data_cpu}."""
        mine = shard_frames(len(tasks), self.world_size, self.rank)
        q: queue.Queue = queue.Queue(maxsize=self.queue_depth)
        phase1 = [0.0]

        def producer():
            for idx in mine:
                t = tasks[idx]
                p0 = time.perf_counter()
                try:
                    payload = self.processor.extract_image_data_cpu(t["src"], **t.get("settings", {}))
                except Exception:  # noqa: BLE001 - same policy as the reference's producer
                    payload = None
                phase1[0] += time.perf_counter() - p0
                q.put((idx, payload))
            q.put(None)

        th = threading.Thread(target=producer, daemon=True)
        t0 = time.perf_counter()
        th.start()
        pipe = self._pipeline()
        done, skipped, pending = [], [], []

        def drain(limit):
            while len(pending) > limit:
                idx, ticket = pending.pop(0)
                sink(idx, pipe.result(ticket))
                done.append(idx)

        while True:
            item = q.get()
            if item is None:
                break
            idx, payload = item
            if payload is None:
                skipped.append(idx)
                continue
            t = tasks[idx]
            if pipe is None:
                image = self.processor.process_preloaded(payload, t["negative_film"], t["grain_size"],
                                                         t["grain_sigma"], **t.get("settings", {}))
                sink(idx, image)
                done.append(idx)
                continue
            ticket = pipe.submit(payload, t["negative_film"], t["grain_size"], t["grain_sigma"],
                                 **t.get("settings", {}))
            pending.append((idx, ticket))
            drain(self.pipeline_depth - 1)
        if pipe is not None:
            drain(0)
        th.join()
        return {"frames": done, "skipped": skipped, "seconds": time.perf_counter() - t0,
                "phase1_seconds": phase1[0]}

    # -- cross-rank bookkeeping (no image data crosses ranks) -------------------------------------
    @staticmethod
    def reduce_report(report: dict, device=None) -> dict:
        """max-over-ranks time and summed frame counts through torch.distributed (gloo or nccl)."""
        import torch
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            return {"seconds": report["seconds"], "frames": len(report["frames"]), "skipped": len(report["skipped"])}
        dev = device if device is not None else "cpu"
        t = torch.tensor([report["seconds"]], dtype=torch.float64, device=dev)
        n = torch.tensor([len(report["frames"]), len(report["skipped"])], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        return {"seconds": float(t[0]), "frames": int(n[0]), "skipped": int(n[1])}
