"""Pipelined end-to-end rendering: host -> device copy, render and device -> host copy of
consecutive frames overlap on three CUDA streams with `depth` in-flight slots.

This is the B200 form of the reference's batch export loop (`GpuWorker.run_tasks`,
src/raw2film/gui_objects.py:65-115: the CPU phase of frame i+1 overlaps the GPU phase of
frame i through a 1-deep queue): here the overlap is pushed down to the copy engines, so that a
batch is bound by max(PCIe H2D, render, PCIe D2H) per frame instead of their sum.
"""
from __future__ import annotations

import numpy as np


class PipelinedRenderer:
    """submit(payload, ...) enqueues one frame and returns a ticket; result(ticket) blocks until that
    frame's uint8 image is in host memory.  At most `depth` frames are in flight; the host array
    returned for ticket t stays valid until ticket t + depth is submitted."""

    def __init__(self, processor, depth: int = 3):
        import torch

        self._torch = torch
        self.proc = processor
        self.depth = max(2, int(depth))
        dev = processor.device
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.s_compute = processor.stream
        self._slots = [dict(dev_in=None, dev_out=None, host_out=None, h2d=None, done=None, d2h=None)
                       for _ in range(self.depth)]
        self._count = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _slot_buffers(self, slot, shape_in, tdtype):
        torch = self._torch
        h, w, ch = shape_in
        if slot["dev_in"] is None or tuple(slot["dev_in"].shape) != (h, w, ch) or slot["dev_in"].dtype != tdtype:
            slot["dev_in"] = torch.empty((h, w, ch), dtype=tdtype, device=self.proc.device)
            slot["dev_out"] = torch.empty((h, w, 3), dtype=torch.uint8, device=self.proc.device)
            slot["host_out"] = torch.empty((h, w, 3), dtype=torch.uint8, pin_memory=True)
            for k in ("h2d", "done", "d2h"):
                slot[k] = torch.cuda.Event()
            slot["used"] = False

    def submit(self, cpu_payload, negative_film, grain_size, grain_sigma, **settings) -> int:
        torch = self._torch
        if cpu_payload.get("_canvas") is not None or (
                cpu_payload.get("_orig_resolution") is not None
                and tuple(cpu_payload["_orig_resolution"]) != tuple(cpu_payload["image_array"].shape[:2])):
            raise NotImplementedError("canvas / post-resize frames go through B200Processor.process_preloaded")
        arr = cpu_payload["image_array"]
        host = cpu_payload.get("_pinned")
        tdtype = torch.uint16 if arr.dtype == np.uint16 else torch.float32
        if host is None:  # foreign payload: stage through pinned memory (extra host copy)
            host = torch.empty(arr.shape, dtype=tdtype, pin_memory=True)
            host.numpy()[...] = arr
        ticket = self._count
        slot = self._slots[ticket % self.depth]
        self._slot_buffers(slot, arr.shape, tdtype)
        if slot["used"]:
            slot["d2h"].synchronize()          # the slot's previous result has left the device
        with torch.cuda.stream(self.s_in):
            if slot["used"]:
                self.s_in.wait_event(slot["done"])   # previous render of this slot finished reading dev_in
            slot["dev_in"].copy_(host, non_blocking=True)
            slot["h2d"].record(self.s_in)
        self.s_compute.wait_event(slot["h2d"])
        if slot["used"]:
            self.s_compute.wait_event(slot["d2h"])
        self.proc.render_device(slot["dev_in"], negative_film, grain_size, grain_sigma, out=slot["dev_out"],
                                stream=self.s_compute, sync_caller=False,
                                input_gain=cpu_payload.get("input_gain", 1.0), **settings)
        slot["done"].record(self.s_compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["done"])
            slot["host_out"].copy_(slot["dev_out"], non_blocking=True)
            slot["d2h"].record(self.s_out)
        slot["used"] = True
        slot["keep"] = host                       # keep the pinned source alive until the copy ran
        self._count += 1
        self.h2d_bytes += host.numel() * host.element_size()
        self.d2h_bytes += slot["host_out"].numel()
        return ticket

    def result(self, ticket: int) -> np.ndarray:
        if not (self._count - self.depth <= ticket < self._count):
            raise ValueError("ticket is not in flight any more")
        slot = self._slots[ticket % self.depth]
        slot["d2h"].synchronize()
        return slot["host_out"].numpy()

    def run(self, payloads, negative_film, grain_size, grain_sigma, sink=None, **settings):
        """Render an iterable of payloads in order; `sink(index, image)` is called as results land."""
        pending = []
        n = 0
        for p in payloads:
            pending.append(self.submit(p, negative_film, grain_size, grain_sigma, **settings))
            n += 1
            if len(pending) >= self.depth:
                t = pending.pop(0)
                img = self.result(t)
                if sink is not None:
                    sink(t, img)
        for t in pending:
            img = self.result(t)
            if sink is not None:
                sink(t, img)
        return n
