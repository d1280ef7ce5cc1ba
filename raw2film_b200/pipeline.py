"""Pipelined end-to-end rendering: host -> device copy, render and device -> host copy of
consecutive frames overlap on three CUDA streams with `depth` in-flight slots.

This is the B200 form of the reference's batch export loop (`GpuWorker.run_tasks`,
src/raw2film/gui_objects.py:65-115: the CPU phase of frame i+1 overlaps the GPU phase of
frame i through a 1-deep queue): here the overlap is pushed down to the copy engines, so that a
batch is bound by max(PCIe H2D, render, PCIe D2H) per frame instead of their sum.

Consecutive frames may use different film stocks and settings (BASELINE config 4, "mixed stocks"):
the processor keeps one device table slot per stock and every table upload is copy-on-write
(include/r2f_b200.h), so a frame in flight never sees tables that a later submit changed.

`B200Processor.process_preloaded` is one submit + result on a three-slot instance of this class.
"""
from __future__ import annotations

import numpy as np

from . import _cabi, hostops


class PipelinedRenderer:
    """submit(payload, ...) enqueues one frame and returns a ticket; result(ticket) blocks until that
    frame's uint8 image is in host memory.  At most `depth` frames are in flight; the host array
    returned for ticket t stays valid until ticket t + depth is submitted."""

    def __init__(self, processor, depth: int = 3, bands: int = 1):
        """`bands` > 1: every frame is copied in and out in that many horizontal bands, the first kernel of the
        render starts on a band as soon as it has arrived and the result of a band leaves while the next one is
        computed (r2f_render_banded).  That shortens a frame's own latency -- what a synchronous caller sees
        (`process_preloaded`) -- and is pointless for throughput, where neighbouring frames already overlap."""
        import torch

        self._torch = torch
        self.proc = processor
        self.depth = max(2, int(depth))
        self.bands = max(1, min(int(bands), 16))
        dev = processor.device
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.s_compute = processor.stream
        self._slots = [dict(dev_in=None, dev_out=None, host_out=None, canvas_dev=None, used=False)
                       for _ in range(self.depth)]
        self._count = 0
        self._last_slot = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        import os

        self._timed = bool(os.environ.get("R2F_TIMELINE"))   # tools/micro/e2e_timeline.py: events that carry times

    def timeline(self, ticket: int):
        """R2F_TIMELINE=1 only: milliseconds after the start of the frame's upload at which each upload band had
        arrived, the render had finished each output band, the whole render and the whole read-back were done."""
        slot = self._slots[ticket % self.depth]
        slot["d2h"].synchronize()
        t0 = slot["t0"]
        out = {"h2d_done": t0.elapsed_time(slot["h2d"]), "render_done": t0.elapsed_time(slot["done"]),
               "d2h_done": t0.elapsed_time(slot["d2h"])}
        if slot.get("in_ev"):
            out["in_bands"] = [t0.elapsed_time(e) for e in slot["in_ev"]]
            out["out_bands"] = [t0.elapsed_time(e) for e in slot["out_ev"]]
        return out

    def _slot_buffers(self, slot, shape_in, tdtype, out_shape, render_hw):
        torch = self._torch
        h, w, ch = shape_in
        dev = self.proc.device
        if slot["dev_in"] is None or tuple(slot["dev_in"].shape) != (h, w, ch) or slot["dev_in"].dtype != tdtype:
            slot["dev_in"] = torch.empty((h, w, ch), dtype=tdtype, device=dev)
        if slot["dev_out"] is None or tuple(slot["dev_out"].shape[:2]) != tuple(render_hw):
            slot["dev_out"] = torch.empty((render_hw[0], render_hw[1], 3), dtype=torch.uint8, device=dev)
        if slot["host_out"] is None or tuple(slot["host_out"].shape) != tuple(out_shape):
            slot["host_out"] = torch.empty(tuple(out_shape), dtype=torch.uint8, pin_memory=True)
            slot["canvas_dev"] = None
        if "h2d" not in slot:
            for k in ("h2d", "done", "d2h", "t0"):
                slot[k] = torch.cuda.Event(enable_timing=self._timed)

    def submit(self, cpu_payload, negative_film, grain_size, grain_sigma, upload: bool = True, readback: bool = True,
               **settings) -> int:
        """Enqueue one frame.  `upload=False` renders the frame that the previous submit left on the device
        again (interactive re-render with changed settings: no host -> device copy)."""
        torch = self._torch
        proc = self.proc
        arr = cpu_payload["image_array"]
        tdtype = torch.uint16 if arr.dtype == np.uint16 else torch.float32
        pre = cpu_payload.get("_pre_resize")          # device resolution_scaling before the path
        h, w = arr.shape[:2] if pre is None else pre
        canvas = cpu_payload.get("_canvas")
        out_shape = (h, w, 3) if canvas is None else (canvas["size"][0], canvas["size"][1], 3)
        orig = cpu_payload.get("_orig_resolution")
        post = None if orig is None else hostops.target_size(out_shape, orig)   # ... and after it
        if post is not None:
            out_shape = (post[0], post[1], 3)
        ticket = self._count
        slot = self._slots[ticket % self.depth]
        band_events, rows, nb = None, None, 1
        if slot["used"]:
            slot["d2h"].synchronize()          # the slot's previous result has left the device
        if upload:
            host = cpu_payload.get("_pinned")
            if host is None:  # foreign payload: stage through pinned memory (extra host copy)
                host = torch.empty(arr.shape, dtype=tdtype, pin_memory=True)
                host.numpy()[...] = arr
            self._slot_buffers(slot, arr.shape, tdtype, out_shape, (h, w))
            nb = 1
            if pre is None and canvas is None and post is None and readback:
                nb = max(1, min(self.bands, h // 256))    # bands of at least 256 rows
            rows = [int(_cabi.lib.r2f_band_row(h, nb, i)) for i in range(nb + 1)]
            if nb > 1 and len(slot.get("in_ev", ())) != nb:
                slot["in_ev"] = [torch.cuda.Event(enable_timing=self._timed) for _ in range(nb)]
                slot["out_ev"] = [torch.cuda.Event(enable_timing=self._timed) for _ in range(nb)]
                for e in slot["in_ev"] + slot["out_ev"]:
                    e.record(self.s_in)               # creates the underlying cudaEvent_t
            with torch.cuda.stream(self.s_in):
                if slot.get("last_read") is not None:
                    self.s_in.wait_event(slot["last_read"])   # the last render that read this dev_in has finished
                slot["t0"].record(self.s_in)
                if nb > 1:
                    for i in range(nb):
                        slot["dev_in"][rows[i]:rows[i + 1]].copy_(host[rows[i]:rows[i + 1]], non_blocking=True)
                        slot["in_ev"][i].record(self.s_in)
                else:
                    slot["dev_in"].copy_(host, non_blocking=True)
                slot["h2d"].record(self.s_in)
            if nb > 1:
                band_events = (slot["in_ev"], slot["out_ev"])  # the render waits per band
            else:
                self.s_compute.wait_event(slot["h2d"])
            slot["keep"] = host                       # keep the pinned source alive until the copy ran
            self.h2d_bytes += host.numel() * host.element_size()
            dev_in = slot["dev_in"]
            reader = slot
        else:
            prev = self._last_slot
            if prev is None or prev["dev_in"] is None or tuple(prev["dev_in"].shape) != tuple(arr.shape):
                raise RuntimeError("upload=False needs the same frame to be on the device from the previous submit")
            dev_in = prev["dev_in"]                   # same compute stream: ordered after the previous render
            reader = prev
            self._slot_buffers(slot, arr.shape, tdtype, out_shape, (h, w))
        if slot["used"]:
            self.s_compute.wait_event(slot["d2h"])
        proc.output_resolution = cpu_payload.get("output_resolution")
        proc.canvas_resolution = cpu_payload.get("canvas_resolution")
        proc.pipeline_resolution = cpu_payload.get("pipeline_resolution")
        if pre is not None:                           # resolution_scaling of the float frame, on the device
            if tdtype != torch.float32:
                raise ValueError("a frame that is resized before the path must be float32")
            slot["dev_pre"] = proc.resize_device(dev_in, pre, out=slot.get("dev_pre"), stream=self.s_compute)
            dev_in = slot["dev_pre"]
        out_dev = proc.render_device(dev_in, negative_film, grain_size, grain_sigma, out=slot["dev_out"],
                                     stream=self.s_compute, sync_caller=False,
                                     input_gain=cpu_payload.get("input_gain", 1.0), bands=band_events, **settings)
        if canvas is not None:                        # add_canvas (cpu_processor.py:409) on the device
            ch_, cw_ = canvas["size"]
            if slot["canvas_dev"] is None or tuple(slot["canvas_dev"].shape) != (ch_, cw_, 3):
                slot["canvas_dev"] = torch.empty((ch_, cw_, 3), dtype=torch.uint8, device=proc.device)
            r, g, b = canvas["colour"]
            _cabi.check(_cabi.lib.r2f_canvas_paste(proc._ctx, out_dev.data_ptr(), h, w, slot["canvas_dev"].data_ptr(),
                                                   ch_, cw_, int(canvas["offset"][0]), int(canvas["offset"][1]),
                                                   r, g, b, self.s_compute.cuda_stream))
            out_dev = slot["canvas_dev"]
        if post is not None:                          # post-step of cpu_processor.py:411-412, on the device
            slot["dev_post"] = proc.resize_device(out_dev, post, out=slot.get("dev_post"), stream=self.s_compute)
            out_dev = slot["dev_post"]
        slot["done"].record(self.s_compute)
        reader["last_read"] = slot["done"]
        slot["result_dev"] = out_dev
        with torch.cuda.stream(self.s_out):
            if band_events is not None:               # copy each band out as soon as its rows are final
                for i in range(nb):
                    self.s_out.wait_event(slot["out_ev"][i])
                    slot["host_out"][rows[i]:rows[i + 1]].copy_(out_dev[rows[i]:rows[i + 1]], non_blocking=True)
            else:
                self.s_out.wait_event(slot["done"])
                if readback:
                    slot["host_out"].copy_(out_dev, non_blocking=True)
            slot["d2h"].record(self.s_out)
        slot["used"] = True
        self._last_slot = slot if upload else self._last_slot
        self._count += 1
        self.d2h_bytes += slot["host_out"].numel()
        return ticket

    def device_result(self, ticket: int):
        """The frame's uint8 result as a CUDA tensor (valid until the slot is reused), ordered on the compute stream."""
        if not (self._count - self.depth <= ticket < self._count):
            raise ValueError("ticket is not in flight any more")
        return self._slots[ticket % self.depth]["result_dev"]

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


class PreviewGraph:
    """One render of a fixed frame buffer, stock and settings captured as a CUDA graph: `replay()` costs one
    cudaGraphLaunch instead of the Python walk through the loaders and the kernel launches (interactive preview,
    BASELINE config 5: the 2 MP pointwise kernel takes ~35 us, comparable to the host work of a normal call).

    The caller writes new frames into `frame` (a device tensor of the captured shape) and calls `replay()`; the
    result lands in `out`.  A graph bakes the table pointers in: after any settings / stock change through the
    processor it refuses to replay -- capture a new one.  The grain seed is baked in as well (the same noise field
    on every replay), so the intended use is the simplified preview (halation / sharpness / grain off, gui.py:2206-2209)."""

    def __init__(self, processor, frame_dev, negative_film, grain_size, grain_sigma, **settings):
        import torch

        self._torch = torch
        self.proc = processor
        self.frame = frame_dev
        h, w = frame_dev.shape[:2]
        self.out = torch.empty((h, w, 3), dtype=torch.uint8, device=processor.device)
        args = (frame_dev, negative_film, grain_size, grain_sigma)
        kw = dict(out=self.out, stream=processor.stream, sync_caller=False, **settings)
        processor.render_device(*args, **kw)            # warm-up: tables, scratch, kernel attributes
        processor.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=processor.stream, capture_error_mode="thread_local"):
            processor.render_device(*args, **kw)
        self._version = processor.table_version
        self._stream = processor.stream

    def replay(self):
        if self.proc.table_version != self._version:
            raise RuntimeError("tables changed since this graph was captured: capture a new PreviewGraph")
        torch = self._torch
        caller = torch.cuda.current_stream(self.proc.device)
        self._stream.wait_stream(caller)              # the caller's writes to `frame` come first ...
        with torch.cuda.stream(self._stream):
            self.graph.replay()
        _cabi.check(_cabi.lib.r2f_stream_mark(self.proc._ctx, self._stream.cuda_stream))
        caller.wait_stream(self._stream)              # ... and its reads of `out` after the replay (no host sync)
        return self.out
