#!/usr/bin/env python
"""Timeline of one synchronous `process_preloaded` call (24 MP full emulation, pinned float32 frame in, host uint8
out): when each upload band has arrived, when the render has finished each output band, when the read-back is done.

    R2F_TIMELINE=1 [R2F_CALL_BANDS=8] python tools/micro/e2e_timeline.py [calls]

Prints the median over the calls of every mark (ms after the start of the frame's upload), the wall time per call
and the same timeline for an upload with no render behind it (the PCIe floor of the call).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("R2F_TIMELINE", "1")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from raw2film_b200 import B200Processor  # noqa: E402
from raw2film_b200.synthetic import SyntheticStock, natural_frame  # noqa: E402

H, W = 4000, 6000
SETTINGS = dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3)


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    proc = B200Processor(device=0)
    stock = SyntheticStock()
    payloads = [proc.extract_image_data_cpu(natural_frame(H, W, i, out=proc.pinned_frame(H, W)), **SETTINGS)
                for i in range(2)]
    for i in range(4):
        proc.process_preloaded(payloads[i % 2], stock, 6.0, 0.4, **SETTINGS)
    pipe = proc._own_pipeline()
    marks, wall = [], []
    for i in range(calls):
        torch.cuda.synchronize()
        t = time.perf_counter()
        proc.process_preloaded(payloads[i % 2], stock, 6.0, 0.4, **SETTINGS)
        wall.append((time.perf_counter() - t) * 1e3)
        marks.append(pipe.timeline(pipe._count - 1))
    med = lambda xs: float(np.median(np.asarray(xs), axis=0)) if np.ndim(xs) == 1 else np.median(np.asarray(xs), axis=0).round(3).tolist()
    out = {"bands": pipe.bands, "wall_ms_per_call": round(med(wall), 3)}
    for k in marks[0]:
        v = [m[k] for m in marks]
        out[k] = med(v) if np.ndim(v) > 1 else round(med(v), 3)
    # the upload alone, same bands
    host = torch.from_numpy(payloads[0]["image_array"])
    dev = torch.empty_like(host, device="cuda")
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        with torch.cuda.stream(s):
            e0.record(s)
            dev.copy_(host, non_blocking=True)
            e1.record(s)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    out["h2d_alone_ms"] = round(float(np.median(ts)), 3)
    out["h2d_alone_gbs"] = round(host.numel() * 4 / np.median(ts) / 1e6, 1)
    print(json.dumps(out))
    proc.close()


if __name__ == "__main__":
    main()
