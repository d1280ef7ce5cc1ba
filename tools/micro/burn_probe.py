"""Per-kernel CUDA-event times of the 24 MP full emulation with the highlight burn on (the staged tail: noise ->
grain correlation -> burn mask -> finish) next to the default fused tail.   python tools/micro/burn_probe.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from raw2film_b200 import B200Processor, _cabi  # noqa: E402
from raw2film_b200.synthetic import SyntheticStock, natural_frame  # noqa: E402

H, W = 4000, 6000
proc = B200Processor(device=0)
stock = SyntheticStock()
frame = torch.from_numpy(natural_frame(H, W, 1)).cuda()
for burn in (0.0, 0.5):
    S = dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3, highlight_burn=burn)
    for _ in range(5):
        proc.render_device(frame, stock, 6.0, 0.4, **S)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(proc.stream)
    for _ in range(10):
        proc.render_device(frame, stock, 6.0, 0.4, **S)
    e1.record(proc.stream)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * len(_cabi.PROF_NAMES))()
    n = (ctypes.c_uint64 * len(_cabi.PROF_NAMES))()
    _cabi.check(_cabi.lib.r2f_profile_read(proc._ctx, ms, n))
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 0))
    print("highlight_burn", burn, "ms/frame", round(e0.elapsed_time(e1) / 10, 3),
          {k: round(m / 10, 4) for k, m, c in zip(_cabi.PROF_NAMES, ms, n) if c})
for name, S in (("grain off", dict(halation=True, sharpness=True, grain=0, halation_green_factor=0.3)),):
    for _ in range(5):
        proc.render_device(frame, stock, 6.0, 0.4, **S)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(proc.stream)
    for _ in range(10):
        proc.render_device(frame, stock, 6.0, 0.4, **S)
    e1.record(proc.stream)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * len(_cabi.PROF_NAMES))()
    n = (ctypes.c_uint64 * len(_cabi.PROF_NAMES))()
    _cabi.check(_cabi.lib.r2f_profile_read(proc._ctx, ms, n))
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 0))
    print(name, "ms/frame", round(e0.elapsed_time(e1) / 10, 3),
          {k: round(m / 10, 4) for k, m, c in zip(_cabi.PROF_NAMES, ms, n) if c})
