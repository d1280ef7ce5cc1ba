"""Full-emulation render time for frame sizes that are / are not multiples of 4 and 64 (cropped or rotated frames):
the aligned fast paths (TMA tiles, 128-bit loads, row-wise copies) against their fallbacks.
    python tools/micro/odd_size_probe.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from raw2film_b200 import B200Processor, _cabi  # noqa: E402
from raw2film_b200.synthetic import SyntheticStock, natural_frame  # noqa: E402

proc = B200Processor(device=0)
stock = SyntheticStock()
S = dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3)
for H, W in ((4000, 6000), (3998, 5998), (3999, 5999), (4001, 6001), (3840, 5760)):
    frame = torch.from_numpy(natural_frame(H, W, 1)).cuda()
    for _ in range(4):
        proc.render_device(frame, stock, 6.0, 0.4, **S)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(proc.stream)
    for _ in range(8):
        proc.render_device(frame, stock, 6.0, 0.4, **S)
    e1.record(proc.stream)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * len(_cabi.PROF_NAMES))()
    n = (ctypes.c_uint64 * len(_cabi.PROF_NAMES))()
    _cabi.check(_cabi.lib.r2f_profile_read(proc._ctx, ms, n))
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 0))
    print(f"{W}x{H}", "ms/frame", round(e0.elapsed_time(e1) / 8, 3),
          {k: round(m / 8, 3) for k, m, c in zip(_cabi.PROF_NAMES, ms, n) if c})
    del frame
