import sys, os, ctypes, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from raw2film_b200 import B200Processor, _cabi
from raw2film_b200.synthetic import SyntheticStock, natural_frame
H, W = 4000, 6000
S = dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3)
proc = B200Processor(device=0); stock = SyntheticStock()
frame = torch.from_numpy(natural_frame(H, W, 1)).cuda()
for nb in (1, 8):
    in_ev = [torch.cuda.Event() for _ in range(nb)]; out_ev = [torch.cuda.Event() for _ in range(nb)]
    for e in in_ev + out_ev: e.record(proc.stream)
    bands = (in_ev, out_ev) if nb > 1 else None
    for _ in range(5): proc.render_device(frame, stock, 6.0, 0.4, bands=bands, **S)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(proc.stream)
    for _ in range(10): proc.render_device(frame, stock, 6.0, 0.4, bands=bands, **S)
    e1.record(proc.stream); torch.cuda.synchronize()
    ms = (ctypes.c_double * len(_cabi.PROF_NAMES))(); n = (ctypes.c_uint64 * len(_cabi.PROF_NAMES))()
    _cabi.check(_cabi.lib.r2f_profile_read(proc._ctx, ms, n)); _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 0))
    print(nb, round(e0.elapsed_time(e1) / 10, 3), {k: (round(m / 10, 4), int(c) // 10) for k, m, c in zip(_cabi.PROF_NAMES, ms, n) if c})
