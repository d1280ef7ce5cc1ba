"""Experiment: do two processors (two streams, two workspaces) on one GPU overlap well?"""
import sys, time
sys.path.insert(0, ".")
import torch
from raw2film_b200 import B200Processor
from raw2film_b200.synthetic import SyntheticStock, natural_frame
import bench

H, W, settings, _ = bench.CONFIGS["C2"]
stock = SyntheticStock()
procs = [B200Processor(device=0) for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2)]
frames = [torch.from_numpy(natural_frame(H, W, i)).cuda() for i in range(3)]
outs = [torch.empty((H, W, 3), dtype=torch.uint8, device="cuda") for _ in procs]
torch.cuda.synchronize()
def run(n):
    for i in range(n):
        p = procs[i % len(procs)]
        p.render_device(frames[i % 3], stock, 6.0, 0.4, out=outs[i % len(procs)], sync_caller=False, **settings)
run(6); torch.cuda.synchronize()
for trial in range(3):
    t0 = time.perf_counter(); run(40); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(len(procs), "streams:", round(dt / 40 * 1e3, 3), "ms/frame")
