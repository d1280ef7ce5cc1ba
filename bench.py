#!/usr/bin/env python
"""bench.py -- headline benchmark of the raw2film render path on B200.

  python bench.py --gpus N --steps K --warmup W [--config C2] [--impl reference]

One "step" = one pass of the hot path over one synthetic frame of the named config
(default C2: 6000x4000 linear XYZ, full emulation: halation 43x43, MTF 17x17x3, RGB grain).
Metric: megapixels/s (whole job, all ranks).  `value` is device-resident (frames already in HBM,
inputs > L2 and rotated between steps); `e2e` goes through the public API
`B200Processor.process_preloaded` with pinned HOST buffers (H2D + render + D2H inside the timed
region).  Multi-GPU (torchrun, one rank per GPU): frames are independent, every rank renders its
own K frames (weak scaling), no data-path collective; timing = max over ranks of CUDA-event time.

`--impl reference` times the reference's CPU algorithm (the oracle port: cv2.filter2D + OpenMP C
restatement, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("NUMBA_THREADING_LAYER", "workqueue")

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (H, W, settings, description)
    "C1": (4000, 6000, dict(halation=False, sharpness=False, grain=0),
           "24MP 6000x4000 pointwise (halation/MTF/grain off)"),
    "C2": (4000, 6000, dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3),
           "24MP 6000x4000 full emulation (halation 43x43, MTF 17x17x3, RGB grain 6um)"),
    "C3": (6336, 9504, dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3, halation_size=2.0),
           "61MP 9504x6336 full emulation, halation_size=2 (133x133), MTF 27x27x3"),
    "C5": (1080, 1920, dict(halation=False, sharpness=False, grain=0),
           "2MP 1920x1080 simplified preview (pointwise)"),
    # not a BASELINE config: the interactive preview with every effect left on (what the GUI renders by default)
    "C5F": (1080, 1920, dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3),
            "2MP 1920x1080 full-emulation preview (halation 15x15, MTF 5x5x3, RGB grain)"),
    # half_size decode of a 24 MP camera (the GUI's default ingest), every effect on
    "C6": (2000, 3000, dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3),
           "6MP 3000x2000 full emulation (half-size decode)"),
}
ALG_BYTES_PER_PX = {"C1": 15, "C2": 63, "C3": 63, "C5": 15, "C5F": 63, "C6": 63}
# per-kernel algorithmic bytes/px: own input + output tensors (SURVEY 8d "same rule per kernel")
KERNEL_BYTES_PER_PX = {"pointwise": 15, "expose": 24, "halation": 24, "density": 24, "mtf": 24, "noise": 12,
                       "grain": 15, "burn": 4, "finish": 15,
                       # FFT halation: 12 B/px frame read + 8 B/px spectrum write | spectrum r+w + 4 B/px kernel
                       # spectrum | spectrum read + frame re-read + 12 B/px density write (padding excluded)
                       # (rows_fwd also leaves the 12 B/px exposure planes behind for rows_inv)
                       "fft_rows_fwd": 32, "fft_cols": 20, "fft_rows_inv": 32}
KERNEL_SASS_NAME = {"pointwise": "k_pointwise", "mtf": "k_conv2d_sym", "halation": "k_conv2d_sym", "grain": "k_grain_finish",
                    "fft_rows_fwd": "k_fft_rows_fwd", "fft_cols": "k_fft_cols", "fft_rows_inv": "k_fft_rows_inv",
                    "finish": "k_finish", "expose": "k_expose", "noise": "k_noise", "density": "k_conv2d"}


def load_traffic(config, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[config][kernel]
    except Exception:  # noqa: BLE001
        return None
GRAIN_SIZE, GRAIN_SIGMA = 6.0, 0.4                                  # gui.py:498, 509


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def cpu_render_once(fo, xyz, stock, settings):
    from tests.helpers import oracle_render

    return oracle_render(fo, xyz, stock, GRAIN_SIZE, GRAIN_SIGMA, settings)


def cpu_sample(cfg_name, H, W, settings, frac=4):
    """Bounded CPU sample: a horizontal band of the frame rendered with the full frame's px/mm
    (so the halation / MTF kernels have the production size)."""
    from raw2film_b200.synthetic import natural_frame

    rows = H if cfg_name in ("C1", "C5") else max(256, H // frac)
    scale_px_mm = max(H, W) / 36.0
    xyz = natural_frame(H, W, 0)[:rows].copy()
    st = dict(settings)
    st["frame_width"] = max(rows, W) / scale_px_mm
    st["frame_height"] = st["frame_width"] * 2 / 3
    return xyz, st, f"{W}x{rows} band of the {W}x{H} natural frame 0, kernels at the full-frame {scale_px_mm:.1f} px/mm"


def run_reference(args, world, rank):
    """--impl reference: the oracle port of the reference CPU path on the host cores."""
    if rank != 0:
        return
    from oracle import film_oracle as fo
    from raw2film_b200.synthetic import SyntheticStock

    H, W, settings, desc = CONFIGS[args.config]
    stock = SyntheticStock()
    fo.use_all_host_threads()
    xyz, st, sample = cpu_sample(args.config, H, W, settings)
    mp = xyz.shape[0] * xyz.shape[1] / 1e6
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_render_once(fo, xyz, stock, st)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_render_once(fo, xyz, stock, st)
    dt = time.perf_counter() - t0
    val = mp * args.steps / dt
    cores = fo.num_threads()
    line = {
        "impl": "reference", "metric": "megapixels_per_second", "value": val, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "name": args.config},
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=3, help="distinct resident input frames rotated between steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    world, rank, local = dist_setup(args.gpus)

    if args.impl == "reference":
        run_reference(args, world, rank)
        return

    import torch
    import torch.distributed as dist

    from raw2film_b200 import B200Processor, _cabi
    from raw2film_b200.synthetic import SyntheticStock, natural_frame

    from raw2film_b200.affinity import bind_to_gpu

    affinity = bind_to_gpu(local)          # before any pinned allocation: first touch on the GPU's NUMA node
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, settings, desc = CONFIGS[args.config]
    mp = H * W / 1e6
    stock = SyntheticStock(variant=rank % 4)   # mixed stocks across ranks
    proc = B200Processor(device=local)

    # --- inputs: pinned host payloads (for e2e) and their device-resident copies (for value) ----------
    n_frames = max(2, args.frames)
    payloads, dev_frames = [], []
    for i in range(n_frames):
        frame = natural_frame(H, W, rank * 1000 + i)
        payloads.append(proc.extract_image_data_cpu(frame, **settings))
        dev_frames.append(torch.from_numpy(payloads[-1]["image_array"]).to(proc.device))
        del frame
    torch.cuda.synchronize()

    def step_device(i):
        return proc.render_device(dev_frames[i % n_frames], stock, GRAIN_SIZE, GRAIN_SIGMA, **settings)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(i)
    barrier()

    # --- timed, device resident -----------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 1))
    launches0 = proc.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(proc.stream)
    for i in range(args.steps):
        step_device(i)
    ev1.record(proc.stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = proc.launch_count - launches0
    import ctypes

    prof_ms = (ctypes.c_double * len(_cabi.PROF_NAMES))()
    prof_n = (ctypes.c_uint64 * len(_cabi.PROF_NAMES))()
    _cabi.check(_cabi.lib.r2f_profile_read(proc._ctx, prof_ms, prof_n))
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 0))

    # --- per-call latency (BASELINE config 5 asks for p50/p99): synchronous device-resident calls ----
    lat_n = 1000 if args.config in ("C5", "C5F") else 50
    lat = []
    for i in range(lat_n):
        t0 = time.perf_counter()
        step_device(i)
        proc.stream.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.sort(np.asarray(lat))
    latency = {"p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[min(len(lat) - 1, int(len(lat) * 0.99))]),
               "min_ms": float(lat[0]), "calls": lat_n,
               "what": "wall time of one synchronous B200Processor.render_device call (device-resident frame)"}

    # --- timed, end to end through the public API (pinned host in, host out) ---------------------
    # (a) synchronous per-frame call, as the reference's single export makes it
    for i in range(2):
        proc.process_preloaded(payloads[i % n_frames], stock, GRAIN_SIZE, GRAIN_SIGMA, **settings)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(proc.stream)
    sync_steps = max(3, min(args.steps, 10))
    for i in range(sync_steps):
        out = proc.process_preloaded(payloads[i % n_frames], stock, GRAIN_SIZE, GRAIN_SIGMA, **settings)
    s1.record(proc.stream)
    barrier()
    ms_sync = s0.elapsed_time(s1) / sync_steps
    # (b) batch export: PipelinedRenderer overlaps H2D / render / D2H of consecutive frames
    from raw2film_b200 import PipelinedRenderer

    pipe = PipelinedRenderer(proc, depth=3)
    checksum = 0

    def sink(idx, img):
        nonlocal checksum
        checksum += int(img[0, 0, 0])

    pipe.run((payloads[i % n_frames] for i in range(3)), stock, GRAIN_SIZE, GRAIN_SIGMA, sink=sink, **settings)
    barrier()
    h2d0, d2h0 = pipe.h2d_bytes, pipe.d2h_bytes
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(pipe.s_in)
    pipe.run((payloads[i % n_frames] for i in range(args.steps)), stock, GRAIN_SIZE, GRAIN_SIGMA, sink=sink,
             **settings)
    e1.record(pipe.s_out)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    h2d_per_step = (pipe.h2d_bytes - h2d0) // args.steps
    d2h_per_step = (pipe.d2h_bytes - d2h0) // args.steps
    # (c) same batch with the uint16 frames rawpy hands over: ingest (/65535, *gain) runs on the device
    gain16 = 2.0 ** 5
    payloads16 = []
    for p in payloads:
        f = p["image_array"]
        u16 = np.clip(f[..., :3] * (65535.0 / gain16), 0, 65535).astype(np.uint16)
        payloads16.append(proc.extract_image_data_cpu(u16, input_gain=gain16, **settings))
        del u16
    pipe.run((payloads16[i % n_frames] for i in range(3)), stock, GRAIN_SIZE, GRAIN_SIGMA, sink=sink, **settings)
    barrier()
    h2d1 = pipe.h2d_bytes
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record(pipe.s_in)
    pipe.run((payloads16[i % n_frames] for i in range(args.steps)), stock, GRAIN_SIZE, GRAIN_SIGMA, sink=sink,
             **settings)
    u1.record(pipe.s_out)
    barrier()
    ms_e2e16 = u0.elapsed_time(u1)
    h2d16_per_step = (pipe.h2d_bytes - h2d1) // args.steps
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, ms_e2e16], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, ms_e2e16 = float(t[0]), float(t[1]), float(t[2])
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln)
        launches = int(ln[0])

    if rank == 0:
        peak, peak_src = load_peaks()
        value = world * args.steps * mp / (ms_total / 1e3)
        e2e_value = world * args.steps * mp / (ms_e2e / 1e3)
        kernels = {}
        for name, ms, n in zip(_cabi.PROF_NAMES, prof_ms, prof_n):
            if n:
                kernels[name] = {"ms": ms / n, "launches_per_step": n / args.steps,
                                 "gbs": KERNEL_BYTES_PER_PX[name] * H * W / (ms / n * 1e-3) / 1e9}
        dom = max(kernels, key=lambda k: kernels[k]["ms"] * kernels[k]["launches_per_step"]) if kernels else None
        roofline = None
        if dom:
            ach = kernels[dom]["gbs"]
            roofline = {"bound": "hbm", "kernel": dom, "sass_name": KERNEL_SASS_NAME.get(dom), "achieved": ach,
                        "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": load_traffic(args.config, dom),
                        "peak_source": peak_src,
                        "ms_per_launch": kernels[dom]["ms"],
                        "share_of_step": kernels[dom]["ms"] * kernels[dom]["launches_per_step"] / (ms_total / args.steps),
                        "step_alg_gbs": ALG_BYTES_PER_PX[args.config] * H * W / (ms_total / args.steps * 1e-3) / 1e9,
                        "step_frac": ALG_BYTES_PER_PX[args.config] * H * W / (ms_total / args.steps * 1e-3) / 1e9 / peak}
            # direct correlations are FP32 bound: report their algorithmic flop rate (2*k*k per px and filtered layer)
            for name in ("halation", "mtf"):
                k = getattr(proc, name + "_kernel", None) if name in kernels else None
                if k is not None:
                    taps = sum(int(np.count_nonzero(k[..., c])) > 1 for c in range(3)) * k.shape[0] * k.shape[1]
                    kernels[name]["alg_fp32_tflops"] = 2.0 * taps * H * W / (kernels[name]["ms"] * 1e-3) / 1e12
            if dom in ("halation", "mtf") and "alg_fp32_tflops" in kernels[dom]:
                roofline["fp32_tflops"] = kernels[dom]["alg_fp32_tflops"]
                roofline["note"] = ("direct 2-D correlation is FP32-FMA bound, not HBM bound (SURVEY 8d); fp32_tflops "
                                    "counts the algorithmic 2*k*k flops/px/layer, the y-symmetric kernel executes "
                                    "about half of them")
            elif dom == "grain":
                roofline["note"] = ("fused grain + tetrahedral LUT + quantise: ~470 instructions per pixel for 15 B/px, "
                                    "bound by instruction issue and L1TEX, not HBM (profiles/)")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import film_oracle as fo

            fo.use_all_host_threads()
            xyz, st, sample = cpu_sample(args.config, H, W, settings, frac=2)
            cpu_render_once(fo, xyz[:64].copy(), stock, st)          # warm caches / thread pools
            t0 = time.perf_counter()
            cpu_render_once(fo, xyz, stock, st)
            dt = time.perf_counter() - t0
            cpu = {"value": xyz.shape[0] * xyz.shape[1] / 1e6 / dt, "unit": "MP/s", "cores": fo.num_threads(),
                   "kind": "port", "sample": sample, "seconds": dt}
        line = {
            "metric": "megapixels_per_second", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "name": args.config, "frames_per_s": value / mp,
                       "l2": f"inputs larger than L2: {n_frames} distinct {H * W * 12 / 1e6:.0f} MB frames rotated",
                       "stocks": "variant = rank % 4 (mixed stocks across GPUs)"},
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": h2d_per_step,
                    "d2h_bytes_per_step": d2h_per_step, "ms_per_step": ms_e2e / args.steps,
                    "api": "PipelinedRenderer.run over extract_image_data_cpu payloads (pinned host float32 in, "
                           "host uint8 out, H2D/render/D2H of consecutive frames overlapped, depth 3)",
                    "sync_call_ms": ms_sync,
                    "sync_call_api": "B200Processor.process_preloaded, one frame at a time"},
            "e2e_u16": {"value": world * args.steps * mp / (ms_e2e16 / 1e3), "unit": "MP/s",
                        "h2d_bytes_per_step": h2d16_per_step, "d2h_bytes_per_step": d2h_per_step,
                        "ms_per_step": ms_e2e16 / args.steps,
                        "note": "same batch from uint16 XYZ frames (what rawpy hands over); /65535 and exposure "
                                "gain applied on the device (SURVEY 8f-1)"},
            "latency": latency, "affinity": affinity, "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "clocks": clocks,
            "checksum": checksum,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    proc.close()


if __name__ == "__main__":
    main()
