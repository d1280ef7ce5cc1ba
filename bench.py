#!/usr/bin/env python
"""bench.py -- headline benchmark of the raw2film render path on B200.

  python bench.py --gpus N --steps K --warmup W [--config C2] [--impl reference]

One "step" = one pass of the hot path over one synthetic frame of the named config
(default C2: 6000x4000 linear XYZ, full emulation: halation 43x43, MTF 17x17x3, RGB grain).
Metric: megapixels/s (whole job, all ranks).  `value` is device-resident (frames already in HBM,
inputs > L2 and rotated between steps).  `e2e` is the reference's own entry point for preloaded frames,
`B200Processor.process_preloaded` (gpu_processor.py:1643-1693), called once per frame with pinned HOST
buffers: H2D + render + D2H inside the timed region, nothing overlapped across calls.  `e2e_pipelined` /
`e2e_u16` are the batch-export form (`PipelinedRenderer`: copies and renders of consecutive frames overlap),
and `c4` is BASELINE config 4 itself: 64 frames, stock = frame index % 4, frame -> rank index % world, through
`BatchExporter`.  Multi-GPU (torchrun, one rank per GPU): frames are independent, every rank renders its
own K frames (weak scaling), no data-path collective; timing = max over ranks of CUDA-event time.

`--impl reference` times the reference's CPU algorithm (the oracle port: cv2.filter2D + OpenMP C
restatement, all host threads) on the same config: the whole frame per step when K + W steps of it fit
in a few minutes, else a horizontal band of it (kernels at the full frame's px/mm).
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
KERNEL_SASS_NAME = {"pointwise": "k_pointwise_fast", "mtf": "k_conv2d_sym", "halation": "k_conv2d_sym", "grain": "k_grain_finish_sym",
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
FFMA_PEAK_TFLOPS = 73.0      # measured: tools/micro/ffma2_rate.cu, profiles/r01_ffma2_rate.txt


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


def config_dict(name):
    """The `config` object of the JSON line: static, identical for both arms (`--impl b200` / `reference`)."""
    H, W, settings, desc = CONFIGS[name]
    return {"workload": desc, "name": name, "frame": f"{W}x{H}x3 float32 linear XYZ",
            "settings": {k: settings[k] for k in sorted(settings)},
            "stock": "SyntheticStock variant = rank % 4 (mixed stocks across GPUs), n2/N/n3 = 64/1024/33",
            "l2": f"inputs larger than L2: distinct {H * W * 12 / 1e6:.0f} MB frames rotated between steps"}


def cpu_sample(cfg_name, H, W, settings, rows=None):
    """CPU sample of the workload: the whole frame, or (`rows`) a horizontal band of it rendered with the full
    frame's px/mm so the halation / MTF kernels keep their production size."""
    from raw2film_b200.synthetic import natural_frame

    rows = H if rows is None else max(64, min(H, int(rows)))
    scale_px_mm = max(H, W) / 36.0
    xyz = natural_frame(H, W, 0)[:rows].copy()
    st = dict(settings)
    st["frame_width"] = max(rows, W) / scale_px_mm
    st["frame_height"] = st["frame_width"] * 2 / 3
    what = "whole frame" if rows == H else f"{W}x{rows} band"
    return xyz, st, f"{what} of the {W}x{H} natural frame 0, kernels at the full-frame {scale_px_mm:.1f} px/mm"


def run_reference(args, world, rank):
    """--impl reference: the oracle port of the reference CPU path on the host cores, same config dict."""
    if rank != 0:
        return
    from oracle import film_oracle as fo
    from raw2film_b200.synthetic import SyntheticStock

    H, W, settings, desc = CONFIGS[args.config]
    stock = SyntheticStock()
    fo.use_all_host_threads()
    # calibrate on a 1/16 band, then take the whole frame if warm-up + K steps of it stay within ~3 minutes
    xyz, st, _ = cpu_sample(args.config, H, W, settings, rows=max(128, H // 16))
    cpu_render_once(fo, xyz[:64].copy(), stock, st)
    t0 = time.perf_counter()
    cpu_render_once(fo, xyz, stock, st)
    per_row = (time.perf_counter() - t0) / xyz.shape[0]
    warm = max(1, min(args.warmup, 1))
    budget_s = 180.0
    rows = H if per_row * H * (args.steps + warm) <= budget_s else int(budget_s / (per_row * (args.steps + warm)))
    xyz, st, sample = cpu_sample(args.config, H, W, settings, rows=rows)
    mp = xyz.shape[0] * xyz.shape[1] / 1e6
    for _ in range(warm):
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
        "config": config_dict(args.config),
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_aux(args, local):
    """CUDA-event times of the entry points either side of the path on a device-resident frame."""
    import ctypes

    import torch

    from raw2film_b200 import B200Processor, _cabi, builders, hostops
    from raw2film_b200.synthetic import SyntheticStock, natural_frame

    torch.cuda.set_device(local)
    proc = B200Processor(device=local)
    H, W = 4000, 6000
    st = proc.stream
    frame = torch.from_numpy(natural_frame(H, W, 0)).to(proc.device)
    u16 = (frame.clamp(0, 1) * 65535).to(torch.int32).to(torch.uint16)
    out8 = proc.render_device(frame, SyntheticStock(), GRAIN_SIZE, GRAIN_SIGMA, halation=False, sharpness=False,
                              grain=0).clone()
    ws = torch.empty(int(_cabi.lib.r2f_workspace_bytes(H, W, 0)), dtype=torch.uint8, device=proc.device)
    res = {}

    def timed(name, fn, bytes_moved, reps=10):
        fn()
        st.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            fn()
        b.record(st)
        st.synchronize()
        ms = a.elapsed_time(b) / reps
        res[name] = {"ms": round(ms, 4), "gbs": round(bytes_moved / ms / 1e6, 1)}

    mp = H * W
    small = torch.empty((1080, 1620, 3), dtype=torch.float32, device=proc.device)
    timed("resize_area_f32_24mp_to_1620x1080", lambda: proc.resize_device(frame, (1080, 1620), out=small), mp * 12)
    half = torch.empty((2000, 3000, 3), dtype=torch.float32, device=proc.device)
    timed("resize_area_f32_24mp_to_half", lambda: proc.resize_device(frame, (2000, 3000), out=half), mp * 15)
    src8 = out8[:2667, :4000].contiguous()
    up8 = torch.empty((H, W, 3), dtype=torch.uint8, device=proc.device)
    timed("resize_lanczos4_u8_10.7mp_to_24mp", lambda: proc.resize_device(src8, (H, 6000), out=up8), mp * 3 + src8.numel())
    taps = builders.chroma_nr_taps(3)
    cn_out = torch.empty_like(frame)
    timed("chroma_nr_size3_24mp", lambda: _cabi.check(_cabi.lib.r2f_chroma_nr(
        proc._ctx, frame.data_ptr(), 3, cn_out.data_ptr(), H, W, _cabi.f32_ptr(taps), taps.shape[0], ws.data_ptr(),
        ws.numel(), st.cuda_stream)), mp * 24)
    mean = ctypes.c_double(0)
    timed("calc_exposure_u16_24mp", lambda: _cabi.check(_cabi.lib.r2f_calc_exposure(
        proc._ctx, u16.data_ptr(), _cabi.IN_U16, H, W, 3, 3.0, ctypes.byref(mean), st.cuda_stream)), mp * 6 / 4)
    size, colour, off = hostops.canvas_geometry((H, W), "Proportional white", 1.1, 1.0)
    canvas = torch.empty((size[0], size[1], 3), dtype=torch.uint8, device=proc.device)
    timed("canvas_paste_24mp", lambda: _cabi.check(_cabi.lib.r2f_canvas_paste(
        proc._ctx, out8.data_ptr(), H, W, canvas.data_ptr(), size[0], size[1], int(off[0]), int(off[1]), 255, 255,
        255, st.cuda_stream)), mp * 3 + canvas.numel())
    mix = np.arange(32, dtype=np.uint8)
    hist = torch.empty((100, 256, 4), dtype=torch.uint8, device=proc.device)
    timed("histogram_image_24mp", lambda: _cabi.check(_cabi.lib.r2f_histogram_image(
        proc._ctx, out8.data_ptr(), H, W, mix.ctypes.data_as(ctypes.c_void_p), 100, hist.data_ptr(),
        st.cuda_stream)), mp * 3)
    widget = torch.empty((1080, 1920, 4), dtype=torch.uint8, device=proc.device)
    proc.pipeline_resolution = proc.output_resolution = (W, H)
    proc.canvas_resolution = None
    timed("present_24mp_to_1920x1080", lambda: proc.present(out8, widget), widget.numel() + 1920 * 1080 * 12)
    print(json.dumps({"aux": True, "frame": f"{W}x{H}", "what": "CUDA-event time per call, device-resident operands; gbs = "
                      "bytes the call has to move / time", "steps": res}))
    proc.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=3, help="distinct resident input frames rotated between steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c4-frames", type=int, default=64, help="frames of the mixed-stock batch leg (0 = skip)")
    ap.add_argument("--kernel-only", action="store_true", help="device-resident leg only (ncu captures)")
    ap.add_argument("--aux", action="store_true",
                    help="time the steps either side of the path (SURVEY 8f) on a device-resident 24 MP frame instead")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    world, rank, local = dist_setup(args.gpus)

    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    if args.aux:
        run_aux(args, local)
        return

    import ctypes

    import torch
    import torch.distributed as dist

    from raw2film_b200 import B200Processor, BatchExporter, PipelinedRenderer, _cabi
    from raw2film_b200.affinity import bind_to_gpu
    from raw2film_b200.synthetic import SyntheticStock, natural_frame

    affinity = bind_to_gpu(local)          # before any pinned allocation: first touch on the GPU's NUMA node
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, settings, desc = CONFIGS[args.config]
    mp = H * W / 1e6
    stock = SyntheticStock(variant=rank % 4)   # mixed stocks across ranks
    proc = B200Processor(device=local)

    # --- inputs: frames generated straight into pinned host memory (phase 1 hands them over without a copy)
    # and their device-resident copies (for `value`) ------------------------------------------------------
    n_frames = max(2, args.frames)
    payloads, dev_frames, host_frames = [], [], []
    for i in range(n_frames):
        frame = natural_frame(H, W, rank * 1000 + i, out=proc.pinned_frame(H, W))
        host_frames.append(frame)
        payloads.append(proc.extract_image_data_cpu(frame, **settings))
        dev_frames.append(torch.from_numpy(payloads[-1]["image_array"]).to(proc.device))
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
    # clock warm-up: a fresh process finds the GPU at idle clocks, and W short steps (0.1-2 ms each) end before
    # the SM clock has ramped; keep stepping (untimed) until 0.4 s have passed
    t_warm, i = time.perf_counter(), args.warmup
    while time.perf_counter() - t_warm < 0.4:
        step_device(i)
        i += 1
        if i % 16 == 0:
            proc.stream.synchronize()
    barrier()

    # --- timed, device resident -----------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 1))
    proc.fast_chain_stats()          # reset the deferred-pixel counter: the share below is of the timed steps only
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
    prof_ms = (ctypes.c_double * len(_cabi.PROF_NAMES))()
    prof_n = (ctypes.c_uint64 * len(_cabi.PROF_NAMES))()
    _cabi.check(_cabi.lib.r2f_profile_read(proc._ctx, prof_ms, prof_n))
    _cabi.check(_cabi.lib.r2f_profile_enable(proc._ctx, 0))
    if args.kernel_only:
        if rank == 0:
            sampler.stop()
            ks = {n: round(ms / k, 5) for n, ms, k in zip(_cabi.PROF_NAMES, prof_ms, prof_n) if k}
            deferred, margin = proc.fast_chain_stats()
            print(json.dumps({"ms_per_step": ms_total / args.steps, "kernel_only": True, "config": args.config,
                              "kernels_ms": ks, "fast_chain": {"margin": margin, "deferred_share": deferred / (
                                  args.steps * H * W)}}))
        proc.close()
        return

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
    if args.config == "C5":
        # (i) the same call replayed as a CUDA graph, (ii) the preview as the GUI drives it: a 24 MP frame resident
        # on the device, shrunk to the widget size (INTER_AREA, utils.py:226-244) and rendered, per call
        from raw2film_b200 import PreviewGraph

        pg = PreviewGraph(proc, dev_frames[0], stock, GRAIN_SIZE, GRAIN_SIGMA, **settings)
        lat = []
        for i in range(lat_n):
            t0 = time.perf_counter()
            pg.replay()
            proc.stream.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.sort(np.asarray(lat))
        latency["graph"] = {"p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]),
                            "min_ms": float(lat[0]), "what": "PreviewGraph.replay() + stream synchronize"}
        big = torch.from_numpy(natural_frame(4000, 6000, 77)).to(proc.device)
        small = None
        lat = []
        for i in range(200):
            t0 = time.perf_counter()
            small = proc.resize_device(big, (H, 1620), out=small)
            proc.render_device(small, stock, GRAIN_SIZE, GRAIN_SIGMA, sync_caller=False, **settings)
            proc.stream.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.sort(np.asarray(lat))
        latency["from_24mp"] = {"p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]),
                                "what": "device INTER_AREA 6000x4000 -> 1620x1080 + pointwise render, per call"}
        del small
        # (iii) the reference's own preview call (gui.py:2197-2224): process(src, resolution=widget size) on a 24 MP
        # frame whose image parameters have not changed, so the frame stays on the device and only the shrink, the
        # render and the 6 MB read-back happen per call
        src24 = natural_frame(4000, 6000, 77, out=proc.pinned_frame(4000, 6000))
        pv = dict(settings, resolution=(H, W), max_scale=None, cache=True, own_result=False)
        proc.process(src24, stock, GRAIN_SIZE, GRAIN_SIGMA, **pv)
        lat = []
        for i in range(200):
            t0 = time.perf_counter()
            proc.process(src24, stock, GRAIN_SIZE, GRAIN_SIGMA, **pv)
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.sort(np.asarray(lat))
        latency["process_call_24mp_source"] = {
            "p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]),
            "what": "B200Processor.process(frame, resolution=(1080, 1920)) with the 24 MP frame cached on the device: "
                    "INTER_AREA shrink + render + read-back of the 1620x1080 result, host array returned"}
        del big

    # --- timed, end to end through the public API (pinned host in, host out) ---------------------
    # (a) the reference's entry point for preloaded frames, one synchronous call per frame
    checksum = 0
    for i in range(3):
        proc.process_preloaded(payloads[i % n_frames], stock, GRAIN_SIZE, GRAIN_SIGMA, **settings)
    own_pipe = proc._own_pipeline()
    barrier()
    h2d0, d2h0 = own_pipe.h2d_bytes, own_pipe.d2h_bytes
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    s0.record(own_pipe.s_in)
    for i in range(args.steps):
        out = proc.process_preloaded(payloads[i % n_frames], stock, GRAIN_SIZE, GRAIN_SIGMA, **settings)
        checksum += int(out[0, 0, 0])
    s1.record(own_pipe.s_out)
    barrier()
    ms_sync = s0.elapsed_time(s1)
    ms_sync_wall = (time.perf_counter() - t_wall) * 1e3
    h2d_sync = (own_pipe.h2d_bytes - h2d0) // args.steps
    d2h_sync = (own_pipe.d2h_bytes - d2h0) // args.steps
    # (b) batch export: PipelinedRenderer overlaps H2D / render / D2H of consecutive frames
    pipe = PipelinedRenderer(proc, depth=3)

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
        u16 = proc.pinned_frame(H, W, dtype=np.uint16)
        u16[...] = np.clip(f[..., :3] * np.float32(65535.0 / gain16), 0, 65535).astype(np.uint16)
        payloads16.append(proc.extract_image_data_cpu(u16, input_gain=gain16, **settings))
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
    del payloads16, pipe

    # (d) BASELINE config 4: a batch of `c4_frames` frames of this config, stock = frame index % 4, frame ->
    # rank index % world, through BatchExporter (producer thread = phase 1, consumer = pipelined phase 2)
    c4 = None
    if args.c4_frames > 0:
        stocks4 = [SyntheticStock(variant=v) for v in range(4)]
        tasks = [{"src": host_frames[i % n_frames], "negative_film": stocks4[i % 4], "grain_size": GRAIN_SIZE,
                  "grain_sigma": GRAIN_SIGMA, "settings": settings} for i in range(args.c4_frames)]
        exporter = BatchExporter(proc, world_size=world, rank=rank)
        exporter.run(tasks[:4 * world], sink)                       # warm every stock's table slot on every rank
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(exporter._pipe.s_in)
        report = exporter.run(tasks, sink)
        c1.record(exporter._pipe.s_out)
        barrier()
        c4 = {"ms": c0.elapsed_time(c1), "wall_s": report["seconds"], "phase1_s": report["phase1_seconds"],
              "frames_rank": len(report["frames"])}
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        vals = [ms_total, ms_e2e, ms_e2e16, ms_sync, ms_sync_wall, c4["ms"] if c4 else 0.0,
                c4["wall_s"] if c4 else 0.0, c4["phase1_s"] if c4 else 0.0]
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, ms_e2e16, ms_sync, ms_sync_wall = (float(t[i]) for i in range(5))
        if c4:
            c4["ms"], c4["wall_s"], c4["phase1_s"] = float(t[5]), float(t[6]), float(t[7])
        ln = torch.tensor([launches, c4["frames_rank"] if c4 else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln)
        launches = int(ln[0])
        if c4:
            c4["frames_rank"] = int(ln[1])

    if rank == 0:
        peak, peak_src = load_peaks()
        value = world * args.steps * mp / (ms_total / 1e3)
        kernels = {}
        for name, ms, n in zip(_cabi.PROF_NAMES, prof_ms, prof_n):
            if n:
                kernels[name] = {"ms": ms / n, "launches_per_step": n / args.steps,
                                 "gbs": KERNEL_BYTES_PER_PX[name] * H * W / (ms / n * 1e-3) / 1e9}
                kernels[name]["hbm_frac"] = kernels[name]["gbs"] / peak
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
            # direct correlations are FP32 bound.  Two rates: the algorithmic 2*k*k flop per pixel and filtered layer
            # (what the reference's filter2D would execute), and the flop the y-symmetric kernel really executes:
            # (k+1)/2 * k multiply-adds + the row-pair additions (2 * NQ*4 per kernel row and 16-pixel strip) --
            # the executed rate is the one to hold against the measured FFMA peak (profiles/r01_ffma2_rate.txt).
            for name in ("halation", "mtf"):
                k = getattr(proc, name + "_kernel", None) if name in kernels else None
                if k is not None:
                    kk = k.shape[0]
                    layers = sum(int(np.count_nonzero(k[..., c])) > 1 for c in range(3))
                    sec = kernels[name]["ms"] * 1e-3
                    kernels[name]["alg_fp32_tflops"] = 2.0 * layers * kk * kk * H * W / sec / 1e12
                    r = kk // 2
                    nq = (16 + kk - 1 + 3) // 4
                    exec_flop_px = (2.0 * (r + 1) * kk) + (r + 1) * (nq * 4) * 2 / 32.0
                    kernels[name]["exec_fp32_tflops"] = layers * exec_flop_px * H * W / sec / 1e12
            if dom in ("halation", "mtf") and "exec_fp32_tflops" in kernels[dom]:
                roofline["fp32"] = {"executed_tflops": kernels[dom]["exec_fp32_tflops"], "peak_tflops": FFMA_PEAK_TFLOPS,
                                    "frac": kernels[dom]["exec_fp32_tflops"] / FFMA_PEAK_TFLOPS,
                                    "algorithmic_tflops": kernels[dom]["alg_fp32_tflops"],
                                    "peak_source": "tools/micro/ffma2_rate.cu on this pool's B200 (profiles/r01_ffma2_rate.txt)"}
                roofline["note"] = ("direct 2-D correlation is FP32-FMA bound, not HBM bound (SURVEY 8d): `fp32.frac` = "
                                    "executed flop rate / measured FFMA peak is the fraction that describes it; `frac` "
                                    "is its HBM fraction as the contract defines it")
            elif dom == "grain":
                roofline["note"] = ("fused MTF/grain + tetrahedral LUT + quantise: bound by instruction issue and the "
                                    "FP32 pipe, not HBM (profiles/)")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import film_oracle as fo

            fo.use_all_host_threads()
            xyz, st, sample = cpu_sample(args.config, H, W, settings, rows=H // 2 if H > 2000 else H)
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
            "config": config_dict(args.config), "frames_per_s": value / mp,
            "e2e": {"value": world * args.steps * mp / (ms_sync / 1e3), "unit": "MP/s", "h2d_bytes_per_step": h2d_sync,
                    "d2h_bytes_per_step": d2h_sync, "ms_per_step": ms_sync / args.steps,
                    "wall_ms_per_step": ms_sync_wall / args.steps,
                    "api": "B200Processor.process_preloaded (the reference's entry point, gpu_processor.py:1643-1693): one "
                           "synchronous call per frame, pinned host float32 in, host uint8 out; H2D, render and D2H of "
                           "a frame run back to back, nothing overlaps across calls"},
            "e2e_pipelined": {"value": world * args.steps * mp / (ms_e2e / 1e3), "unit": "MP/s",
                              "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": d2h_per_step,
                              "ms_per_step": ms_e2e / args.steps,
                              "api": "PipelinedRenderer.run (what BatchExporter drives): H2D / render / D2H of "
                                     "consecutive frames overlapped, depth 3"},
            "e2e_u16": {"value": world * args.steps * mp / (ms_e2e16 / 1e3), "unit": "MP/s",
                        "h2d_bytes_per_step": h2d16_per_step, "d2h_bytes_per_step": d2h_per_step,
                        "ms_per_step": ms_e2e16 / args.steps,
                        "note": "same pipelined batch from uint16 XYZ frames (what rawpy hands over); /65535 and "
                                "exposure gain applied on the device (SURVEY 8f-1)"},
            "latency": latency, "affinity": affinity, "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
            "cpu_baseline": cpu, "clocks": clocks, "checksum": checksum,
        }
        if c4:
            nf = c4["frames_rank"]
            line["c4"] = {"frames": nf, "frames_per_s": nf / (c4["ms"] / 1e3), "mp_per_s": nf * mp / (c4["ms"] / 1e3),
                          "ms_total": c4["ms"], "wall_s": c4["wall_s"], "phase1_s_max_rank": c4["phase1_s"],
                          "what": "BASELINE config 4: batch of frames of this config through BatchExporter, stock = frame "
                                  "index % 4 (4 table slots per GPU), frame -> rank index % world; source frames live "
                                  "in pinned host memory (B200Processor.pinned_frame), so phase 1 hands them to the "
                                  "copy engine without a host copy; device-timed, max over ranks"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    proc.close()


if __name__ == "__main__":
    main()
