"""Whole-path parity through the reference-facing API (process / process_preloaded) and
size-independent properties at the BASELINE sizes."""
import numpy as np
import pytest

from oracle import film_oracle as fo
from raw2film_b200.synthetic import SyntheticStock, natural_frame
from tests.helpers import oracle_render, small_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proc():
    from raw2film_b200 import B200Processor

    p = B200Processor(device=0)
    yield p
    p.close()


def _lsb_report(got, want):
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    return int(diff.max()), float(np.mean(diff != 0))


@pytest.mark.parametrize("grain_mode", [2, 1])
def test_process_full_emulation_injected_noise(proc, grain_mode):
    """Full emulation with the oracle's own noise field injected: <= 1 LSB at 8 bit."""
    stock = SyntheticStock(n3=17)
    xyz = small_frame(240, 360, seed=21)
    noise = fo.white_noise(xyz.shape, grain_mode == 1, seed=5)
    st = dict(frame_width=3.0, frame_height=2.0, grain=grain_mode, halation_green_factor=0.3)
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st, noise=noise)
    got = proc.process(xyz, stock, 6.0, 0.4, grain_noise=noise, **st)
    assert got.dtype == np.uint8 and got.shape == want.shape
    mx, rate = _lsb_report(got, want)
    assert mx <= 1 and rate < 2e-3, (mx, rate)


@pytest.mark.parametrize("shape", [(241, 363), (250, 361), (239, 358), (130, 1027)])
def test_full_emulation_odd_widths_injected_noise(proc, shape):
    """Widths that are not multiples of 4: the density planes between the row-inverse FFT kernel, the MTF kernel and
    the grain kernel use a padded row pitch (TMA / 16-byte paths); the bytes must still be the oracle's (<= 1 LSB),
    and the same as through the generic kernels, whose planes are unpadded."""
    stock = SyntheticStock(n3=17)
    xyz = small_frame(*shape, seed=shape[1])
    noise = fo.white_noise(xyz.shape, False, seed=9)
    st = dict(frame_width=3.0, frame_height=3.0 * shape[0] / shape[1], grain=2, halation_green_factor=0.3)
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st, noise=noise)
    got = proc.process(xyz, stock, 6.0, 0.4, grain_noise=noise, **st)
    mx, rate = _lsb_report(got, want)
    assert mx <= 1 and rate < 2e-3, (mx, rate)
    proc.set_conv_sym(False)
    try:
        generic = proc.process(xyz, stock, 6.0, 0.4, grain_noise=noise, **st)
    finally:
        proc.set_conv_sym(True)
    mx, rate = _lsb_report(got, generic)
    assert mx <= 1 and rate < 2e-3, (mx, rate)
    # regenerated noise (no injection): the same field through both kernel families
    a = proc.process(xyz, stock, 6.0, 0.4, grain_seed=77, **st)
    proc.set_conv_sym(False)
    try:
        b = proc.process(xyz, stock, 6.0, 0.4, grain_seed=77, **st)
    finally:
        proc.set_conv_sym(True)
    mx, rate = _lsb_report(a, b)
    assert mx <= 1 and rate < 2e-3, (mx, rate)


def test_process_pointwise_config_bit_exact(proc):
    """Config C1 through the public API: stages off -> bit-exact uint8."""
    stock = SyntheticStock()
    xyz = small_frame(200, 300, seed=4)
    st = dict(halation=False, sharpness=False, grain=0)
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st)
    got = proc.process(xyz, stock, 6.0, 0.4, **st)
    assert np.array_equal(got, want)


def test_process_preloaded_matches_process_and_accepts_reference_payload(proc):
    stock = SyntheticStock()
    xyz = small_frame(128, 192, seed=8)
    st = dict(halation=True, sharpness=True, grain=0, frame_width=2.0, frame_height=1.5)
    a = proc.process(xyz, stock, 6.0, 0.4, **st)
    payload = proc.extract_image_data_cpu(xyz, **st)
    assert payload["pipeline_resolution"] == (192, 128) and payload["output_resolution"] == (192, 128)
    b = proc.process_preloaded(payload, stock, 6.0, 0.4, **st)
    # reference-style foreign payload: XYZ + ones alpha, plain (unpinned) numpy
    foreign = {"image_array": np.dstack((xyz, np.ones_like(xyz[..., :1]))), "output_resolution": (192, 128),
               "canvas_resolution": None, "pipeline_resolution": (192, 128)}
    c = proc.process_preloaded(foreign, stock, 6.0, 0.4, **st)
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_stage_gating_follows_stock(proc):
    """No MTF data / no rms_density -> the stages are skipped (cpu_processor.py:382, 387)."""
    plain = SyntheticStock(with_mtf=False, with_grain=False)
    xyz = small_frame(90, 120, seed=3)
    st = dict(halation=False, frame_width=2.0, frame_height=1.5)
    want = oracle_render(fo, xyz, plain, 6.0, 0.4, st)
    got = proc.process(xyz, plain, 6.0, 0.4, **st)
    assert np.array_equal(got, want)      # effectively pointwise


def test_grain_statistics_device_noise(proc):
    """Statistical equivalence of the on-device Philox grain: white N(0,1) field (mean, variance,
    kurtosis, flat spectrum), independent channels, seed-determinism."""
    import torch
    from raw2film_b200 import _cabi

    h, w = 512, 768
    out = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
    _cabi.check(_cabi.lib.r2f_generate_noise(proc._ctx, out.data_ptr(), h, w, 3, 1234, None))
    torch.cuda.synchronize()
    n = out.cpu().numpy().astype(np.float64)
    assert abs(n.mean()) < 5e-3 and abs(n.var() - 1.0) < 1e-2
    assert abs(((n - n.mean()) ** 4).mean() / n.var() ** 2 - 3.0) < 0.05
    for a in range(3):
        for b in range(a + 1, 3):
            assert abs(np.corrcoef(n[..., a].ravel(), n[..., b].ravel())[0, 1]) < 5e-3
    spec = np.abs(np.fft.fft2(n[..., 0])) ** 2 / (h * w)
    bands = [spec[:h // 8, :w // 8].mean(), spec[h // 4:h // 2, w // 4:w // 2].mean()]
    assert abs(bands[0] - 1.0) < 0.05 and abs(bands[1] - 1.0) < 0.05
    out2 = torch.empty_like(out)
    _cabi.check(_cabi.lib.r2f_generate_noise(proc._ctx, out2.data_ptr(), h, w, 3, 1234, None))
    out3 = torch.empty_like(out)
    _cabi.check(_cabi.lib.r2f_generate_noise(proc._ctx, out3.data_ptr(), h, w, 3, 1235, None))
    torch.cuda.synchronize()
    assert torch.equal(out, out2) and not torch.equal(out, out3)


def test_grain_device_noise_matches_oracle_statistics(proc):
    """Per-density-bin mean/variance of the added grain (device noise vs oracle noise)."""
    import torch

    stock = SyntheticStock(n3=17)
    xyz = small_frame(384, 512, seed=31, highlights=False)
    st = dict(halation=False, sharpness=False, grain=2, frame_width=4.0, frame_height=3.0, grain_seed=77)
    x = torch.from_numpy(xyz).cuda()
    dens = proc.render_tap(x, "density", stock, 6.0, 0.4, **st).cpu().numpy()
    grained = proc.render_tap(x, "grain", stock, 6.0, 0.4, **st).cpu().numpy()
    stages = {}
    oracle_render(fo, xyz, stock, 6.0, 0.4, st, noise=None, stages=stages)
    g_gpu, g_orc = grained - dens, stages["grain"] - stages["density"]
    bins = np.quantile(dens[..., 1], np.linspace(0, 1, 9))
    for lo, hi in zip(bins[:-1], bins[1:]):
        m = (dens[..., 1] >= lo) & (dens[..., 1] < hi) & (grained[..., 1] > 0)
        if m.sum() < 2000:
            continue
        a, b = g_gpu[..., 1][m], g_orc[..., 1][m]
        assert abs(a.mean() - b.mean()) < 4 * b.std() / np.sqrt(m.sum()) + 1e-4
        assert abs(a.std() / b.std() - 1.0) < 0.05


def test_full_size_properties_24mp(proc):
    """BASELINE config C2 at 6000x4000 through size-independent properties:
    (1) the render is deterministic for a fixed seed; (2) with halation/MTF/grain on, a frame whose
    exposure is uniform renders to a uniform image away from nothing (kernels sum to 1);
    (3) pointwise-only rows of the same frame agree bit-exactly with the oracle on a row sample."""
    import torch

    stock = SyntheticStock()
    h, w = 4000, 6000
    xyz = natural_frame(h, w, 0)
    x = torch.from_numpy(xyz).cuda()
    st = dict(grain=2, grain_seed=5, halation_green_factor=0.3)
    a = proc.render_device(x, stock, 6.0, 0.4, **st).clone()
    b = proc.render_device(x, stock, 6.0, 0.4, **st).clone()
    proc.stream.synchronize()
    assert torch.equal(a, b)
    flat = torch.full((h, w, 3), 0.2, dtype=torch.float32, device="cuda")
    st2 = dict(grain=0)
    c = proc.render_device(flat, stock, 6.0, 0.4, **st2).cpu().numpy()
    ref = oracle_render(fo, np.full((8, 8, 3), 0.2, np.float32), stock, 6.0, 0.4,
                        dict(halation=False, sharpness=False, grain=0))
    assert np.abs(c.astype(np.int16) - ref[0, 0].astype(np.int16)).max() <= 1
    assert (c == c[0, 0]).all(axis=-1).mean() > 0.999
    rows = slice(1234, 1250)
    off = dict(halation=False, sharpness=False, grain=0)
    got = proc.render_device(x[rows].contiguous(), stock, 6.0, 0.4, **off).cpu().numpy()
    assert np.array_equal(got, oracle_render(fo, xyz[rows], stock, 6.0, 0.4, off))


@pytest.mark.parametrize("which", ["halation", "mtf"])
def test_full_size_24mp_spatial_stages_are_linear_and_shift_consistent(proc, which):
    """The two spatial stages at the BASELINE size (6000x4000) through properties that do not need a CPU truth:
    the correlation is linear -- K (*) (a x + b y) == a K (*) x + b K (*) y -- and translation-consistent away
    from the borders, for the halation kernel (43x43, FFT path: compile-time row/column plans, in-place column
    kernel) and for the MTF kernel (17x17x3, y-symmetric FFMA2 kernel with TMA-staged tiles)."""
    import torch
    from raw2film_b200 import _cabi, builders

    h, w = 4000, 6000
    scale = w / 36
    if which == "halation":
        kern = fo.compute_halation_kernel(scale, 1.0, 1.0, 0.3, 0.0, 1.0)
        path = "fft"
    else:
        kern = builders.mtf_kernel(SyntheticStock().mtf, scale)
        path = "direct"
    kern = np.ascontiguousarray(kern, np.float32)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand((h, w, 3), device="cuda", generator=g)
    y = torch.rand((h, w, 3), device="cuda", generator=g)
    ws = torch.empty(int(_cabi.lib.r2f_workspace_bytes(h, w, 0)), dtype=torch.uint8, device="cuda")

    def conv(img):
        out = torch.empty_like(img)
        torch.cuda.synchronize()
        _cabi.check(_cabi.lib.r2f_convolve2d(proc._ctx, img.data_ptr(), out.data_ptr(), h, w, _cabi.f32_ptr(kern),
                                             kern.shape[0], ws.data_ptr(), ws.numel(), None))
        torch.cuda.synchronize()
        return out

    proc.set_conv_path(path)
    try:
        cx, cy = conv(x), conv(y)
        mix = conv((0.75 * x + 0.5 * y).contiguous())
        assert (mix - (0.75 * cx + 0.5 * cy)).abs().max().item() <= 2e-6
        # shift by (64, 128): interior outputs must follow the input
        r = kern.shape[0] // 2
        xs = torch.roll(x, shifts=(64, 128), dims=(0, 1)).contiguous()
        cs = conv(xs)
        a = cs[64 + r:h - r, 128 + r:w - r]
        b = cx[r:h - r - 64, r:w - r - 128]
        assert (a - b).abs().max().item() <= 2e-6
        # the kernel layers sum to 1: a flat field stays flat (borders included: REFLECT_101)
        flat = torch.full((h, w, 3), 0.37, device="cuda")
        assert (conv(flat) - 0.37).abs().max().item() <= 2e-6
    finally:
        proc.set_conv_path("auto")


def test_errors_are_loud(proc):
    import torch
    from raw2film_b200 import _cabi

    stock = SyntheticStock()
    with pytest.raises(NotImplementedError):
        proc.process("some_file.ARW", stock, 6.0, 0.4)
    x = torch.zeros((8, 8, 3), device="cuda")
    with pytest.raises(_cabi.R2FError):
        _cabi.check(_cabi.lib.r2f_render(proc._ctx, x.data_ptr(), 8, 8, 5, x.data_ptr(), 0, None, 0, None, 0, None))


@pytest.mark.parametrize("depth", [2, 3])
def test_pipelined_renderer_matches_synchronous_calls(proc, depth):
    """Overlapped H2D / render / D2H (PipelinedRenderer) returns, in order, exactly what
    process_preloaded returns frame by frame (fixed grain seed), for more frames than slots."""
    from raw2film_b200 import PipelinedRenderer

    stock = SyntheticStock(n3=17)
    st = dict(frame_width=2.0, frame_height=1.5, grain=2, grain_seed=11)
    frames = [small_frame(120, 180, seed=100 + i) for i in range(7)]
    want = [proc.process_preloaded(proc.extract_image_data_cpu(f, **st), stock, 6.0, 0.4, **st).copy() for f in frames]
    pipe = PipelinedRenderer(proc, depth=depth)
    got = {}
    n = pipe.run((proc.extract_image_data_cpu(f, **st) for f in frames), stock, 6.0, 0.4,
                 sink=lambda i, img: got.__setitem__(i, img.copy()), **st)
    assert n == len(frames) and sorted(got) == list(range(len(frames)))
    for i in range(len(frames)):
        assert np.array_equal(got[i], want[i]), f"frame {i}"
    assert pipe.h2d_bytes == len(frames) * 120 * 180 * 3 * 4 and pipe.d2h_bytes == len(frames) * 120 * 180 * 3
    with pytest.raises(ValueError):
        pipe.result(0)


@pytest.mark.parametrize("mode", ["Proportional white", "Uniform black", "Fixed white"])
def test_canvas_modes_match_oracle(proc, mode):
    """add_canvas (effects.py:338-357) pasted on the device, then the reference's post-step
    resolution_scaling(image, resolution) (cpu_processor.py:409-412) on the host."""
    from raw2film_b200 import hostops

    stock = SyntheticStock(n3=17)
    xyz = small_frame(90, 140, seed=17)
    st = dict(halation=False, sharpness=False, grain=0, canvas_mode=mode, canvas_scale=1.2, canvas_ratio=0.8)
    core = oracle_render(fo, xyz, stock, 6.0, 0.4, {k: v for k, v in st.items() if not k.startswith("canvas")})
    want = hostops.resolution_scaling(fo.add_canvas(core, mode, 1.2, 0.8), (90, 140))
    got = proc.process(xyz, stock, 6.0, 0.4, **st)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_preview_resolution_and_max_scale(proc):
    """`resolution` (preview) and `max_scale` (dense frames) resize on the host before / after the
    path exactly like cpu_processor.py:122-134 and :411-412."""
    from raw2film_b200 import hostops

    stock = SyntheticStock(n3=17)
    xyz = small_frame(240, 360, seed=23)
    st = dict(halation=False, sharpness=False, grain=0)
    # preview: fit into 100 x 100
    small = hostops.resolution_scaling(xyz, (100, 100))
    want = hostops.resolution_scaling(oracle_render(fo, small, stock, 6.0, 0.4, st), (100, 100))
    got = proc.process(xyz, stock, 6.0, 0.4, resolution=(100, 100), **st)
    assert got.shape == want.shape and np.array_equal(got, want)
    # max_scale: 360 px over a 0.5 mm frame = 720 px/mm > 400 -> render at 400 px/mm, Lanczos back up
    st2 = dict(st, frame_width=0.5, frame_height=0.3333, max_scale=400.0)
    sf = 400.0 / (360 / 0.5)
    res = [round(240 * sf), round(360 * sf)]
    small = hostops.resolution_scaling(xyz, res)
    want = hostops.resolution_scaling(oracle_render(fo, small, stock, 6.0, 0.4, st2), (240, 360))
    got = proc.process(xyz, stock, 6.0, 0.4, **st2)
    assert got.shape == want.shape and got.shape[1] == 360 and np.array_equal(got, want)


@pytest.mark.parametrize("size", [1, 3, 8])
def test_chroma_nr_bit_exact_vs_reference_golden(proc, size):
    """tests/golden/chroma_nr.npz holds outputs of the reference's own chroma_nr_filter (effects.py:547-561)."""
    g = np.load("tests/golden/chroma_nr.npz")
    got = proc.chroma_nr_filter(g["xyz"], size)
    assert got.dtype == np.float32 and np.array_equal(got, g[f"ref_out{size}"])


@pytest.mark.parametrize("shape,size", [((90, 1100), 20), ((70, 513), 2), ((33, 65), 10), ((1, 1), 3), ((3, 700), 7)])
def test_chroma_nr_tile_seams_vs_oracle(proc, shape, size):
    """Frames that cross the 512-pixel row segments and the 64x32 column tiles of k_cnr_rows / k_cnr_cols, up to
    the largest filter the reference GUI can ask for (slider 10, doubled on full-size export: effects.py:547-561),
    with zero and near-zero denominators sprinkled in."""
    rng = np.random.default_rng(size)
    xyz = rng.random(shape + (3,), dtype=np.float32)
    xyz[rng.random(shape) < 0.02] = 0.0
    xyz[rng.random(shape) < 0.02] *= np.float32(1e-9)
    assert np.array_equal(proc.chroma_nr_filter(xyz, size), fo.chroma_nr_filter(xyz, size))


def test_chroma_nr_rejects_oversized_filter(proc):
    with pytest.raises(Exception, match="tap"):
        proc.chroma_nr_filter(np.ones((8, 8, 3), np.float32), 30)


def test_process_with_chroma_nr_matches_oracle(proc):
    stock = SyntheticStock(n3=17)
    xyz = small_frame(100, 150, seed=41)
    st = dict(halation=False, sharpness=False, grain=0, chroma_nr=4)
    want = oracle_render(fo, fo.chroma_nr_filter(xyz, 4), stock, 6.0, 0.4, st)
    got = proc.process(xyz, stock, 6.0, 0.4, **st)
    assert np.array_equal(got, want)


def test_histogram_matches_reference_golden(proc):
    """generate_histogram (utils.py:145-223): device counts + host raster vs the reference's output."""
    import torch

    g = np.load("tests/golden/histogram.npz")
    img = torch.from_numpy(g["img"]).cuda()
    counts = proc.histogram_counts(img)
    want = np.stack([np.bincount(g["img"][..., c].ravel(), minlength=256) for c in range(3)])
    assert np.array_equal(counts, want)
    for h in (100, 64):
        assert np.array_equal(proc.generate_histogram(g["mix"], h, img), g[f"ref_hist{h}"])        # passes 1-3 on device
        assert np.array_equal(proc.generate_histogram(g["mix"], h, img, on_device=False), g[f"ref_hist{h}"])
    # ragged pixel count (npix % 4 != 0) and a big frame
    rng = np.random.default_rng(0)
    for shape in ((7, 9), (1001, 777)):
        a = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
        got = proc.histogram_counts(torch.from_numpy(a).cuda())
        assert np.array_equal(got, np.stack([np.bincount(a[..., c].ravel(), minlength=256) for c in range(3)]))


def test_c_abi_host_entry_point_matches_device_path(proc):
    """r2f_render_host (the host-buffer C-ABI call a non-torch binding of process_preloaded would make):
    plain pageable numpy buffers in and out, staging owned by the context."""
    import ctypes

    from raw2film_b200 import _cabi

    stock = SyntheticStock(n3=17)
    xyz = small_frame(150, 222, seed=55)
    noise = fo.white_noise(xyz.shape, False, seed=2)
    st = dict(frame_width=2.5, frame_height=1.7, grain=2, grain_noise=noise)
    want = proc.process(xyz, stock, 6.0, 0.4, **st)          # also loads every table into the context
    out = np.empty(xyz.shape, np.uint8)
    flags = _cabi.HALATION | _cabi.MTF | _cabi.GRAIN
    _cabi.check(_cabi.lib.r2f_render_host(proc._ctx, xyz.ctypes.data_as(ctypes.c_void_p), 150, 222, 3,
                                          out.ctypes.data_as(ctypes.c_void_p), flags,
                                          noise.ctypes.data_as(ctypes.c_void_p), 3))
    assert np.array_equal(out, want)
    # pointwise flags = 0 through the same entry point
    off = dict(halation=False, sharpness=False, grain=0)
    want0 = proc.process(xyz, stock, 6.0, 0.4, **off)
    _cabi.check(_cabi.lib.r2f_render_host(proc._ctx, xyz.ctypes.data_as(ctypes.c_void_p), 150, 222, 3,
                                          out.ctypes.data_as(ctypes.c_void_p), 0, None, 0))
    assert np.array_equal(out, want0)


def test_black_and_white_stock_halation(proc):
    """density_measure == "bw": the halation kernel uses the green factor on all three layers
    (effects.py:248-250), so no layer is a pass-through and the direct correlation runs on all three."""
    stock = SyntheticStock(n3=17, density_measure="bw")
    xyz = small_frame(130, 170, seed=9)
    st = dict(frame_width=2.0, frame_height=1.5, grain=0, sharpness=False)
    stages = {}
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st, stages=stages)
    got = proc.process(xyz, stock, 6.0, 0.4, **st)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert diff.max() <= 1 and np.mean(diff != 0) < 2e-3
    assert not np.array_equal(stages["halation"][..., 2], stages["exposure"][..., 2])


def test_fused_grain_kernel_matches_staged_kernels(proc):
    """The fused grain+finish kernel (normal render) and the staged k_noise / k_conv2d / k_finish kernels
    (tap path) implement the same stream and arithmetic: with device-generated Philox noise and a fixed
    seed, the uint8 render equals quantise(tetra(grain tap)) within 1 LSB, and the RGB tap agrees."""
    import torch

    stock = SyntheticStock(n3=17)
    # both fused kernels (y-symmetric packed-FMA k_grain_finish_sym and the generic k_grain_finish), RGB and
    # B/W grain, grain kernels from 5x5 to 15x15, a ragged frame (W % 4 != 0, partial tiles)
    for shape, grain_size, sym in (((200, 264), 6.0, True), ((200, 264), 6.0, False), ((131, 203), 3.0, True),
                                   ((131, 203), 14.0, True), ((131, 203), 14.0, False)):
        xyz = small_frame(*shape, seed=77)
        proc.set_conv_sym(sym)
        for grain_mode in (2, 1):
            st = dict(frame_width=1.2, frame_height=0.8, grain=grain_mode, grain_seed=1234)
            x = torch.from_numpy(xyz).cuda()
            rgb = proc.render_tap(x, "rgb", stock, grain_size, 0.4, **st).cpu().numpy()
            want = fo.quantise_u8(rgb)
            got = proc.render_device(x, stock, grain_size, 0.4, **st)
            proc.stream.synchronize()
            got = got.cpu().numpy()
            diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
            assert diff.max() <= 1 and np.mean(diff != 0) < 1e-3, (shape, grain_size, sym, grain_mode, diff.max(),
                                                                  np.mean(diff != 0))
    proc.set_conv_sym(True)


def test_process_skips_ingest_when_image_parameters_are_unchanged(proc):
    """cpu_processor.py:88-105: same image parameters -> the frame on the device is reused; a changed
    effect setting still re-renders, a changed image parameter (or cache=False) re-ingests."""
    stock = SyntheticStock(n3=17)
    xyz = small_frame(64, 96, seed=61)
    calls = {"n": 0}
    real = proc.extract_image_data_cpu

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    proc.extract_image_data_cpu = counting
    try:
        st = dict(halation=False, sharpness=False, grain=0)
        a = proc.process(xyz, stock, 6.0, 0.4, **st)
        b = proc.process(xyz, stock, 6.0, 0.4, **st)
        c = proc.process(xyz, stock, 6.0, 0.4, exp_comp=1.0, **st)          # effect setting only
        assert calls["n"] == 1 and np.array_equal(a, b) and not np.array_equal(a, c)
        assert np.array_equal(c, oracle_render(fo, xyz, stock, 6.0, 0.4, dict(st, exp_comp=1.0)))
        proc.process(xyz, stock, 6.0, 0.4, canvas_mode="Uniform white", canvas_scale=1.1, **st)   # image parameter
        assert calls["n"] == 2
        proc.process(xyz, stock, 6.0, 0.4, cache=False, **st)
        assert calls["n"] == 3
        other = small_frame(64, 96, seed=62)
        d = proc.process(other, stock, 6.0, 0.4, **st)
        assert calls["n"] == 4 and np.array_equal(d, oracle_render(fo, other, stock, 6.0, 0.4, st))
    finally:
        proc.extract_image_data_cpu = real


@pytest.mark.parametrize("shape", [(9, 13), (4, 40), (33, 5)])
def test_full_emulation_on_tiny_frames(proc, shape):
    """Frames smaller than the halation / MTF kernels (multiple REFLECT_101 folds at every border)."""
    stock = SyntheticStock(n3=9)
    xyz = small_frame(*shape, seed=shape[0], highlights=False)
    noise = fo.white_noise((*shape, 3), False, seed=1)
    st = dict(frame_width=0.3, frame_height=0.2, grain=2, grain_noise=noise)     # 43..133 px/mm -> kernels > frame
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, {k: v for k, v in st.items() if k != "grain_noise"}, noise=noise)
    got = proc.process(xyz, stock, 6.0, 0.4, **st)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert got.shape == want.shape and diff.max() <= 1


def test_full_size_61mp_fft_halation_agrees_with_direct_correlation(proc):
    """BASELINE config C3 at full size (9504x6336, halation_size=2 -> 133x133 kernel): the FFT path
    (padded 10240 x 6912 transforms) against the direct 17 689-tap correlation on the same frame, through
    the halation tap; plus determinism of the whole render."""
    import torch

    stock = SyntheticStock()
    h, w = 6336, 9504
    xyz = natural_frame(h, w, 3)
    x = torch.from_numpy(xyz).cuda()
    st = dict(halation_size=2.0, halation_green_factor=0.3, grain=2, grain_seed=9)
    proc.set_conv_path("fft")
    a = proc.render_tap(x, "halation", stock, 6.0, 0.4, **st)
    proc.set_conv_path("direct")
    b = proc.render_tap(x, "halation", stock, 6.0, 0.4, **st)
    proc.set_conv_path("auto")
    assert proc.halation_kernel.shape[0] == 133
    rel = (a - b).abs() / b.abs().clamp_min(1e-3)
    assert float(rel.max()) <= 2e-4, float(rel.max())
    assert torch.equal(a[..., 2], b[..., 2])
    del a, b, rel
    o1 = proc.render_device(x, stock, 6.0, 0.4, **st).clone()
    o2 = proc.render_device(x, stock, 6.0, 0.4, **st).clone()
    proc.stream.synchronize()
    assert torch.equal(o1, o2)


def test_full_size_24mp_full_emulation_against_oracle(proc):
    """BASELINE config C2 at full size against the oracle itself (cv2.filter2D halation 43x43 and MTF
    17x17x3, injected noise field): <= 1 LSB at 8 bit, LSB flip rate reported and bounded, and the
    density working image within 1e-4 on a row band."""
    import torch

    stock = SyntheticStock()
    h, w = 4000, 6000
    xyz = natural_frame(h, w, 1)
    noise = fo.white_noise((h, w, 3), False, seed=4)
    st = dict(grain=2, halation_green_factor=0.3)
    fo.use_all_host_threads()
    stages = {}
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st, noise=noise, stages=stages)
    got = proc.process(xyz, stock, 6.0, 0.4, grain_noise=noise, **st)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    rate = float(np.mean(diff != 0))
    print(f"24 MP full emulation: max |diff| = {diff.max()} LSB, flip rate = {rate:.2e}")
    assert diff.max() <= 1 and rate < 2e-3
    x = torch.from_numpy(xyz).cuda()
    dens = proc.render_tap(x, "density", stock, 6.0, 0.4, **st).cpu().numpy()
    err = np.abs(dens - stages["density"])
    print(f"24 MP density after FFT halation vs cv2 oracle: max abs err = {err.max():.2e}")
    assert err.max() <= 1e-4
    mtf = proc.render_tap(x, "mtf", stock, 6.0, 0.4, **st).cpu().numpy()
    assert np.abs(mtf - stages["mtf"]).max() <= 1e-4
