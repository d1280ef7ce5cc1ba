"""Stage-level parity of the CUDA path against the oracle through the r2f_render_tap /
r2f_convolve2d entry points.  Tolerances (BASELINE north_star): deterministic stages max-abs
<= 1e-4 in the float working space (density / display RGB); linear exposure, whose range is
unbounded, is held to 1e-5 of the frame maximum."""
import ctypes

import numpy as np
import pytest

from oracle import film_oracle as fo
from raw2film_b200 import settings as S
from raw2film_b200.synthetic import SyntheticStock
from tests.helpers import oracle_render, small_frame

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def proc():
    from raw2film_b200 import B200Processor

    p = B200Processor(device=0)
    yield p
    p.close()


def _tap(proc, xyz, stage, stock, **settings):
    import torch

    x = torch.from_numpy(np.ascontiguousarray(xyz)).cuda()
    return proc.render_tap(x, stage, stock, 6.0, 0.4, **settings).cpu().numpy()


def _gpu_convolve(proc, img, kernel, path="auto"):
    import torch
    from raw2film_b200 import _cabi

    proc.set_conv_path(path)

    h, w = img.shape[:2]
    x = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    out = torch.empty_like(x)
    ws = torch.empty(int(_cabi.lib.r2f_workspace_bytes(h, w, 0)), dtype=torch.uint8, device="cuda")
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib.r2f_convolve2d(proc._ctx, x.data_ptr(), out.data_ptr(), h, w, _cabi.f32_ptr(kernel),
                                         kernel.shape[0], ws.data_ptr(), ws.numel(), None))
    torch.cuda.synchronize()
    proc.set_conv_path("auto")
    return out.cpu().numpy()


@pytest.mark.parametrize("shape,scale,size", [((300, 200), 150.0, 1.0), ((211, 307), 102.0, 1.0), ((97, 130), 40.0, 1.0),
                                              ((402, 603), 6000 / 36, 1.0), ((530, 801), 9504 / 36, 2.0),
                                              # tall strip: column length 4096 -> the compile-time in-place column
                                              # kernel (k_fft_cols_ip) of the 24 MP frame; the 6912 plan of the 61 MP
                                              # frame is covered by the full-size test in test_gpu_full.py
                                              ((4000, 128), 6000 / 36, 1.0),
                                              # wide strip: row length 6144 -> the compile-time row plan, ragged rows
                                              ((130, 6000), 6000 / 36, 1.0)])
def test_halation_fft_path_vs_truth_and_direct(proc, shape, scale, size):
    """The FFT path (packed R+iG, real kernel spectrum) against the float64 direct truth and against
    the direct CUDA kernel, on ragged sizes (odd H -> single-row tail CTA, W % 4 != 0) with real
    halation kernels up to 133x133 and point highlights 200x above the surround."""
    rng = np.random.default_rng(shape[0])
    img = (rng.random((*shape, 3), dtype=np.float32) * 0.5).astype(np.float32)
    for _ in range(12):
        img[rng.integers(0, shape[0]), rng.integers(0, shape[1])] = 100.0
    kern = fo.compute_halation_kernel(scale, size, 1.0, 0.3, 0.0, 1.0)
    truth = fo.correlate_truth_f64(img, kern)
    fft = _gpu_convolve(proc, img, kern, "fft")
    direct = _gpu_convolve(proc, img, kern, "direct")
    # float32 accumulation next to 100.0-valued highlights: bound the error relative to the local value
    for name, got in (("direct", direct), ("fft", fft)):
        rel = np.abs(got - truth) / np.maximum(np.abs(truth), 0.05)
        assert rel.max() <= 1e-4, f"{name}: max rel err {rel.max():.2e}, max abs {np.abs(got - truth).max():.2e}"
    assert np.array_equal(fft[..., 2], img[..., 2])          # blue layer untouched


def test_fft_path_refuses_ineligible_kernel(proc):
    from raw2film_b200 import _cabi

    rng = np.random.default_rng(0)
    img = rng.random((64, 64, 3), dtype=np.float32)
    kern = rng.random((9, 9, 3), dtype=np.float32)            # asymmetric, three filtered layers
    with pytest.raises(_cabi.R2FError):
        _gpu_convolve(proc, img, kern, "fft")
    proc.set_conv_path("auto")


def test_convolve_matches_reference_golden(proc):
    """tests/golden/convolve.npz was produced by the reference's own convolve_2d (effects.py:146-156):
    asymmetric kernels prove correlation orientation, centre anchor and REFLECT_101 borders."""
    g = np.load("tests/golden/convolve.npz")
    for name in ("small", "big"):
        got = _gpu_convolve(proc, g["img"], g["k_" + name])
        assert np.abs(got - g["ref_" + name]).max() <= 2e-6


@pytest.mark.parametrize("shape,k", [((5, 7), 3), ((9, 4), 7), ((16, 16), 43), ((70, 130), 43), ((33, 200), 1),
                                     ((130, 70), 17), ((64, 64), 133), ((200, 150), 133)])
def test_convolve_vs_float64_truth(proc, shape, k):
    """Ragged sizes, images smaller than the kernel (multiple reflections), k = 1 .. 133."""
    rng = np.random.default_rng(k * 100 + shape[0])
    img = rng.random((*shape, 3), dtype=np.float32)
    kern = rng.random((k, k, 3), dtype=np.float32)
    kern /= kern.sum(axis=(0, 1), keepdims=True)
    if min(shape) > k // 2:     # scipy 'mirror' handles one reflection per side like cv2 for these sizes
        truth = fo.correlate_truth_f64(img, kern)
    else:                        # fold indices explicitly (cv2.borderInterpolate REFLECT_101 loop)
        r = k // 2
        def fold(p, n):
            if n == 1:
                return np.zeros_like(p)
            period = 2 * n - 2
            p = np.mod(p, period)
            return np.where(p >= n, period - p, p)
        yy = fold(np.arange(-r, shape[0] + r), shape[0])
        xx = fold(np.arange(-r, shape[1] + r), shape[1])
        pad = img[yy][:, xx].astype(np.float64)
        truth = np.zeros(img.shape, np.float64)
        for i in range(k):
            for j in range(k):
                truth += kern[i, j].astype(np.float64) * pad[i:i + shape[0], j:j + shape[1]]
    got = _gpu_convolve(proc, img, kern)
    assert np.abs(got - truth).max() <= 5e-6


@pytest.mark.parametrize("k", [3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31, 33])
def test_symmetric_kernel_path_vs_truth_and_generic(proc, k):
    """Kernels mirrored in y (but NOT in x: the tap orientation stays checked) take k_conv2d_sym (row-pair
    sums + packed FMA); it must agree with the float64 truth and with the generic kernel on ragged frames,
    frames smaller than a tile and frames smaller than the kernel radius."""
    rng = np.random.default_rng(k)
    for shape in ((70 + k, 131), (64, 64), (5, 9), (130, 67), (256, 320)):   # the last one has interior (TMA) tiles
        img = rng.random((*shape, 3), dtype=np.float32)
        half = rng.random((k // 2 + 1, k, 3), dtype=np.float32)
        kern = np.concatenate([half[:0:-1], half], axis=0)
        kern[:, :, 2] = 0.0
        kern[k // 2, k // 2, 2] = 1.0                               # one identity layer
        kern /= kern.sum(axis=(0, 1), keepdims=True)
        assert np.array_equal(kern, kern[::-1])
        proc.set_conv_sym(True)
        sym = _gpu_convolve(proc, img, kern, "direct")
        proc.set_conv_sym(False)
        gen = _gpu_convolve(proc, img, kern, "direct")
        proc.set_conv_sym(True)
        if min(shape) > k // 2:
            truth = fo.correlate_truth_f64(img, kern)
            assert np.abs(sym - truth).max() <= 5e-6
        assert np.abs(sym - gen).max() <= 5e-6
        assert np.array_equal(sym[..., 2], img[..., 2])


@pytest.mark.parametrize("scale,strength", [(6000 / 36, 0.0), (9504 / 36, 0.5), (1920 / 36, 0.0)])
def test_symmetric_kernel_path_real_mtf(proc, scale, strength):
    from raw2film_b200 import builders

    rng = np.random.default_rng(7)
    img = (rng.random((200, 333, 3), dtype=np.float32) * 3.5).astype(np.float32)
    kern = builders.mtf_kernel(SyntheticStock().mtf, scale, strength, 1.0)
    truth = fo.correlate_truth_f64(img, kern)
    got = _gpu_convolve(proc, img, kern, "direct")
    assert np.abs(got - truth).max() <= 1e-5


def test_identity_channel_is_exact(proc):
    """A centre-delta channel (halation blue layer, effects.py:255-262 with factor 0) passes through bit-exactly."""
    rng = np.random.default_rng(1)
    img = rng.random((50, 60, 3), dtype=np.float32)
    kern = fo.compute_halation_kernel(40.0)
    got = _gpu_convolve(proc, img, kern)
    assert np.array_equal(got[..., 2], img[..., 2])
    assert not np.array_equal(got[..., 0], img[..., 0])


@pytest.mark.parametrize("shape,frame_w", [((96, 144), 36), ((211, 307), 36), ((300, 200), 24)])
def test_taps_match_oracle_stages(proc, shape, frame_w):
    stock = SyntheticStock(n3=17)
    xyz = small_frame(*shape, seed=7)
    # scale chosen through frame_width so that halation / MTF kernels have realistic sizes
    st = dict(frame_width=frame_w / 12.0, frame_height=frame_w / 18.0, grain=2, highlight_burn=0.0)
    noise = fo.white_noise(xyz.shape, False, seed=99)
    stages = {}
    want_u8 = oracle_render(fo, xyz, stock, 6.0, 0.4, st, noise=noise, stages=stages)
    st_gpu = dict(st, grain_noise=noise)

    exp = _tap(proc, xyz, "exposure", stock, **st_gpu)
    assert np.array_equal(exp, stages["exposure"]), "2-D LUT stage must be bit-exact"

    hal = _tap(proc, xyz, "halation", stock, **st_gpu)
    lim = 1e-5 * max(1.0, float(np.abs(stages["halation"]).max()))
    assert np.abs(hal - stages["halation"]).max() <= lim
    assert np.array_equal(hal[..., 2], exp[..., 2]), "blue layer is not halated"

    for name in ("density", "mtf", "grain", "rgb"):
        got = _tap(proc, xyz, name, stock, **st_gpu)
        err = np.abs(got - stages[name]).max()
        assert err <= TOL, f"{name}: max abs err {err}"

    import torch
    out = proc.render_device(torch.from_numpy(xyz).cuda(), stock, 6.0, 0.4, **st_gpu)
    proc.stream.synchronize()
    got_u8 = out.cpu().numpy()
    diff = np.abs(got_u8.astype(np.int16) - want_u8.astype(np.int16))
    assert diff.max() <= 1, "more than 1 LSB off at 8-bit output"
    assert np.mean(diff != 0) < 2e-3, f"LSB flip rate {np.mean(diff != 0):.2e}"


def test_density_tap_without_spatial_stages_is_bit_exact(proc):
    stock = SyntheticStock()
    xyz = small_frame(64, 80, seed=2)
    st = dict(halation=False, sharpness=False, grain=0)
    stages = {}
    oracle_render(fo, xyz, stock, 6.0, 0.4, st, stages=stages)
    assert np.array_equal(_tap(proc, xyz, "density", stock, **st), stages["density"])
    assert np.array_equal(_tap(proc, xyz, "rgb", stock, **st), stages["rgb"])


@pytest.mark.parametrize("shape,burn,scale", [((72, 108), 0.5, 10.0), ((57, 83), 0.8, 7.0), ((160, 240), 1.0, 50.0)])
def test_burn_tap(proc, shape, burn, scale):
    """Highlight burn (effects.py:392-418) on the density image, halation/MTF/grain off."""
    stock = SyntheticStock()
    xyz = small_frame(*shape, seed=13)
    st = dict(halation=False, sharpness=False, grain=0, highlight_burn=burn, burn_scale=scale)
    stages = {}
    oracle_render(fo, xyz, stock, 6.0, 0.4, st, stages=stages)
    got = _tap(proc, xyz, "burn", stock, **st)
    assert np.abs(got - stages["burn"]).max() <= TOL
