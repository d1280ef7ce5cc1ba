"""Internal consistency of the oracle: the C restatement vs its NumPy twins for the stages that
have no reference golden (third-party: apply_2d_lut, log_clip, multi_channel_interp), the fused
C chain vs the staged functions, and edge cases (empty-ish, ragged, extreme values)."""
import numpy as np
import pytest

from oracle import film_oracle as fo
from raw2film_b200.synthetic import SyntheticStock
from tests.helpers import oracle_render, small_frame


def _edge_frame(rng, h=32, w=48):
    xyz = rng.random((h, w, 3), dtype=np.float32)
    xyz[0] = 0.0
    xyz[1] = 1e-14
    xyz[2] = 1e6
    xyz[3, :, 0] = -0.05
    xyz[4, :, 2] = -0.2
    xyz[5] = np.float32(1 / 3)
    xyz[6, :, 0] = xyz[6, :, 1]
    xyz[7] = np.float32(1e-7)
    return xyz


@pytest.mark.parametrize("n", [2, 5, 16, 64])
def test_apply_2d_lut_c_equals_numpy(n):
    rng = np.random.default_rng(n)
    lut = rng.random((n, n, 3), dtype=np.float32)
    xyz = _edge_frame(rng)
    a, b = fo.apply_2d_lut(xyz, lut), fo.apply_2d_lut_np(xyz, lut)
    assert np.array_equal(a, b)
    assert np.array_equal(a[0], np.zeros_like(a[0]))           # S < 1e-12 -> 0 (lut_2d.wgsl:47)
    xyza = np.concatenate([xyz, np.ones_like(xyz[..., :1])], axis=-1)
    assert np.array_equal(fo.apply_2d_lut(xyza, lut), a)      # XYZ + alpha payload layout


def test_apply_2d_lut_reproduces_lattice_values():
    """On lattice chromaticities the interpolation returns lut[i, j] * S exactly."""
    n = 9
    rng = np.random.default_rng(0)
    lut = rng.random((n, n, 3), dtype=np.float32)
    i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    s = np.float32(8.0)                                         # (n-1)/S = 1: r, g are exact integers
    xyz = np.stack([i, j, 8 - i - j], axis=-1).astype(np.float32)
    out = fo.apply_2d_lut(xyz, lut)
    assert np.array_equal(out, lut[i, j] * s)


def test_log_clip_and_interp_c_equals_numpy():
    rng = np.random.default_rng(3)
    img = np.exp(rng.uniform(-20, 6, (64, 64, 3))).astype(np.float32)
    img[0, :4] = [[0, -1, 1e-6], [1e-7, 1, 10], [np.float32(1e-6), 100, 1e30], [0.5, 2, 4]]
    logged = fo.log_clip(img.copy())
    assert np.array_equal(logged, fo.log_clip_np(img))
    assert logged.min() == np.float32(-6.0)                     # eps = 1e-6 (lut_1d.wgsl:24)
    for N in (2, 7, 1024):
        curve = SyntheticStock(n1=N).get_density_curve()
        a, b = fo.multi_channel_interp(logged, curve), fo.multi_channel_interp_np(logged, curve)
        assert np.array_equal(a, b)
        lo = fo.multi_channel_interp(np.full((1, 1, 3), -50, np.float32), curve)[0, 0]
        hi = fo.multi_channel_interp(np.full((1, 1, 3), 50, np.float32), curve)[0, 0]
        assert np.array_equal(lo, curve[1:, 0]) and np.abs(hi - curve[1:, -1]).max() < 1e-6   # clamped ends


def test_interp_hits_knots_exactly():
    curve = SyntheticStock(n1=65).get_density_curve()
    x = np.linspace(-4, 2, 65).astype(np.float32)
    img = np.repeat(x[:, None, None], 3, axis=2)
    out = fo.multi_channel_interp(img, curve)
    assert np.abs(out[:, 0, :] - curve[1:].T).max() < 2e-6


def test_quantise_truncates():
    v = np.array([0.0, 0.999 / 255, 1.0 / 255, 0.5, 254.999 / 255, 1.0, 1.5, -0.1, np.nan], np.float32)
    assert fo.quantise_u8(v).tolist() == [0, 0, 1, 127, 254, 255, 255, 0, 0]
    ok = v[:6]
    assert np.array_equal(fo.quantise_u8(ok), (ok * (2 ** 8 - 1)).astype(np.uint8))   # cpu_processor.py:407


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (37, 53)])
def test_fused_chain_equals_staged(shape):
    stock = SyntheticStock(n3=9)
    xyz = small_frame(*shape, seed=1)
    st = dict(halation=False, sharpness=False, grain=0)
    staged = oracle_render(fo, xyz, stock, 6.0, 0.4, st)
    fused = fo.pointwise_chain(xyz, stock.get_input_lut(6500, 0.0, 0.0), stock.get_density_curve(), stock.create_lut())
    assert np.array_equal(staged, fused)


def test_render_stage_order_and_gating():
    """Stage taps appear exactly for the enabled stages, in the order of cpu_processor.py:363-407."""
    stock = SyntheticStock(n3=9)
    xyz = small_frame(40, 60, seed=2)
    stages = {}
    oracle_render(fo, xyz, stock, 6.0, 0.4, dict(frame_width=2.0, frame_height=1.5, highlight_burn=0.5,
                                                 burn_scale=8.0), stages=stages)
    assert list(stages) == ["exposure", "halation", "density", "mtf", "grain", "burn", "rgb"]
    assert np.array_equal(stages["halation"][..., 2], stages["exposure"][..., 2]) or \
        np.abs(stages["halation"][..., 2] - stages["exposure"][..., 2]).max() < 1e-4    # blue: delta kernel
    assert stages["grain"].min() >= 0.0 and stages["burn"].min() >= 0.0
    stages = {}
    plain = SyntheticStock(n3=9, with_mtf=False, with_grain=False, density_measure="status_a")
    oracle_render(fo, xyz, plain, 6.0, 0.4, dict(halation=False, highlight_burn=0.5), stages=stages)
    assert list(stages) == ["exposure", "density", "rgb"]


def test_halation_kernel_properties():
    k = fo.compute_halation_kernel(6000 / 36, 1.0, 1.0, 0.3, 0.0, 1.0)
    assert k.shape == (43, 43, 3) and k.dtype == np.float32
    assert np.allclose(k.sum(axis=(0, 1)), 1.0, atol=3e-6)
    blue = k[..., 2]
    assert blue[21, 21] == 1.0 and np.count_nonzero(blue) == 1
    assert np.array_equal(k[..., 0], k[..., 0].T) and np.array_equal(k[..., 0], k[::-1, ::-1, 0])
    bw = fo.compute_halation_kernel(50.0, bw=True)
    assert np.array_equal(bw[..., 0], bw[..., 1]) and np.array_equal(bw[..., 1], bw[..., 2])


def test_grain_field_statistics():
    for scale, expect_kernel in ((6000 / 36, True), (20.0, False)):
        kern = fo.grain_kernel(1 / scale, 0.006, 0.4)
        assert (kern is not None) == expect_kernel
        field = fo.generate_grain((256, 256, 3), scale, 0.006, False, 0.4, seed=1)
        assert field.shape == (256, 256, 3) and abs(field.std() - 1.0) < 0.05 and abs(field.mean()) < 0.02
    bw = fo.generate_grain((64, 64, 3), 166.0, 0.006, True, 0.4, seed=1)
    assert np.array_equal(bw[..., 0], bw[..., 1]) and np.array_equal(bw[..., 1], bw[..., 2])


def test_nonuniform_curve_is_np_interp_and_uniform_detection():
    """a (4, N) table with a non-uniform abscissa is evaluated like np.interp (bit for bit);
    a float32 linspace counts as uniform and keeps the reference GPU path's normalised lookup."""
    from raw2film_b200.synthetic import SyntheticStock

    rng = np.random.default_rng(3)
    warped = SyntheticStock(warped_curve=True)
    curve = warped.get_density_curve()
    assert not fo.abscissa_uniform(curve[0]) and np.all(np.diff(curve[0]) > 0)
    assert fo.abscissa_uniform(SyntheticStock().get_density_curve()[0])
    assert fo.abscissa_uniform(np.linspace(0, 4, 4096).astype(np.float32))
    img = rng.uniform(-6.5, 2.5, (64, 80, 3)).astype(np.float32)
    got = fo.multi_channel_interp(img, curve)
    want = np.stack([np.interp(img[..., k].astype(np.float64), curve[0].astype(np.float64),
                               curve[k + 1].astype(np.float64)).astype(np.float32) for k in range(3)], axis=-1)
    assert np.array_equal(got, want)
    # on a uniform table both rules agree to rounding
    uni = SyntheticStock().get_density_curve()
    a = fo.multi_channel_interp(img, uni)
    b = fo.multi_channel_interp_nonuniform(img, uni)
    assert np.abs(a - b).max() < 2e-6
