"""K1 (fused pointwise chain) parity: bit-exact uint8 against the oracle (BASELINE configs C1/C5)."""
import numpy as np
import pytest

from oracle import film_oracle as fo
from raw2film_b200.synthetic import SyntheticStock, adversarial_frame, natural_frame
from tests.helpers import oracle_luts, small_frame
from raw2film_b200 import settings as S

pytestmark = pytest.mark.gpu

OFF = dict(halation=False, sharpness=False, grain=0, highlight_burn=0.0)


@pytest.fixture(scope="module")
def proc():
    from raw2film_b200 import B200Processor

    p = B200Processor(device=0)
    yield p
    p.close()


def _gpu_u8(proc, xyz, stock, **settings):
    import torch

    x = torch.from_numpy(np.ascontiguousarray(xyz)).cuda()
    out = proc.render_device(x, stock, 6.0, 0.4, **settings)
    proc.stream.synchronize()
    return out.cpu().numpy().copy()


def _oracle_u8(xyz, stock, **settings):
    s = S.merged(settings)
    luts = oracle_luts(stock, s, *xyz.shape[:2])
    return fo.pointwise_chain(xyz, luts["lut2d"], luts["curve"], luts["lut3d"])


@pytest.mark.parametrize("shape", [(1, 1), (1, 3), (3, 5), (17, 31), (64, 64), (255, 257), (1080, 1920)])
def test_pointwise_bit_exact_small(proc, shape):
    stock = SyntheticStock()
    xyz = small_frame(*shape, seed=shape[0] * 1000 + shape[1])
    got, want = _gpu_u8(proc, xyz, stock, **OFF), _oracle_u8(xyz, stock, **OFF)
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} mismatching bytes"


def test_pointwise_four_channel_payload(proc):
    """XYZ + alpha layout of the reference GPU payload (gpu_processor.py:765)."""
    stock = SyntheticStock()
    xyz = small_frame(97, 131, seed=5)
    xyza = np.concatenate([xyz, np.ones_like(xyz[..., :1])], axis=-1)
    assert np.array_equal(_gpu_u8(proc, xyza, stock, **OFF), _oracle_u8(xyz, stock, **OFF))


def test_pointwise_edge_values(proc):
    """zeros (S < 1e-12), negatives, huge values, values on LUT lattice lines."""
    stock = SyntheticStock(n2=16, n1=64, n3=9)
    rng = np.random.default_rng(11)
    xyz = rng.random((40, 64, 3), dtype=np.float32)
    xyz[0] = 0.0
    xyz[1] = 1e-14
    xyz[2] = 1e6
    xyz[3, :, 0] = -0.05
    xyz[4, :, 2] = -0.2
    xyz[5] = np.float32(1.0 / 3.0)                 # chromaticity exactly on a lattice diagonal
    xyz[6, :, 0] = xyz[6, :, 1]                    # fs == 1 boundary candidates
    xyz[7] = np.float32(1e-7)                      # below the log clip
    assert np.array_equal(_gpu_u8(proc, xyz, stock, **OFF), _oracle_u8(xyz, stock, **OFF))


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_pointwise_mixed_stocks_and_settings(proc, variant):
    stock = SyntheticStock(variant=variant, n3=17 + 8 * (variant % 2))
    xyz = small_frame(120, 200, seed=variant)
    st = dict(OFF, exp_comp=0.5 * variant - 0.5, exp_kelvin=5000 + 700 * variant, tint=variant - 1.0,
              push_pull=0.5 * (variant - 1), sat_adjust=1.0 + 0.1 * variant)
    assert np.array_equal(_gpu_u8(proc, xyz, stock, **st), _oracle_u8(xyz, stock, **st))


def test_pointwise_large_tables_global_path(proc):
    """2-D LUT too large for shared memory -> global-memory table path of K1."""
    stock = SyntheticStock(n2=128, n1=4096, n3=33)
    xyz = small_frame(100, 100, seed=3)
    assert np.array_equal(_gpu_u8(proc, xyz, stock, **OFF), _oracle_u8(xyz, stock, **OFF))


@pytest.mark.parametrize("kind", ["natural", "adversarial"])
def test_pointwise_24mp_bit_exact(proc, kind):
    """BASELINE config C1 at full size: 6000x4000, stages off, uint8 bit-exact."""
    stock = SyntheticStock()
    xyz = natural_frame(4000, 6000, 0) if kind == "natural" else adversarial_frame(4000, 6000, 0)
    got, want = _gpu_u8(proc, xyz, stock, **OFF), _oracle_u8(xyz, stock, **OFF)
    bad = np.count_nonzero(got != want)
    assert bad == 0, f"{bad} of {got.size} bytes differ (max |d| = {np.abs(got.astype(int) - want).max()})"


@pytest.mark.parametrize("kind", ["natural", "adversarial"])
def test_fast_chain_equals_exact_chain_and_defers_few_pixels(proc, kind):
    """The guarded float32 fast path (csrc/fast_chain.cuh) must give the bytes of the exact chain for every pixel,
    and only a small share of the pixels may need the exact fallback (otherwise it is not a fast path)."""
    stock = SyntheticStock()
    h, w = 2000, 3000
    xyz = natural_frame(h, w, 5) if kind == "natural" else adversarial_frame(h, w, 5)
    proc.fast_chain_stats()                              # reset the counter
    fast = _gpu_u8(proc, xyz, stock, **OFF)
    deferred, margin = proc.fast_chain_stats()
    assert 0 < margin < 0.05, margin                     # the default tables qualify, with a tight bound
    proc.set_fast_chain(False)
    try:
        exact = _gpu_u8(proc, xyz, stock, **OFF)
        assert proc.fast_chain_stats()[0] == 0
    finally:
        proc.set_fast_chain(True)
    assert np.array_equal(fast, exact)
    assert np.array_equal(exact, _oracle_u8(xyz, stock, **OFF))
    assert 0 < deferred < 0.15 * h * w, (deferred, h * w)
    print(f"fast chain [{kind}]: margin {margin:.2e}, deferred {deferred / (h * w):.3%} of the pixels")


def test_fast_chain_disqualified_tables_take_the_exact_chain(proc):
    """Tables outside the fast path's preconditions (3-D LUT values outside [0, 1]) silently use the exact chain."""

    class Hot(SyntheticStock):
        def create_lut(self, *a, **k):
            return (super().create_lut(*a, **k) * np.float32(1.5) - np.float32(0.2)).astype(np.float32)

    stock = Hot(name="lut outside unit range")
    xyz = small_frame(150, 210, seed=9)
    got = _gpu_u8(proc, xyz, stock, **OFF)
    assert proc.fast_chain_stats()[1] == -1.0
    assert np.array_equal(got, _oracle_u8(xyz, stock, **OFF))


def test_fast_chain_zero_density_toe(proc):
    """A curve whose toe is exactly 0 puts dark pixels ON the lower face of the 3-D LUT lattice, where the fast
    chain's roundings can land a hair below 0 (cell -1 of the guarded table): still bit-exact, still the fast path."""

    class ZeroToe(SyntheticStock):
        def get_density_curve(self, push_pull=0.0, color_masking=None):
            c = super().get_density_curve(push_pull, color_masking)
            c[1:] -= c[1:].min(axis=1, keepdims=True)        # every layer starts at exactly 0
            c[1:, :40] = 0.0                                   # ... and stays there for a while
            return c

    stock = ZeroToe(name="zero toe")
    rng = np.random.default_rng(8)
    xyz = small_frame(300, 400, seed=13)
    xyz[:100] *= np.float32(1e-5)                              # deep shadows: below the first curve sample
    xyz[100:120] = 0.0
    xyz[120:140] = rng.random((20, 400, 3), dtype=np.float32) * np.float32(1e-4)
    got = _gpu_u8(proc, xyz, stock, **OFF)
    assert proc.fast_chain_stats()[1] > 0                      # the tables qualify for the fast chain
    assert np.array_equal(got, _oracle_u8(xyz, stock, **OFF))
