"""Container-only: fuzz the oracle restatements against the live, unmodified reference functions
(imported through oracle/ref_loader.py).  Skipped wherever /root/reference is absent (GPU box)."""
import numpy as np
import pytest

from oracle import film_oracle as fo
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return load_reference()


@pytest.mark.parametrize("n,seed", [(5, 0), (33, 1)])
def test_tetra_fuzz(ref, n, seed):
    _, utils = ref
    rng = np.random.default_rng(seed)
    lut = rng.random((n, n, n, 3), dtype=np.float32)
    img = (rng.random((64, 96, 3), dtype=np.float32) * 4.2).astype(np.float32)
    assert np.array_equal(fo.apply_lut_tetrahedral(img, lut, 0.25), utils.apply_lut_tetrahedral(img, lut, 0.25))


def test_kernel_builders_fuzz(ref):
    effects, _ = ref
    rng = np.random.default_rng(5)
    for _ in range(6):
        scale = float(rng.uniform(15, 300))
        size, gf, inten = float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.1, 0.9)), float(rng.uniform(0.3, 2))
        a = fo.compute_halation_kernel(scale, size, 1.0, gf, 0.0, inten)
        b = effects.compute_halation_kernel(scale, halation_size=size, halation_green_factor=gf,
                                            halation_intensity=inten)
        assert np.array_equal(a, b)


def test_burn_fuzz(ref):
    effects, _ = ref

    class Stock:
        d_ref = (0.4, 0.55, 0.7)

    rng = np.random.default_rng(9)
    img = (rng.random((90, 140, 3), dtype=np.float32) * 3).astype(np.float32)
    assert np.array_equal(fo.burn(img.copy(), Stock.d_ref, 0.6, 12.0), effects.burn(img.copy(), Stock, 0.6, 12.0))
