"""The PRODUCT's host-side kernel builders (raw2film_b200/builders.py) against kernels produced by
the reference's own builders -- they must not depend on the oracle."""
import numpy as np

from raw2film_b200 import builders

G = "tests/golden/"


def test_halation_builder_bit_exact():
    g = np.load(G + "halation_kernels.npz")
    for i, c in enumerate(g["cases"]):
        k = builders.halation_kernel(c[0], c[1], 1.0, c[2], 0.0, c[3], bool(c[4]))
        assert k.dtype == np.float32 and np.array_equal(k, g[f"ref_kernel{i}"])


def test_mtf_builder_bit_exact():
    g = np.load(G + "mtf_kernels.npz")
    mtf = [(tuple(g["logf"][c]), tuple(g["vals"][c])) for c in range(3)]
    for i, c in enumerate(g["cases"]):
        k = builders.mtf_kernel(mtf, c[0], c[1], c[2])
        assert k.dtype == np.float32 and np.array_equal(k, g[f"ref_kernel{i}"])


def test_chroma_nr_taps_bit_exact():
    g = np.load(G + "chroma_nr.npz")
    for size in (1, 3, 8):
        assert np.array_equal(builders.chroma_nr_taps(size), g[f"ref_kernel{size}"])


def test_builders_do_not_import_oracle():
    import sys

    import raw2film_b200.builders  # noqa: F401
    import raw2film_b200.settings  # noqa: F401
    import raw2film_b200.synthetic  # noqa: F401

    src = open(builders.__file__).read()
    assert "oracle" not in src.replace("# oracle", "")
    for mod in ("processor", "batch", "_cabi", "settings", "synthetic", "builders", "pipeline", "hostops"):
        text = open(builders.__file__.replace("builders.py", mod + ".py")).read()
        assert "import oracle" not in text and "from oracle" not in text, mod
    assert "raw2film_b200" in sys.modules
