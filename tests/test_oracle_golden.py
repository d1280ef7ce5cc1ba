"""The oracle against the golden vectors minted from the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle where /root/reference is absent."""
import numpy as np
import pytest

from oracle import film_oracle as fo

G = "tests/golden/"
MODES = ["Proportional white", "Proportional black", "Uniform white", "Uniform black", "Fixed white", "Fixed black"]


@pytest.mark.parametrize("n", [9, 17])
def test_tetrahedral_lut_bit_exact(n):
    """apply_lut_tetrahedral (utils.py:247-380): C restatement and NumPy twin, both bit-exact,
    including the top-of-table clamp, lattice hits, ties and the grey axis."""
    g = np.load(G + "tetra.npz")
    img, lut = g[f"img{n}"], g[f"lut{n}"]
    assert np.array_equal(fo.apply_lut_tetrahedral(img, lut, 0.25), g[f"ref_out{n}"])
    assert np.array_equal(fo.apply_lut_tetrahedral_np(img, lut, 0.25), g[f"ref_out{n}"])
    assert np.array_equal(fo.apply_lut_tetrahedral((img / 4.4).astype(np.float32), lut, 1.0), g[f"ref_out{n}_s1"])


def test_halation_kernels_bit_exact():
    """compute_halation_kernel (effects.py:239-263) at the BASELINE scales (43x43 @24MP, 133x133 @61MP x2)."""
    g = np.load(G + "halation_kernels.npz")
    sizes = []
    for i, c in enumerate(g["cases"]):
        k = fo.compute_halation_kernel(c[0], c[1], 1.0, c[2], 0.0, c[3], bool(c[4]))
        assert np.array_equal(k, g[f"ref_kernel{i}"])
        sizes.append(k.shape[0])
    assert sizes[0] == 43 and sizes[2] == 133


def test_mtf_kernels_bit_exact():
    """mtf_kernel (effects.py:165-185) incl. the unsharp term; 17x17x3 @24MP, 27x27x3 @61MP."""
    g = np.load(G + "mtf_kernels.npz")
    mtf = [(g["logf"][c], g["vals"][c]) for c in range(3)]
    sizes = []
    for i, c in enumerate(g["cases"]):
        k = fo.mtf_kernel(mtf, c[0], c[1], c[2])
        assert np.array_equal(k, g[f"ref_kernel{i}"])
        sizes.append(k.shape[0])
    assert sizes[:2] == [17, 27]


def test_convolve_2d_matches_reference_and_truth():
    """convolve_2d (effects.py:146-156): same cv2 call; also the float64 'mirror' truth proves
    correlation orientation + REFLECT_101."""
    g = np.load(G + "convolve.npz")
    for name in ("small", "big"):
        out = fo.convolve_2d(g["img"].copy(), g["k_" + name])
        assert np.array_equal(out, g["ref_" + name])
        assert np.abs(out - fo.correlate_truth_f64(g["img"], g["k_" + name])).max() < 5e-7


def test_burn_bit_exact():
    g = np.load(G + "burn.npz")
    assert np.array_equal(fo.burn(g["density"].copy(), g["d_ref"], *g["params"]), g["ref_out"])
    assert np.array_equal(fo.burn(g["density2"].copy(), g["d_ref"], *g["params2"]), g["ref_out2"])


def test_canvas_bit_exact():
    g = np.load(G + "canvas_resize.npz")
    for i, mode in enumerate(MODES):
        assert np.array_equal(fo.add_canvas(g["img"], mode, 1.2, 0.8), g[f"ref_{i}"])
    img = g["img"]
    assert fo.add_canvas(img, "No") is img


def test_chroma_nr_bit_exact():
    """chroma_nr_filter + gaussian_kernel_1d (effects.py:421-561) incl. the zero-denominator branches."""
    g = np.load(G + "chroma_nr.npz")
    for size in (1, 3, 8):
        n = size * 2 + 1
        assert np.array_equal(fo.gaussian_kernel_1d(n, 0.3 * ((n - 1) * 0.5 - 1) + 0.8), g[f"ref_kernel{size}"])
        assert np.array_equal(fo.chroma_nr_filter(g["xyz"], size), g[f"ref_out{size}"])


def _exposure_cases():
    g = np.load(G + "calc_exposure.npz")
    metas = []
    for f, iso, t in zip(g["meta_fnumber"], g["meta_iso"], g["meta_time"]):
        if iso == 0:
            metas.append(None)
        else:
            metas.append({"EXIF:FNumber": "undef" if f < 0 else float(f), "EXIF:ISO": float(iso),
                          "EXIF:ExposureTime": float(t)})
    return g, metas


def test_calc_exposure_matches_reference_golden():
    """calc_exposure (color_processing.py:71-99) on frames ingested like raw_conversion.py:51; the values were
    produced by the reference function itself.  Same NumPy expressions -> equal to the last bit on the NumPy
    that minted them; 1e-6 stops of slack for a different libm powf."""
    g, metas = _exposure_cases()
    for i in range(3):
        rgb = g[f"u16_{i}"].astype(np.float32) / np.float32(65535.0)
        got = [fo.calc_exposure(rgb, metadata=m) for m in metas]
        assert np.allclose(got, g[f"ref_exp_{i}"], rtol=0, atol=1e-6), (i, got, g[f"ref_exp_{i}"])


def test_resize_oracle_matches_reference_resolution_scaling():
    """oracle/resize_oracle.py against the reference's own resolution_scaling (cv2.resize) outputs: INTER_AREA
    (float32 and uint8, integer and fractional factors) and uint8 INTER_LANCZOS4 bit for bit; float32 INTER_LANCZOS4
    to 1e-6 of the data range (cv2's vertical pass is host-SIMD dependent, see the oracle's header)."""
    from oracle import resize_oracle as ro

    g = np.load(G + "resize.npz")
    names = [k[4:] for k in g.files if k.startswith("box_")]
    assert len(names) >= 12
    for name in names:
        box = tuple(int(v) for v in g["box_" + name])
        for kind in ("f32", "u8"):
            src, want = g[kind], g[f"ref_{kind}_{name}"]
            got = ro.resolution_scaling(src, box)
            assert got.shape == want.shape and got.dtype == want.dtype, (name, kind)
            shrink = want.shape[0] < src.shape[0]
            if shrink or kind == "u8":
                assert np.array_equal(got, want), (name, kind, np.abs(got.astype(np.float64) - want).max())
            else:
                assert np.abs(got - want).max() <= 1e-6 * float(src.max()), (name, kind)


def test_third_party_goldens():
    """Pins the WGSL-restated third-party stages against vectors minted from the real spectral_film_lut
    (`python tests/golden/make_golden.py third_party`).  The package is not installable offline, so the fixture
    does not exist yet and the stages stay PARITY UNPINNED; the test documents the procedure
    and starts guarding the day the fixture is committed."""
    import os

    import pytest

    path = G + "third_party.npz"
    if not os.path.exists(path):
        pytest.skip("tests/golden/third_party.npz not minted: spectral_film_lut is not installable offline")
    g = np.load(path)
    if "ref_apply_2d_lut" in g.files:
        assert np.allclose(fo.apply_2d_lut(g["xyz"], g["lut2d"]), g["ref_apply_2d_lut"], rtol=1e-6, atol=1e-7)
    if "ref_log_clip" in g.files:
        assert np.allclose(fo.log_clip(g["exposure"].copy()), g["ref_log_clip"], rtol=0, atol=1e-6)
    if "ref_interp" in g.files:
        assert np.allclose(fo.multi_channel_interp(g["logs"], g["curve"]), g["ref_interp"], rtol=0, atol=2e-6)
        assert np.allclose(fo.multi_channel_interp(g["logs"], g["curve_warped"]), g["ref_interp_warped"], rtol=0, atol=2e-6)
