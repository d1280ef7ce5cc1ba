"""Host-side logic that needs no GPU: settings merge / stage gating (cpu_processor.py:368-403),
pixel scale, frame sharding, the synthetic stock surface."""
import numpy as np
import pytest

from raw2film_b200 import flags as F
from raw2film_b200 import settings as S
from raw2film_b200.synthetic import SyntheticStock, adversarial_frame, mixed_stocks, natural_frame


def test_defaults_match_reference_signature():
    """Spot values of CpuProcessor.process defaults (cpu_processor.py:269-322)."""
    d = S.DEFAULTS
    assert d["halation_green_factor"] == 0.4 and d["grain"] == 2 and d["burn_scale"] == 50.0
    assert d["max_scale"] == 400.0 and d["frame_width"] == 36 and d["frame_height"] == 24
    assert d["canvas_mode"] == "No" and d["highlight_burn"] == 0.0 and d["exp_kelvin"] == 6500
    g = {**S.GUI_PROFILE_DEFAULTS, **S.GUI_IMAGE_DEFAULTS}          # gui.py:486-531
    assert g["halation_green_factor"] == 0.3 and g["exp_kelvin"] == 6000 and g["grain_size"] == 6


def test_merge_keeps_unknown_keys():
    s = S.merged({"film_format": "135", "profile": "Default", "grain": 1})
    assert s["film_format"] == "135" and s["grain"] == 1 and s["halation"] is True


def test_stage_gating():
    full = SyntheticStock()
    plain = SyntheticStock(with_mtf=False, with_grain=False, density_measure="status_a")
    assert S.stage_flags(S.merged({}), full) == F.HALATION | F.MTF | F.GRAIN
    assert S.stage_flags(S.merged({"grain": 1}), full) == F.HALATION | F.MTF | F.GRAIN | F.GRAIN_BW
    assert S.stage_flags(S.merged({"grain": 0, "halation": False, "sharpness": False}), full) == 0
    assert S.stage_flags(S.merged({}), plain) == F.HALATION                       # stock gates MTF/grain
    assert S.stage_flags(S.merged({"highlight_burn": 0.5}), full) & F.BURN          # status_m
    assert not S.stage_flags(S.merged({"highlight_burn": 0.5}), plain) & F.BURN     # no print film, status_a
    assert S.stage_flags(S.merged({"highlight_burn": 0.5, "print_film": full}), plain) & F.BURN


def test_pixels_per_mm_matches_survey_sizes():
    assert S.pixels_per_mm(4000, 6000, 36, 24) == pytest.approx(166.6667, rel=1e-6)
    assert S.pixels_per_mm(6336, 9504, 36, 24) == 264.0
    assert S.pixels_per_mm(6000, 4000, 24, 36) == pytest.approx(166.6667, rel=1e-6)


@pytest.mark.parametrize("n,world", [(64, 1), (64, 2), (64, 8), (7, 4), (0, 2), (3, 8)])
def test_shard_frames_partitions_exactly(n, world):
    shards = [S.shard_frames(n, world, r) for r in range(world)]
    allf = sorted(i for sh in shards for i in sh)
    assert allf == list(range(n))
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        S.shard_frames(4, 2, 2)


def test_synthetic_stock_surface():
    st = SyntheticStock(n2=32, n1=128, n3=9)
    lut2 = st.get_input_lut(6000, 0.0, 0.0)
    assert lut2.shape == (32, 32, 3) and lut2.dtype == np.float32 and lut2.min() > 0
    curve = st.get_density_curve(push_pull=0.0, color_masking=1.0)
    assert curve.shape == (4, 128) and np.all(np.diff(curve[0]) > 0) and np.all(np.diff(curve[1:], axis=1) >= 0)
    lut3 = st.create_lut(None)
    assert lut3.shape == (9, 9, 9, 3) and 0.0 <= lut3.min() and lut3.max() <= 1.0
    gc = st.get_grain_curve(166.0, adx=False, bw_grain=False)
    assert gc.shape == (4, 128) and gc[1:].min() > 0
    assert len(st.mtf) == 3 and st.rms_density is not None and len(st.d_ref) == 3
    assert hash(st) == hash(SyntheticStock(n2=32, n1=128, n3=9)) and st == SyntheticStock(n2=32, n1=128, n3=9)
    four = mixed_stocks(4)
    assert len({s.name for s in four}) == 4
    assert not np.array_equal(four[0].get_density_curve(), four[1].get_density_curve())


def test_synthetic_frames():
    a, b = natural_frame(96, 128, 0), natural_frame(96, 128, 0)
    assert a.shape == (96, 128, 3) and a.dtype == np.float32 and np.array_equal(a, b)
    assert not np.array_equal(a, natural_frame(96, 128, 1))
    assert a.min() >= 0 and a[..., 1].max() == 16.0
    adv = adversarial_frame(16, 16, 0)
    assert adv.min() >= 0 and adv.max() <= 2


def test_hostops_match_reference_goldens():
    """resolution_scaling (utils.py:226-244) and get_canvas_data (effects.py:290-335) restated in
    raw2film_b200/hostops.py against outputs of the reference functions."""
    from raw2film_b200 import hostops

    g = np.load("tests/golden/canvas_resize.npz")
    assert np.array_equal(hostops.resolution_scaling(g["img_resize"], (32, 32)), g["ref_down"])
    assert np.array_equal(hostops.resolution_scaling(g["img_resize"], (128, 400)), g["ref_up"])
    same = g["img_resize"]
    assert hostops.resolution_scaling(same, same.shape[:2]) is same
    modes = ["Proportional white", "Proportional black", "Uniform white", "Uniform black", "Fixed white", "Fixed black"]
    for i, m in enumerate(modes):
        size, colour, off = hostops.canvas_geometry(g["img"].shape, m, 1.2, 0.8)
        ref = g[f"ref_{i}"]
        assert tuple(size) == ref.shape[:2]
        assert tuple(ref[0, 0]) == colour
        assert np.array_equal(ref[off[0]:off[0] + 20, off[1]:off[1] + 30], g["img"])


def test_histogram_postprocessing_matches_reference_golden():
    from raw2film_b200 import hostops

    g = np.load("tests/golden/histogram.npz")
    counts = np.stack([np.bincount(g["img"][..., c].ravel(), minlength=256) for c in range(3)])
    for h in (100, 64):
        assert np.array_equal(hostops.histogram_image(counts, g["mix"], h), g[f"ref_hist{h}"])
    assert hostops.histogram_image(np.zeros((3, 256), np.int64), g["mix"], 10).shape == (10, 256, 4)


def test_present_geometry_letterbox_and_canvas():
    """_bind_copy_to_dst (reference gpu_processor.py:1416-1512) restated in hostops.present_geometry."""
    from raw2film_b200 import hostops

    # 3:2 image into a square widget: full width, centred vertically, no canvas rectangle
    sx, sy, ox, oy, *canvas = hostops.present_geometry((300, 200), (240, 240), (300, 200), (300, 200), None)
    assert (round(1 / sx), round(1 / sy), ox, oy) == (240, 160, 0.0, 40.0) and canvas == [0.0, 0.0, 0.0, 0.0]
    # tall widget, wide image with a canvas 20 % larger: the canvas fills the width, the image is inset
    sx, sy, ox, oy, x0, y0, x1, y1 = hostops.present_geometry((300, 200), (600, 900), (300, 200), (300, 200), (360, 240))
    assert (x0, x1) == (0.0, 600.0) and abs((y1 - y0) - 400.0) < 1e-9 and abs(y0 - 250.0) < 1e-9
    assert abs(1 / sx - 500.0) < 1e-9 and abs(ox - 50.0) < 1e-9 and abs(1 / sy - 400.0 * 200 / 240) < 1e-9
