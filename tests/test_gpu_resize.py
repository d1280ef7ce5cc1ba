"""Device resolution_scaling (reference utils.py:226-244 = cv2.resize INTER_AREA / INTER_LANCZOS4):
r2f_resize against the reference's own outputs (tests/golden/resize.npz), against the pinned oracle at larger
sizes, and inside the processor API (preview `resolution`, `max_scale` round trip)."""
import numpy as np
import pytest

from oracle import film_oracle as fo
from oracle import resize_oracle as ro
from raw2film_b200.synthetic import SyntheticStock, natural_frame
from tests.helpers import oracle_render, small_frame

pytestmark = pytest.mark.gpu
G = "tests/golden/"


@pytest.fixture(scope="module")
def proc():
    from raw2film_b200 import B200Processor

    p = B200Processor(device=0)
    yield p
    p.close()


def _dev_resize(proc, img, size):
    import torch

    out = proc.resize_device(torch.from_numpy(np.ascontiguousarray(img)).cuda(), size)
    proc.stream.synchronize()
    return out.cpu().numpy()


def test_resize_matches_reference_goldens(proc):
    g = np.load(G + "resize.npz")
    names = [k[4:] for k in g.files if k.startswith("box_")]
    for name in names:
        for kind in ("f32", "u8"):
            src, want = g[kind], g[f"ref_{kind}_{name}"]
            got = _dev_resize(proc, src, want.shape[:2])
            assert got.shape == want.shape and got.dtype == want.dtype
            if want.shape[0] < src.shape[0] or kind == "u8":      # INTER_AREA, uint8 INTER_LANCZOS4: bit-exact
                assert np.array_equal(got, want), (name, kind, np.abs(got.astype(np.float64) - want).max())
            else:                                                 # float32 INTER_LANCZOS4: a few ulp (host SIMD FMA)
                assert np.abs(got - want).max() <= 1e-6 * float(src.max()), (name, kind)


@pytest.mark.parametrize("shape,size", [((1000, 1500), (333, 500)), ((1000, 1500), (500, 750)), ((999, 1501), (412, 619)),
                                        ((600, 900), (200, 300)), ((4000, 6000), (1080, 1620))])
def test_area_large_frames_vs_oracle_and_alpha_channel(proc, shape, size):
    """Preview-sized shrinks (24 MP -> 2 MP included) of float frames, 3 and 4 channels, bit-exact vs the oracle."""
    xyz = natural_frame(shape[0], shape[1], 9)
    want = ro.resize_area(xyz, size[1], size[0])
    assert np.array_equal(_dev_resize(proc, xyz, size), want)
    xyza = np.concatenate([xyz, np.ones_like(xyz[..., :1])], axis=-1)
    assert np.array_equal(_dev_resize(proc, xyza, size), want)


def test_lanczos_u8_large_vs_oracle(proc):
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (480, 720, 3), dtype=np.uint8)
    for size in [(960, 1440), (661, 991), (481, 721)]:
        assert np.array_equal(_dev_resize(proc, img, size), ro.resize_lanczos4(img, size[1], size[0]))


def test_process_with_preview_resolution_bit_exact(proc):
    """update_preview passes `resolution` = widget size (gui.py:2197-2209): the frame is shrunk (INTER_AREA) before
    the path.  Device resize + pointwise chain == oracle chain on the cv2-resized frame, bit for bit."""
    import cv2 as cv

    stock = SyntheticStock()
    xyz = natural_frame(1200, 1800, 2)
    st = dict(halation=False, sharpness=False, grain=0, resolution=(400, 640), max_scale=None)
    small = cv.resize(xyz, (600, 400), interpolation=cv.INTER_AREA)       # fit (400, 640) keeping 3:2 -> 400 x 600
    want = oracle_render(fo, small, stock, 6.0, 0.4, dict(st, resolution=None))
    got = proc.process(xyz, stock, 6.0, 0.4, **st)
    assert got.shape == (400, 600, 3) and np.array_equal(got, want)
    host = proc.process(xyz, stock, 6.0, 0.4, device_resize=False, cache=False, **st)   # host cv2 resize, as before
    assert np.array_equal(host, want)


def test_max_scale_round_trip_bit_exact(proc):
    """cpu_processor.py:122-134 + 411-412: a frame sampled finer than max_scale px/mm is shrunk before the path and
    the uint8 result enlarged (INTER_LANCZOS4) back to the original size."""
    import cv2 as cv

    stock = SyntheticStock()
    xyz = small_frame(300, 450, seed=6)
    # 450 px over a 1 mm frame = 450 px/mm > max_scale 300 -> factor 2/3 -> 200 x 300 inside the path
    st = dict(halation=False, sharpness=False, grain=0, frame_width=1.0, frame_height=2 / 3, max_scale=300.0)
    small = cv.resize(xyz, (300, 200), interpolation=cv.INTER_AREA)
    inner = oracle_render(fo, small, stock, 6.0, 0.4, dict(st, max_scale=None))
    want = cv.resize(inner, (450, 300), interpolation=cv.INTER_LANCZOS4)
    got = proc.process(xyz, stock, 6.0, 0.4, **st)
    assert got.shape == (300, 450, 3) and np.array_equal(got, want)


def test_preview_graph_replay_matches_direct_call_and_detects_stale_tables(proc):
    """PreviewGraph: a captured render replays bit-identically on new frame contents, and refuses to replay once the
    processor's tables changed (the graph holds the old table pointers)."""
    import torch
    from raw2film_b200 import PreviewGraph

    stock = SyntheticStock()
    st = dict(halation=False, sharpness=False, grain=0)
    a, b = small_frame(270, 480, seed=1), small_frame(270, 480, seed=2)
    frame = torch.from_numpy(a).cuda()
    pg = PreviewGraph(proc, frame, stock, 6.0, 0.4, **st)
    got_a = pg.replay().cpu().numpy()
    frame.copy_(torch.from_numpy(b).cuda())
    torch.cuda.synchronize()
    got_b = pg.replay().cpu().numpy()
    assert np.array_equal(got_a, oracle_render(fo, a, stock, 6.0, 0.4, st))
    assert np.array_equal(got_b, oracle_render(fo, b, stock, 6.0, 0.4, st))
    proc.render_device(frame, stock, 6.0, 0.4, **dict(st, exp_comp=0.5))      # uploads a new 2-D LUT
    with pytest.raises(RuntimeError, match="tables changed"):
        pg.replay()


@pytest.mark.parametrize("dst,canvas", [((300, 500), False), ((500, 300), False), ((360, 640), True)])
def test_present_blit_matches_wgsl_restatement(proc, dst, canvas):
    """shaders/copy_to_int.wgsl: letterboxed bilinear blit into the widget's RGBA8 buffer (canvas area filled,
    rest transparent), geometry as _bind_copy_to_dst builds it."""
    import torch
    from raw2film_b200 import hostops

    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    proc.pipeline_resolution, proc.output_resolution = (300, 200), (300, 200)
    proc.canvas_resolution = (360, 260) if canvas else None
    dst_t = torch.zeros((dst[0], dst[1], 4), dtype=torch.uint8, device="cuda")
    proc.present(torch.from_numpy(img).cuda(), dst_t, canvas_colour=(10, 128, 250))
    proc.stream.synchronize()
    t = hostops.present_geometry((300, 200), (dst[1], dst[0]), (300, 200), (300, 200), proc.canvas_resolution)
    want = fo.present(img, dst, t, (10, 128, 250))
    got = dst_t.cpu().numpy()
    assert np.array_equal(got, want)
    assert (got[..., 3] == 255).any() and ((got[..., 3] == 0).any() or canvas)
    if canvas:
        assert (got[..., :3] == np.array([10, 128, 250], np.uint8)).all(axis=-1).any()


def test_process_preloaded_presents_into_a_device_texture(proc):
    """gpu_processor.py:1866-1890: with a destination texture the result is presented and None is returned."""
    import torch

    stock = SyntheticStock()
    xyz = small_frame(120, 180, seed=3)
    st = dict(halation=False, sharpness=False, grain=0)
    payload = proc.extract_image_data_cpu(xyz, **st)
    tex = torch.zeros((240, 240, 4), dtype=torch.uint8, device="cuda")
    assert proc.process_preloaded(payload, stock, 6.0, 0.4, dst_texture=tex, **st) is None
    got = tex.cpu().numpy()
    want_img = oracle_render(fo, xyz, stock, 6.0, 0.4, st)
    from raw2film_b200 import hostops

    t = hostops.present_geometry((180, 120), (240, 240), (180, 120), (180, 120), None)
    assert np.array_equal(got, fo.present(want_img, (240, 240), t))
    assert got[0, 0, 3] == 0 and got[120, 120, 3] == 255          # letterbox bars above / below a 3:2 image
