"""Round-2 parity cases: mixed stocks through the pipelined batch path (BASELINE config 4), every stock
variant under full emulation, burn together with the spatial stages, the ICC branch, table sizes other
than the defaults, a non-uniform curve abscissa, and config C3 at full size against the cv2 oracle."""
import numpy as np
import pytest

from oracle import film_oracle as fo
from raw2film_b200.synthetic import SyntheticStock, natural_frame
from tests.helpers import make_test_icc_transform, oracle_render, small_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proc():
    from raw2film_b200 import B200Processor

    p = B200Processor(device=0)
    yield p
    p.close()


def _lsb(got, want):
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    return int(diff.max()), float(np.mean(diff != 0))


# ---------------------------------------------------------------------------------------------------
# C4: frames of different stocks and settings back to back, nothing synchronised in between
# (reference gui_objects.py:65-115: the batch consumer renders whatever the producer queued next)
# ---------------------------------------------------------------------------------------------------
def test_pipelined_mixed_stocks_and_settings_match_oracle(proc):
    """12 frames through PipelinedRenderer(depth=3): stock = frame % 4, exp_comp and the stage set change
    from frame to frame.  Table uploads happen while earlier frames are still in flight; every frame must
    equal the oracle render made with ITS tables (pointwise frames bit-exact, full emulation <= 1 LSB)."""
    from raw2film_b200 import PipelinedRenderer

    stocks = [SyntheticStock(variant=v) for v in range(4)]
    h, w = 1200, 1800
    frames, jobs = [], []
    for i in range(12):
        xyz = natural_frame(h, w, 100 + i)
        full = i % 3 != 0
        st = dict(exp_comp=0.25 * (i % 5) - 0.5, halation=full, sharpness=full, grain=2 if full else 0,
                  frame_width=9.0, frame_height=6.0, halation_green_factor=0.3)
        noise = fo.white_noise(xyz.shape, False, seed=40 + i) if full else None
        frames.append(xyz)
        jobs.append((stocks[i % 4], st, noise, full))
    pipe = PipelinedRenderer(proc, depth=3)
    got = {}
    tickets = []
    for i, (stock, st, noise, full) in enumerate(jobs):
        payload = proc.extract_image_data_cpu(frames[i], **st)
        extra = {"grain_noise": noise} if full else {}
        tickets.append(pipe.submit(payload, stock, 6.0, 0.4, **st, **extra))
        if len(tickets) >= 3:
            t = tickets.pop(0)
            got[t] = pipe.result(t).copy()
    for t in tickets:
        got[t] = pipe.result(t).copy()
    for i, (stock, st, noise, full) in enumerate(jobs):
        want = oracle_render(fo, frames[i], stock, 6.0, 0.4, st, noise=noise)
        if full:
            mx, rate = _lsb(got[i], want)
            assert mx <= 1 and rate < 2e-3, (i, mx, rate)
        else:
            assert np.array_equal(got[i], want), f"pointwise frame {i} is not bit-exact"


def test_table_slots_are_per_stock_and_evict_lru():
    """Each stock name owns a slot; returning to a stock re-uses its tables (no rebuild), and with more stocks
    than slots the least recently used one is evicted and rebuilt correctly."""
    from raw2film_b200 import B200Processor

    p = B200Processor(device=0, table_slots=2)
    try:
        stocks = [SyntheticStock(variant=v, n3=17) for v in range(3)]
        xyz = small_frame(96, 128, seed=2)
        st = dict(halation=False, sharpness=False, grain=0)
        want = [oracle_render(fo, xyz, s, 6.0, 0.4, st) for s in stocks]
        calls = {"n": 0}
        orig = SyntheticStock.get_input_lut

        def counting(self, *a, **k):
            calls["n"] += 1
            return orig(self, *a, **k)

        SyntheticStock.get_input_lut = counting
        try:
            for order in ([0, 1, 0, 1], [2, 0, 1, 2]):
                for v in order:
                    assert np.array_equal(p.process(xyz, stocks[v], 6.0, 0.4, **st), want[v]), v
            # pass 1 builds stocks 0 and 1 once each and then alternates for free; in pass 2 every request
            # evicts the least recently used of the two slots: 2 (evicts 0), 0 (evicts 1), 1 (evicts 2), 2 (evicts 0)
            assert calls["n"] == 6
        finally:
            SyntheticStock.get_input_lut = orig
        assert len(p._slots) == 2
    finally:
        p.close()


# ---------------------------------------------------------------------------------------------------
# full emulation for every stock variant, burn with the spatial stages, ICC, table sizes
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [1, 2, 3])
def test_full_emulation_every_stock_variant(proc, variant):
    stock = SyntheticStock(variant=variant)
    xyz = small_frame(300, 420, seed=50 + variant)
    noise = fo.white_noise(xyz.shape, False, seed=9 + variant)
    st = dict(frame_width=3.5, frame_height=2.5, grain=2, halation_green_factor=0.3)
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st, noise=noise)
    got = proc.process(xyz, stock, 6.0, 0.4, grain_noise=noise, **st)
    mx, rate = _lsb(got, want)
    assert mx <= 1 and rate < 2e-3, (mx, rate)


@pytest.mark.parametrize("grain_mode", [2, 1, 0])
def test_burn_with_halation_mtf_grain(proc, grain_mode):
    """The staged path burn takes when the spatial stages are on (k_noise / conv(EPI_GRAIN) / k_burn_* / k_finish):
    float taps <= 1e-4, uint8 <= 1 LSB (cpu_processor.py:387-407)."""
    import torch

    stock = SyntheticStock(n3=17)
    xyz = small_frame(320, 480, seed=61)
    xyz[100:180, 200:330] *= 40.0                       # a burnt-out window so the mask is not empty
    noise = fo.white_noise(xyz.shape, grain_mode == 1, seed=3) if grain_mode else None
    st = dict(frame_width=4.0, frame_height=3.0, grain=grain_mode, halation_green_factor=0.3, highlight_burn=0.6,
              burn_scale=40.0)
    stages = {}
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st, noise=noise, stages=stages)
    assert stages["burn"].min() >= 0 and np.abs(stages["burn"] - stages.get("grain", stages["mtf"])).max() > 1e-2
    extra = {"grain_noise": noise} if grain_mode else {}
    x = torch.from_numpy(xyz).cuda()
    tap = proc.render_tap(x, "burn", stock, 6.0, 0.4, **st, **extra).cpu().numpy()
    assert np.abs(tap - stages["burn"]).max() <= 1e-4
    got = proc.process(xyz, stock, 6.0, 0.4, **st, **extra)
    mx, rate = _lsb(got, want)
    assert mx <= 1 and rate < 2e-3, (mx, rate)


def test_icc_transform_branch_bit_exact(proc):
    """cpu_processor.py:255-263: the ICC transform is applied to the 3-D LUT through an 8-bit PIL image."""
    icc = make_test_icc_transform()
    stock = SyntheticStock()
    xyz = small_frame(160, 240, seed=12)
    st = dict(halation=False, sharpness=False, grain=0, icc_transform=icc)
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st)
    plain = oracle_render(fo, xyz, stock, 6.0, 0.4, dict(st, icc_transform=None))
    assert not np.array_equal(want, plain)
    got = proc.process(xyz, stock, 6.0, 0.4, **st)
    assert np.array_equal(got, want)
    assert np.array_equal(proc.process(xyz, stock, 6.0, 0.4, **dict(st, icc_transform=None)), plain)


@pytest.mark.parametrize("n2,n1,n3", [(32, 256, 17), (64, 1024, 33), (128, 4096, 65), (32, 4096, 65), (128, 256, 33)])
def test_pointwise_bit_exact_over_table_sizes(proc, n2, n1, n3):
    """Third-party table sizes are not visible offline: every size combination, including
    tables too large for shared memory (n2 = 128) and the 4.4 MB n3 = 65 cube, must stay bit-exact."""
    stock = SyntheticStock(n2=n2, n1=n1, n3=n3)
    st = dict(halation=False, sharpness=False, grain=0)
    for xyz in (small_frame(257, 391, seed=n2 + n3), (np.random.default_rng(5).random((130, 200, 3)) * 2).astype(np.float32)):
        want = oracle_render(fo, xyz, stock, 6.0, 0.4, st)
        got = proc.process(xyz, stock, 6.0, 0.4, **st)
        assert np.array_equal(got, want), (n2, n1, n3)


def test_n3_65_full_size_pointwise_and_full_emulation(proc):
    """'default 33, also run 65'.  24 MP pointwise bit-exact on a row band; full emulation <= 1 LSB."""
    import torch

    stock = SyntheticStock(n3=65)
    h, w = 4000, 6000
    xyz = natural_frame(h, w, 3)
    st = dict(halation=False, sharpness=False, grain=0)
    got = proc.render_device(torch.from_numpy(xyz).cuda(), stock, 6.0, 0.4, **st).cpu().numpy()
    rows = slice(1700, 2100)
    want = fo.pointwise_chain(xyz[rows], stock.get_input_lut(6500, 0.0, 0.0), stock.get_density_curve(),
                              stock.create_lut(None))
    assert np.array_equal(got[rows], want)
    small = small_frame(280, 400, seed=77)
    noise = fo.white_noise(small.shape, False, seed=2)
    st2 = dict(frame_width=3.0, frame_height=2.0, grain=2, halation_green_factor=0.3)
    want2 = oracle_render(fo, small, stock, 6.0, 0.4, st2, noise=noise)
    got2 = proc.process(small, stock, 6.0, 0.4, grain_noise=noise, **st2)
    mx, rate = _lsb(got2, want2)
    assert mx <= 1 and rate < 2e-3, (mx, rate)


def test_nonuniform_curve_abscissa_follows_np_interp(proc):
    """a (4, N) table whose row 0 is not uniform is looked up with np.interp semantics, not with
    the uniform-grid shortcut (which would silently return wrong densities)."""
    import torch

    warped = SyntheticStock(warped_curve=True)
    assert not fo.abscissa_uniform(warped.get_density_curve()[0])
    xyz = small_frame(200, 300, seed=14)
    st = dict(halation=False, sharpness=False, grain=0)
    want = oracle_render(fo, xyz, warped, 6.0, 0.4, st)
    got = proc.process(xyz, warped, 6.0, 0.4, **st)
    assert np.array_equal(got, want)
    # the uniform shortcut on this table would be visibly different: make sure the test can tell
    fake = warped.get_density_curve().copy()
    fake[0] = np.linspace(fake[0, 0], fake[0, -1], fake.shape[1])
    dens_wrong = fo.multi_channel_interp(fo.log_clip(fo.apply_2d_lut(xyz, warped.get_input_lut())), fake)
    dens_right = proc.render_tap(torch.from_numpy(xyz).cuda(), "density", warped, 6.0, 0.4, **st).cpu().numpy()
    assert np.abs(dens_wrong - dens_right).max() > 0.05
    # grain amplitude curve on a non-uniform density abscissa, full emulation
    noise = fo.white_noise(xyz.shape, False, seed=8)
    st2 = dict(frame_width=3.0, frame_height=2.0, grain=2, halation_green_factor=0.3)
    want2 = oracle_render(fo, xyz, warped, 6.0, 0.4, st2, noise=noise)
    got2 = proc.process(xyz, warped, 6.0, 0.4, grain_noise=noise, **st2)
    mx, rate = _lsb(got2, want2)
    assert mx <= 1 and rate < 2e-3, (mx, rate)


def test_decreasing_abscissa_is_rejected(proc):
    from raw2film_b200 import _cabi

    class Bad(SyntheticStock):
        def get_density_curve(self, push_pull=0.0, color_masking=None):
            c = super().get_density_curve(push_pull, color_masking)
            c[0, 10], c[0, 11] = c[0, 11], c[0, 10]          # not monotonic, far from uniform
            c[0, 500] = c[0, 499]
            return c

    with pytest.raises(_cabi.R2FError, match="strictly increasing"):
        proc.load_density_curve(Bad(name="bad abscissa"), 0.0)


# ---------------------------------------------------------------------------------------------------
# C3 at full size against the cv2 oracle itself
# ---------------------------------------------------------------------------------------------------
def test_c3_61mp_full_emulation_vs_cv2_oracle(proc):
    """BASELINE config 3: 9504x6336, halation_size=2 (133x133 halation through the FFT path), MTF 27x27x3, RGB
    grain with the oracle's noise injected.  uint8 <= 1 LSB against oracle_render (cv2.filter2D)."""
    import torch

    stock = SyntheticStock()
    h, w = 6336, 9504
    xyz = natural_frame(h, w, 7)
    noise = fo.white_noise(xyz.shape, False, seed=11)
    st = dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3, halation_size=2.0)
    fo.use_all_host_threads()
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st, noise=noise)
    x = torch.from_numpy(xyz).cuda()
    got = proc.render_device(x, stock, 6.0, 0.4, grain_noise=noise, **st).cpu().numpy()
    assert proc.halation_kernel.shape[0] == 133 and proc.mtf_kernel.shape[0] == 27
    mx, rate = _lsb(got, want)
    assert mx <= 1 and rate < 2e-3, (mx, rate)


@pytest.mark.parametrize("st", [
    dict(halation=False, sharpness=False, grain=0),                                            # k_pointwise_fast per band
    dict(halation=True, sharpness=True, grain=2, grain_seed=11, halation_green_factor=0.3),    # FFT rows + fused grain
    dict(halation=True, sharpness=True, grain=0, frame_width=360.0, frame_height=240.0),       # k_expose, k_finish
    dict(halation=True, sharpness=False, grain=2, grain_seed=5, highlight_burn=0.5),           # burn: unbanded tail
])
def test_banded_streaming_call_equals_plain_render(proc, st):
    """process_preloaded streams the frame in and the result out in four bands around the first / last kernel
    (r2f_render_banded); the bytes must be those of the plain device-resident render."""
    import torch

    stock = SyntheticStock()
    xyz = natural_frame(1100, 1500, 21)            # 1100 rows -> four bands of 320 / 256 / 256 / 268 rows
    want = proc.render_device(torch.from_numpy(xyz).cuda(), stock, 6.0, 0.4, **st).cpu().numpy()
    payload = proc.extract_image_data_cpu(xyz, **st)
    got = proc.process_preloaded(payload, stock, 6.0, 0.4, **st)
    assert proc._own_pipeline().bands >= 4
    assert np.array_equal(got, want)
    u16 = np.clip(xyz * (65535.0 / 32.0), 0, 65535).astype(np.uint16)
    want16 = proc.render_device(torch.from_numpy(u16).cuda(), stock, 6.0, 0.4, input_gain=32.0, **st).cpu().numpy()
    got16 = proc.process_preloaded(proc.extract_image_data_cpu(u16, input_gain=32.0, **st), stock, 6.0, 0.4, **st)
    assert np.array_equal(got16, want16)


def test_banded_call_at_an_odd_width_equals_plain_render(proc):
    """Odd width (padded plane pitch between rows_inv, MTF and grain) through the banded synchronous call."""
    import torch

    stock = SyntheticStock()
    st = dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3, grain_seed=4321)
    xyz = natural_frame(1100, 1501, 29)
    want = proc.render_device(torch.from_numpy(xyz).cuda(), stock, 6.0, 0.4, **st).cpu().numpy()
    got = proc.process_preloaded(proc.extract_image_data_cpu(xyz, **st), stock, 6.0, 0.4, **st)
    assert proc._own_pipeline().bands >= 4 and np.array_equal(got, want)


def test_banded_mtf_switch_gives_the_same_bytes(proc):
    """R2F_OPT_FUSE_MTF: the MTF issued band by band with the grain kernel or as one whole-frame launch."""
    stock = SyntheticStock()
    st = dict(halation=True, sharpness=True, grain=2, halation_green_factor=0.3, grain_seed=1234)
    xyz = natural_frame(1100, 1500, 23)
    payload = proc.extract_image_data_cpu(xyz, **st)
    a = np.array(proc.process_preloaded(payload, stock, 6.0, 0.4, **st))
    proc.set_fuse_mtf(False)
    try:
        b = np.array(proc.process_preloaded(payload, stock, 6.0, 0.4, **st))
    finally:
        proc.set_fuse_mtf(True)
    assert np.array_equal(a, b)


def test_black_and_white_stock_halation_through_fft(proc):
    """B/W stocks filter all three layers with the same kernel (effects.py:248-250): the FFT path runs a second
    transform pair for the third layer.  Taps vs the oracle, FFT vs forced direct correlation, uint8 <= 1 LSB."""
    import torch

    stock = SyntheticStock(density_measure="bw")
    xyz = small_frame(400, 560, seed=19)
    st = dict(frame_width=14.0, frame_height=10.0, grain=0, sharpness=True, halation_green_factor=0.4)   # 40 px/mm: 11x11
    stages = {}
    want = oracle_render(fo, xyz, stock, 6.0, 0.4, st, stages=stages)
    x = torch.from_numpy(xyz).cuda()
    proc.set_conv_path("fft")
    try:
        hal_fft = proc.render_tap(x, "halation", stock, 6.0, 0.4, **st).cpu().numpy()
        dens_fft = proc.render_tap(x, "density", stock, 6.0, 0.4, **st).cpu().numpy()
        got = proc.process(xyz, stock, 6.0, 0.4, **st)
    finally:
        proc.set_conv_path("direct")
    try:
        hal_dir = proc.render_tap(x, "halation", stock, 6.0, 0.4, **st).cpu().numpy()
    finally:
        proc.set_conv_path("auto")
    assert proc.halation_kernel.shape[0] >= 9
    assert not np.array_equal(stages["halation"][..., 2], stages["exposure"][..., 2])     # the blue layer is filtered
    scale = float(stages["halation"].max())
    assert np.abs(hal_fft - stages["halation"]).max() <= 1e-5 * scale
    assert np.abs(hal_fft - hal_dir).max() <= 1e-5 * scale
    assert np.abs(dens_fft - stages["density"]).max() <= 1e-4
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert diff.max() <= 1 and np.mean(diff != 0) < 2e-3
