"""uint16 ingest on device: the frame rawpy hands over (uint16 XYZ,
raw_conversion.py:38-48) is uploaded as is; the device applies the reference's
`astype(float32) / 65535.0` and `*= 2**calc_exposure` (raw_conversion.py:51-53).  The result must
equal rendering the host-converted float32 frame bit for bit."""
import numpy as np
import pytest

from oracle import film_oracle as fo
from raw2film_b200.synthetic import SyntheticStock
from tests.helpers import oracle_render

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proc():
    from raw2film_b200 import B200Processor

    p = B200Processor(device=0)
    yield p
    p.close()


def _u16_frame(h, w, seed):
    rng = np.random.default_rng(seed)
    lum = np.exp2(rng.uniform(-9, 0, (h, w, 1)))
    frame = np.clip(lum * rng.uniform(0.6, 1.0, (h, w, 3)) * 65535, 0, 65535).astype(np.uint16)
    edge = np.array([[0, 0, 0], [65535] * 3, [1, 0, 0], [0, 1, 0], [0, 0, 1], [65535, 0, 0], [2, 3, 5], [32768] * 3],
                    np.uint16)
    frame[0, :min(8, w)] = edge[:min(8, w)]
    return frame


def _host_linear(frame, gain):
    """raw_conversion.py:51-53 on the host (float32 division, float32 product)."""
    rgb = frame.astype(np.float32) / 65535.0
    rgb *= gain
    return rgb


def test_u16_conversion_is_exact_for_every_code_value(proc):
    import torch

    codes = np.arange(65536, dtype=np.uint16).reshape(256, 256, 1).repeat(3, axis=2)
    gain = 2 ** 1.37
    want = _host_linear(codes, gain)
    stock = SyntheticStock(n2=2)
    # a 2x2 LUT of ones makes the 2-D LUT stage return (X+Y+Z)*1 ... use the exposure tap of an identity-like LUT
    x = torch.from_numpy(codes).cuda()
    exp_u16 = proc.render_tap(x, "exposure", stock, 6.0, 0.4, input_gain=gain, halation=False).cpu().numpy()
    exp_f32 = proc.render_tap(torch.from_numpy(want).cuda(), "exposure", stock, 6.0, 0.4, halation=False).cpu().numpy()
    assert np.array_equal(exp_u16, exp_f32)
    assert np.array_equal(exp_f32, fo.apply_2d_lut(want, stock.get_input_lut(6500, 0.0, 0.0)))


@pytest.mark.parametrize("shape,alpha", [((97, 131), False), ((64, 80), True), ((33, 7), False)])
def test_u16_pointwise_bit_exact(proc, shape, alpha):
    stock = SyntheticStock()
    frame = _u16_frame(*shape, seed=shape[0])
    gain = 2 ** 0.75
    st = dict(halation=False, sharpness=False, grain=0)
    want = oracle_render(fo, _host_linear(frame, gain), stock, 6.0, 0.4, st)
    payload = proc.extract_image_data_cpu(frame, input_gain=gain, alpha=alpha, **st)
    assert payload["image_array"].dtype == np.uint16 and payload["image_array"].shape[2] == (4 if alpha else 3)
    got = proc.process_preloaded(payload, stock, 6.0, 0.4, **st)
    assert np.array_equal(got, want)


def test_u16_full_emulation_matches_float_path_and_oracle(proc):
    from raw2film_b200 import PipelinedRenderer

    stock = SyntheticStock(n3=17)
    frame = _u16_frame(200, 300, seed=3)
    gain = 2 ** 1.1
    noise = fo.white_noise((200, 300, 3), False, seed=8)
    st = dict(frame_width=3.0, frame_height=2.0, grain=2, grain_noise=noise)
    lin = _host_linear(frame, gain)
    a = proc.process(lin, stock, 6.0, 0.4, **st)
    b = proc.process_preloaded(proc.extract_image_data_cpu(frame, input_gain=gain, **st), stock, 6.0, 0.4, **st)
    assert np.array_equal(a, b), "device-side ingest must reproduce the host-side conversion bit for bit"
    want = oracle_render(fo, lin, stock, 6.0, 0.4, {k: v for k, v in st.items() if k != "grain_noise"}, noise=noise)
    diff = np.abs(b.astype(np.int16) - want.astype(np.int16))
    assert diff.max() <= 1 and np.mean(diff != 0) < 2e-3
    pipe = PipelinedRenderer(proc, depth=2)
    got = {}
    pipe.run([proc.extract_image_data_cpu(frame, input_gain=gain, **st)] * 3, stock, 6.0, 0.4,
             sink=lambda i, im: got.__setitem__(i, im.copy()), **st)
    assert all(np.array_equal(got[i], b) for i in range(3))
    assert pipe.h2d_bytes == 3 * 200 * 300 * 3 * 2


def test_calc_exposure_on_device_matches_reference_golden():
    """r2f_calc_exposure (strided power-mean reduction of color_processing.py:71-99 on the device) against
    values produced by the reference function: uint16 and float32 frames, 3 and 4 channels, odd sizes, with and
    without EXIF metadata.  The device sums float32-rounded terms in binary64 (the reference: float32 pairwise),
    so the bar is 1e-5 stops (a gain error of 7e-6 relative, far below one 16-bit code value)."""
    from raw2film_b200 import B200Processor
    from tests.test_oracle_golden import _exposure_cases

    g, metas = _exposure_cases()
    proc = B200Processor(device=0)
    try:
        for i in range(3):
            u16 = g[f"u16_{i}"]
            f32 = u16.astype(np.float32) / np.float32(65535.0)
            f32x4 = np.concatenate([f32, np.ones_like(f32[..., :1])], axis=2)
            for frame in (u16, f32, f32x4):
                got = [proc.calc_exposure(frame, metadata=m) for m in metas]
                assert np.allclose(got, g[f"ref_exp_{i}"], rtol=0, atol=1e-5), (i, frame.dtype, got)
        # full-size frame against the oracle restatement
        rng = np.random.default_rng(5)
        big = (rng.random((1200, 1800, 3)) ** 3 * 65535).astype(np.uint16)
        want = fo.calc_exposure(big.astype(np.float32) / np.float32(65535.0))
        assert abs(proc.calc_exposure(big) - want) <= 1e-5
    finally:
        proc.close()
