"""Shared helpers for the parity tests: a SyntheticStock -> oracle LUT dict adapter and
settings used on both sides."""
from __future__ import annotations

import numpy as np

from raw2film_b200 import settings as S
from raw2film_b200.synthetic import SyntheticStock


def icc_through_8bit(lut: np.ndarray, icc_transform) -> np.ndarray:
    """The reference bakes an ICC transform into the 3-D LUT through an 8-bit PIL image
    (cpu_processor.py:255-263): restated here for the oracle side of the parity tests."""
    from PIL import Image, ImageCms

    shape = lut.shape
    img = Image.fromarray((lut * 255).astype(np.uint8).reshape(shape[0], -1, shape[-1]))
    ImageCms.applyTransform(img, icc_transform, inPlace=True)
    return (np.array(img, np.uint8).reshape(shape) / 255.0).astype(np.float32)


def make_test_icc_transform():
    """sRGB -> a wide-gamut matrix/TRC RGB profile (Adobe-RGB-like colorants, gamma 2.2) built in memory:
    there are no .icc files on the box.  Returns a PIL ImageCms transform like gui.py hands to the processors."""
    import io
    import struct

    from PIL import ImageCms

    def s15(v):
        return struct.pack(">i", int(round(v * 65536)))

    def xyz(x, y, z):
        return b"XYZ " + b"\0" * 4 + s15(x) + s15(y) + s15(z)

    def curv(gamma):
        return b"curv" + b"\0" * 4 + struct.pack(">I", 1) + struct.pack(">H", int(round(gamma * 256))) + b"\0\0"

    def pad4(d):
        return d + b"\0" * ((4 - len(d) % 4) % 4)

    def text(t):
        return pad4(b"text" + b"\0" * 4 + t.encode() + b"\0")

    def desc(t):
        b = t.encode() + b"\0"
        return pad4(b"desc" + b"\0" * 4 + struct.pack(">I", len(b)) + b + b"\0" * 78)

    tags = [(b"desc", desc("r2f test wide RGB")), (b"cprt", text("none")), (b"wtpt", xyz(0.9642, 1.0, 0.8249)),
            (b"rXYZ", xyz(0.60974, 0.31111, 0.01947)), (b"gXYZ", xyz(0.20528, 0.62567, 0.06087)),
            (b"bXYZ", xyz(0.14919, 0.06322, 0.74457)), (b"rTRC", curv(2.19921875)), (b"gTRC", curv(2.19921875)),
            (b"bTRC", curv(2.19921875))]
    off = 128 + 4 + 12 * len(tags)
    table = data = b""
    for sig, d in tags:
        table += sig + struct.pack(">II", off + len(data), len(d))
        data += d
    hdr = (struct.pack(">I", off + len(data)) + b"\0" * 4 + struct.pack(">I", 0x02100000) + b"mntr" + b"RGB " + b"XYZ "
           + b"\0" * 12 + b"acsp" + b"\0" * 28 + s15(0.9642) + s15(1.0) + s15(0.8249) + b"\0" * 48)
    assert len(hdr) == 128
    profile = ImageCms.ImageCmsProfile(io.BytesIO(hdr + struct.pack(">I", len(tags)) + table + data))
    return ImageCms.buildTransform(ImageCms.createProfile("sRGB"), profile, "RGB", "RGB")


def oracle_luts(stock: SyntheticStock, s: dict, h: int, w: int) -> dict:
    scale = S.pixels_per_mm(h, w, s["frame_width"], s["frame_height"])
    bw_grain = s["grain"] == 1
    lut3d = stock.create_lut(s["print_film"], **{k: s[k] for k in (
        "red_light", "green_light", "blue_light", "projector_kelvin", "shadow_comp", "sat_adjust", "gamma_func",
        "inversion_gamma", "idealized_curve", "inversion", "white_balance", "white_clip", "color_masking")},
        linear_scaling=4.0)
    if s.get("icc_transform") is not None:
        lut3d = icc_through_8bit(lut3d, s["icc_transform"])
    return {
        "lut2d": stock.get_input_lut(s["exp_kelvin"], s["tint"], s["exp_comp"]),
        "curve": stock.get_density_curve(push_pull=s["push_pull"], color_masking=s["color_masking"]),
        "lut3d": lut3d,
        "mtf": stock.mtf,
        "grain_curve": stock.get_grain_curve(scale, adx=False, bw_grain=bw_grain) if stock.rms_density is not None else None,
        "d_ref": stock.d_ref,
    }


def oracle_render(fo, xyz, stock, grain_size, grain_sigma, settings, noise=None, stages=None):
    """oracle.render with the same flat settings dict the processor takes."""
    s = S.merged(settings)
    h, w = xyz.shape[:2]
    luts = oracle_luts(stock, s, h, w)
    burn_on = bool(s["highlight_burn"]) and (s["print_film"] is not None or stock.density_measure in ["status_m", "bw"])
    return fo.render(
        xyz, luts, frame_width=s["frame_width"], frame_height=s["frame_height"], halation_on=bool(s["halation"]),
        halation_size=s["halation_size"], halation_green_factor=s["halation_green_factor"],
        halation_intensity=s["halation_intensity"], bw=stock.density_measure == "bw",
        sharpness=bool(s["sharpness"]), sharpening_strength=s["sharpening_strength"],
        sharpening_sigma=s["sharpening_sigma"], grain=s["grain"] if stock.rms_density is not None else 0,
        grain_size=grain_size, grain_sigma=grain_sigma, noise=noise,
        highlight_burn=s["highlight_burn"] if burn_on else 0.0, burn_scale=s["burn_scale"], stages=stages)


def small_frame(h, w, seed=0, highlights=True):
    """Small natural-ish XYZ frame with a wide exposure range and a few clipped highlights."""
    rng = np.random.default_rng(seed)
    y = (0.18 * np.exp2(rng.uniform(-6, 3, (h, w)))).astype(np.float32)
    cx = rng.uniform(0.2, 0.45, (h, w)).astype(np.float32)
    cy = rng.uniform(0.2, 0.45, (h, w)).astype(np.float32)
    if highlights:
        for _ in range(max(1, h * w // 2000)):
            py, px = int(rng.integers(0, h)), int(rng.integers(0, w))
            y[py, px] = 16.0
    out = np.empty((h, w, 3), np.float32)
    out[..., 0] = cx * y / cy
    out[..., 1] = y
    out[..., 2] = (1 - cx - cy) * y / cy
    return out
