"""Shared helpers for the parity tests: a SyntheticStock -> oracle LUT dict adapter and
settings used on both sides."""
from __future__ import annotations

import numpy as np

from raw2film_b200 import settings as S
from raw2film_b200.synthetic import SyntheticStock


def oracle_luts(stock: SyntheticStock, s: dict, h: int, w: int) -> dict:
    scale = S.pixels_per_mm(h, w, s["frame_width"], s["frame_height"])
    bw_grain = s["grain"] == 1
    return {
        "lut2d": stock.get_input_lut(s["exp_kelvin"], s["tint"], s["exp_comp"]),
        "curve": stock.get_density_curve(push_pull=s["push_pull"], color_masking=s["color_masking"]),
        "lut3d": stock.create_lut(s["print_film"], **{k: s[k] for k in (
            "red_light", "green_light", "blue_light", "projector_kelvin", "shadow_comp", "sat_adjust", "gamma_func",
            "inversion_gamma", "idealized_curve", "inversion", "white_balance", "white_clip", "color_masking")},
            linear_scaling=4.0),
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
