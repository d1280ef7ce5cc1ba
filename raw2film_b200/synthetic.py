"""Deterministic synthetic film stocks and frames.

The reference gets its stocks from `spectral_film_lut.FilmSpectral`, which is not in the
reference tree and not installable offline.  `SyntheticStock` is an analytic stand-in with the
same SURFACE the processors consume (reference cpu_processor.py:151-182, 375-401;
effects.py:174, 233, 406; gpu_processor.py:860, 898, 913, 954): `.name`, `.density_measure`,
`.mtf`, `.rms_density`, `.d_ref`, `.get_input_lut()`, `.get_density_curve()`,
`.get_grain_curve()`, `.grain_transform()`, hashable.  It additionally offers `.create_lut()`,
the stand-in for the module-level `spectral_film_lut.utils.create_lut` (cpu_processor.py:232-253).
Table sizes (n2, N, n3) are parameters because the third-party defaults are not visible.
Every shape assumption here is UNVERIFIED against the real package.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32

# XYZ -> linear "film RGB" (a wide-gamut matrix; rows sum so that D65 white maps near (1,1,1))
_XYZ_TO_FILM = np.array([[1.64, -0.33, -0.24], [-0.67, 1.62, 0.02], [0.01, -0.09, 0.99]], dtype=np.float64)


def _mired_gain(kelvin: float, tint: float) -> np.ndarray:
    shift = (1e6 / 6500.0 - 1e6 / float(kelvin)) / 100.0
    return np.array([2.0 ** (-0.35 * shift), 2.0 ** (0.02 * tint), 2.0 ** (0.45 * shift - 0.02 * tint)])


class SyntheticStock:
    """Analytic negative stock.  `variant` selects one of the distinct parameter sets used for the
    mixed-stock batch config (BASELINE config 4)."""

    def __init__(self, name: str | None = None, variant: int = 0, n2: int = 64, n1: int = 1024, n3: int = 33,
                 density_measure: str = "status_m", with_mtf: bool = True, with_grain: bool = True,
                 warped_curve: bool = False):
        # The processors key their LUT caches on `.name` (reference cpu_processor.py:151, 174, 211), so two
        # stocks with different tables must not share a name: the default name encodes every parameter.
        if name is None:
            name = (f"Synthetic {100 * (int(variant) + 1)} [{n2}/{n1}/{n3} {density_measure}"
                    f"{'' if with_mtf else ' no-mtf'}{'' if with_grain else ' no-grain'}"
                    f"{' warped' if warped_curve else ''}]")
        self.name = name
        self.variant = int(variant)
        self.n2, self.n1, self.n3 = int(n2), int(n1), int(n3)
        self.density_measure = density_measure
        # warped_curve: the H-D and grain curves are sampled on a NON-uniform abscissa (denser around the
        # mid-tones), which the real package is free to do; lookups are then np.interp
        self.warped_curve = bool(warped_curve)
        v = self.variant
        self.d_ref = (0.50 + 0.03 * v, 0.60 + 0.02 * v, 0.70 + 0.01 * v)
        self.rms_density = (0.012 + 0.002 * v) if with_grain else None
        self.gamma = np.array([0.62, 0.66, 0.70]) * (1.0 + 0.04 * v)
        self.d_min = np.array([0.12, 0.18, 0.25]) + 0.01 * v
        self.d_max = np.array([2.9, 3.1, 3.3]) - 0.05 * v
        freqs = np.array([0.0, 5.0, 10.0, 20.0, 50.0, 100.0, 200.0])
        bump = 1.0 + 0.01 * v
        self.mtf = None
        if with_mtf:
            self.mtf = [
                (tuple(np.log1p(freqs)), (1.0, 1.03 * bump, 1.05 * bump, 0.92, 0.55, 0.22, 0.05)),
                (tuple(np.log1p(freqs)), (1.0, 1.04 * bump, 1.08 * bump, 0.98, 0.62, 0.27, 0.06)),
                (tuple(np.log1p(freqs)), (1.0, 1.02 * bump, 1.03 * bump, 0.85, 0.45, 0.15, 0.03)),
            ]

    # hashable + comparable by identity of its parameters (used as a cache key like FilmSpectral)
    def _key(self):
        return (self.name, self.variant, self.n2, self.n1, self.n3, self.density_measure, self.mtf is None,
                self.rms_density is None, self.warped_curve)

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        return isinstance(other, SyntheticStock) and self._key() == other._key()

    # ---- (n2, n2, 3) chromaticity LUT: film exposure per unit (X+Y+Z) -------------------------
    def get_input_lut(self, exp_kelvin=6500, tint=0.0, exp_comp=0.0) -> np.ndarray:
        n = self.n2
        x = np.arange(n, dtype=np.float64) / (n - 1)
        xx, yy = np.meshgrid(x, x, indexing="ij")           # lut[x_idx, y_idx]
        xyz = np.stack([xx, yy, 1.0 - xx - yy], axis=-1)
        rgb = xyz @ _XYZ_TO_FILM.T
        soft = 0.02
        rgb = 0.5 * (rgb + np.sqrt(rgb * rgb + soft * soft))  # smooth positive part
        rgb = rgb * _mired_gain(exp_kelvin, tint) * (2.0 ** float(exp_comp)) * 3.0
        return rgb.astype(F32)

    # ---- (4, N) H-D curve: row 0 log10 exposure (uniform), rows 1..3 density --------------------
    def get_density_curve(self, push_pull=0.0, color_masking=None) -> np.ndarray:
        loge = np.linspace(-4.0, 2.0, self.n1)
        if self.warped_curve:
            u = np.linspace(-1.0, 1.0, self.n1)
            loge = -1.0 + 3.0 * (0.35 * u + 0.65 * u ** 3)
        mask = 1.0 if color_masking is None else float(color_masking)
        rows = [loge]
        for c in range(3):
            g = self.gamma[c] * (1.0 + 0.15 * float(push_pull))
            mid = -1.2 + 0.1 * c
            span = self.d_max[c] - self.d_min[c]
            dens = self.d_min[c] * (0.5 + 0.5 * mask) + span / (1.0 + np.exp(-4.0 * g * (loge - mid) / span))
            rows.append(dens)
        return np.stack(rows).astype(F32)

    # ---- (4, N) grain amplitude over density -----------------------------------------------------
    def get_grain_curve(self, scale, adx=False, bw_grain=False) -> np.ndarray:
        dens = np.linspace(0.0, 4.0, self.n1)
        if self.warped_curve:
            dens = 4.0 * np.linspace(0.0, 1.0, self.n1) ** 1.7
        rms = self.rms_density or 0.0
        # Selwyn: rms over a 48 um aperture -> per-pixel sigma grows with sampling density
        amp = rms * math.sqrt(max(scale, 1.0) * 0.048 * math.sqrt(math.pi) / 2.0)
        rows = [dens]
        for c in range(3):
            peak = 1.1 + 0.15 * c
            shape = 0.35 + 0.65 * np.exp(-0.5 * ((dens - peak) / 0.9) ** 2)
            rows.append(amp * (1.0 + 0.2 * c) * shape * (0.8 if bw_grain else 1.0))
        return np.stack(rows).astype(F32)

    def grain_transform(self, rgb, scale, adx=False, bw_grain=False) -> np.ndarray:
        """CPU form of the per-pixel grain factor (reference effects.py:233); the CUDA path evaluates
        the same curve on device.  Provided only so the object has FilmSpectral's surface."""
        curve = self.get_grain_curve(scale, adx, bw_grain)
        return np.stack([np.interp(rgb[..., c], curve[0], curve[c + 1]) for c in range(3)], axis=-1).astype(F32)

    # ---- (n3, n3, n3, 3) output LUT: density*0.25 -> display RGB in [0,1] ---------------------------
    def create_lut(self, print_film=None, red_light=0.0, green_light=0.0, blue_light=0.0, projector_kelvin=6500,
                   shadow_comp=0.0, sat_adjust=1.0, gamma_func="sRGB", inversion_gamma=4.0, idealized_curve=False,
                   inversion=False, white_balance=False, white_clip=False, linear_scaling=4.0, color_masking=None,
                   **_) -> np.ndarray:
        n = self.n3
        d = np.arange(n, dtype=np.float64) / (n - 1) * float(linear_scaling)
        dr, dg, db = np.meshgrid(d, d, d, indexing="ij")     # lut[r, g, b]
        dens = np.stack([dr, dg, db], axis=-1)
        lights = np.array([red_light, green_light, blue_light], dtype=np.float64)
        rel = dens - np.asarray(self.d_ref) - 0.6 + 0.05 * lights
        pg = 2.4 if print_film is not None else float(inversion_gamma) * 0.45
        lin = 10.0 ** (-pg * 0.55 * rel) * 0.18               # print-through: more density -> darker print...
        lin = 0.18 * 0.18 / np.maximum(lin, 1e-9)             # ...of a negative: invert around mid grey
        cross = np.array([[0.88, 0.08, 0.04], [0.06, 0.89, 0.05], [0.03, 0.09, 0.88]])
        lin = lin @ cross.T
        luma = lin @ np.array([0.2126, 0.7152, 0.0722])
        lin = luma[..., None] + float(sat_adjust) * (lin - luma[..., None])
        lin = lin * _mired_gain(projector_kelvin, 0.0) + 0.002 * float(shadow_comp)
        lin = lin / (1.0 + lin) * 1.18                        # shoulder
        lin = np.clip(lin, 0.0, 1.0)
        if gamma_func == "sRGB":
            out = np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * np.power(lin, 1 / 2.4) - 0.055)
        else:
            out = np.power(lin, 1 / 2.2)
        return np.clip(out, 0.0, 1.0).astype(F32)


def mixed_stocks(count: int = 4, **kw) -> list[SyntheticStock]:
    return [SyntheticStock(variant=i, **kw) for i in range(count)]


# -----------------------------------------------------------------------------------------------
# frames
# -----------------------------------------------------------------------------------------------
def _smooth_field(rng, h, w, cells):
    """Band-limited noise: coarse Gaussian grid, bicubic-upsampled (cheap at 24-61 MP)."""
    import cv2 as cv

    gh, gw = max(4, h // cells), max(4, w // cells)
    coarse = rng.standard_normal((gh, gw)).astype(F32)
    coarse = cv.GaussianBlur(coarse, (0, 0), 1.5)
    coarse /= max(float(coarse.std()), 1e-6)
    return cv.resize(coarse, (w, h), interpolation=cv.INTER_CUBIC)


def natural_frame(h: int, w: int, frame_idx: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """'Natural' synthetic linear XYZ frame, float32 (H, W, 3): smooth log-luminance
    spanning 0.18 * 2^[-6,+3], D65-centred chromaticity wobble, 64 point highlights of Y=16."""
    rng = np.random.default_rng(1000 + frame_idx)
    cells = max(8, h // 64)
    lum = _smooth_field(rng, h, w, cells)
    fine = rng.standard_normal((h, w), dtype=F32) * F32(0.05)
    stops = np.clip(-1.5 + 2.2 * lum + fine, -6.0, 3.0)
    y = (0.18 * np.exp2(stops)).astype(F32)
    cx = np.clip(0.3127 + 0.04 * _smooth_field(rng, h, w, cells), 0.05, 0.6).astype(F32)
    cy = np.clip(0.3290 + 0.04 * _smooth_field(rng, h, w, cells), 0.05, 0.6).astype(F32)
    for _ in range(64):
        py, px = int(rng.integers(1, h - 1)), int(rng.integers(1, w - 1))
        y[py - 1:py + 2, px - 1:px + 2] = 16.0
    if out is None:
        out = np.empty((h, w, 3), F32)
    out[..., 0] = cx * y / cy
    out[..., 1] = y
    out[..., 2] = (1.0 - cx - cy) * y / cy
    return out


def adversarial_frame(h: int, w: int, frame_idx: int = 0) -> np.ndarray:
    """i.i.d. U(0,2) per component: worst-case LUT locality."""
    rng = np.random.default_rng(1000 + frame_idx)
    return (rng.random((h, w, 3), dtype=F32) * F32(2.0)).astype(F32)
