"""Host-side builders of the small spatial kernels (run once per settings change, cached by the
processor; SURVEY 8a rows a3/a6 "host builders").  NumPy only; the per-pixel work is CUDA.

Each builder mirrors a reference function and is checked against kernels produced by the
unmodified reference (tests/golden/halation_kernels.npz, mtf_kernels.npz).
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage

F32 = np.float32


def halation_kernel(scale: float, halation_size: float = 1.0, red_factor: float = 1.0, green_factor: float = 0.4,
                    blue_factor: float = 0.0, intensity: float = 1.0, bw: bool = False) -> np.ndarray:
    """(k, k, 3) float32 halation kernel.

    reference effects.py:200-217 (exponential_blur_kernel: 1/d^2 * max((r-d)/r, 0), centre 1,
    normalised) and effects.py:239-263 (per channel (f_c*K + delta)/(f_c + 1); B/W stocks use the
    green factor for all layers).  `size` = scale/4 * halation_size pixels (effects.py:252).
    """
    size = scale / 4 * halation_size
    radius = size / 2
    k = 2 * math.floor(math.ceil(size) / 2) + 1
    half = k // 2
    ax = np.arange(-half, half + 1, dtype=np.float64)
    d2 = np.add.outer(ax * ax, ax * ax)
    d2[half, half] = 1.0                                   # placeholder, centre is overwritten below
    base = np.maximum((radius - np.sqrt(d2)) / radius, 0.0) / d2
    base[half, half] = 1.0
    base = (base / base.sum()).astype(F32)
    if bw:
        red_factor = blue_factor = green_factor
    factors = F32(intensity) * np.array([red_factor, green_factor, blue_factor], dtype=F32)
    kern = base[:, :, None] * factors
    kern[half, half, :] += F32(1.0)
    return (kern / (factors + F32(1.0))).astype(F32)


def mtf_kernel(mtf, scale: float, sharpening_strength: float = 0.0, sharpening_sigma: float = 1.0) -> np.ndarray:
    """(k, k, 3) float32 MTF kernel from per-channel (log1p(f), response) polylines.

    reference effects.py:114-143 (radial response on the fftfreq grid, |ifft2|, fftshift,
    normalise), :159-162 (0.1 mm support), :165-185 (stack + optional unsharp term, which
    Gaussian-filters the stacked kernel along all three axes).
    """
    pixel_mm = 1 / scale
    k = round(0.1 / pixel_mm)
    k += 1 - (k % 2)
    freq = np.fft.fftfreq(k, d=pixel_mm)
    radial = np.hypot(freq[None, :], freq[:, None])
    layers = []
    for logf, vals in mtf:
        response = np.interp(np.log1p(radial), np.asarray(logf), np.asarray(vals), left=1, right=0)
        spatial = np.fft.fftshift(np.abs(np.fft.ifft2(response)))
        layers.append(spatial / spatial.sum())
    kern = np.stack(layers, axis=-1, dtype=F32)
    if sharpening_strength:
        soft = ndimage.gaussian_filter(kern, sigma=sharpening_sigma * scale / 50)
        kern += sharpening_strength * (kern - soft)
    return kern


def grain_kernel(pixel_size_mm: float, grain_size_mm: float = 0.01, grain_sigma: float = 0.4):
    """Stand-in for spectral_film_lut.grain_generation.grain_kernel (not in the reference tree;
    call site gpu_processor.py:927-929).  Used only when the real package is absent.

    3-point quadrature of log-normally distributed Gaussian grain blobs (median radius
    grain_size/2, log-std grain_sigma), unit L2 norm.  None when finer than the pixel grid
    (the caller then uses a 1x1 kernel, gpu_processor.py:931-932).
    """
    radius_px = 0.5 * grain_size_mm / pixel_size_mm
    if radius_px < 0.2:
        return None
    sigmas = [radius_px * math.exp(grain_sigma * q) for q in (-1.0, 0.0, 1.0)]
    half = max(1, int(math.ceil(3.0 * sigmas[-1])))
    ax = np.arange(-half, half + 1, dtype=np.float64)
    d2 = np.add.outer(ax * ax, ax * ax)
    kern = np.zeros_like(d2)
    for weight, s in zip((0.25, 0.5, 0.25), sigmas):
        kern += weight * np.exp(-d2 / (2 * s * s)) / (2 * math.pi * s * s)
    kern /= math.sqrt(float(np.sum(kern * kern)))
    return kern.astype(F32)


def chroma_nr_taps(chroma_nr: int) -> np.ndarray:
    """float32 Gaussian taps of the chroma noise reduction (reference effects.py:421-435, 552-554):
    size = 2*chroma_nr + 1, sigma = 0.3*((size-1)/2 - 1) + 0.8, normalised by the float32 running
    sum (numba sums a float32 array in float32)."""
    size = int(chroma_nr) * 2 + 1
    sigma = 0.3 * ((size - 1) * 0.5 - 1) + 0.8
    half = size // 2
    taps = np.array([math.exp(-((i - half) * (i - half)) / (2.0 * sigma * sigma)) for i in range(size)], dtype=F32)
    total = F32(0.0)
    for v in taps:
        total = F32(total + v)
    return (taps / total).astype(F32)
