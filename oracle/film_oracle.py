"""CPU oracle for the raw2film per-pixel film-emulation render path.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package
(raw2film_b200/) never imports it and has no CPU fallback.

What it is: a restatement of `CpuProcessor.process` (reference
src/raw2film/cpu_processor.py:363-407) and of the stage functions it calls
(src/raw2film/effects.py, src/raw2film/utils.py:247-380), assembled in the
reference's stage order.  Every function cites the reference lines it follows.

Pinning (see DESIGN.md "Oracle"):
  * in-tree reference functions (tetrahedral LUT, halation/MTF kernel builders,
    convolve_2d, burn, canvas) are PINNED: tests/golden/*.npz were produced by
    importing the unmodified reference (oracle/ref_loader.py,
    tests/golden/make_golden.py) and the restatements reproduce them.
  * the third-party stages (apply_2d_lut, log_clip, multi_channel_interp,
    generate_grain, grain_kernel, grain_transform from spectral-film-lut >=0.8.0,
    reference pyproject.toml:28) are NOT in the reference tree and cannot be
    installed here: they are "PARITY UNPINNED" and follow the WGSL restatements
    the reference ships (shaders/lut_2d.wgsl, lut_1d.wgsl, grain.wgsl, noise.wgsl).

Like the reference, spatial filters go through cv2.filter2D (effects.py:146-156)
so the oracle carries the same library arithmetic (correlation, anchor at the
centre, BORDER_REFLECT_101, DFT path for kernels larger than 11x11).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import cv2 as cv
import numpy as np
from scipy import ndimage

F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

LOG_CLIP_EPS = 1e-6  # shaders/lut_1d.wgsl:24 (CPU value of the third-party log_clip unverified)


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/pointwise_oracle.c -> liboracle.so (gcc, OpenMP when available)."""
    src = os.path.join(_HERE, "pointwise_oracle.c")
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
        return _LIB_PATH
    base = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math"]
    for extra in (["-fopenmp"], []):
        r = subprocess.run(base + extra + ["-o", _LIB_PATH, src, "-lm"], capture_output=True, text=True,
                           env={**os.environ, "PATH": "/usr/bin:/bin:" + os.environ.get("PATH", "")})
        if r.returncode == 0:
            return _LIB_PATH
    raise RuntimeError("could not build oracle C library:\n" + r.stderr)


def _c():
    global _lib
    if _lib is None:
        build_c_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        fp, u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
        i64, ci, cf, cd = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double
        lib.orc_num_threads.restype = ci
        lib.orc_set_num_threads.argtypes = [ci]
        lib.orc_set_num_threads.restype = None
        lib.orc_apply_2d_lut.argtypes = [fp, i64, ci, fp, ci, fp]
        lib.orc_log_clip.argtypes = [fp, i64, cf]
        lib.orc_curve_interp.argtypes = [fp, i64, fp, ci, cf, fp]
        lib.orc_tetra.argtypes = [fp, i64, fp, ci, cd, fp]
        lib.orc_quantise_u8.argtypes = [fp, i64, u8p]
        lib.orc_clip_min0.argtypes = [fp, i64]
        lib.orc_pointwise_chain.argtypes = [fp, i64, ci, fp, ci, fp, ci, cf, cf, fp, ci, cd, u8p]
        for f in ("orc_apply_2d_lut", "orc_log_clip", "orc_curve_interp", "orc_tetra", "orc_quantise_u8",
                  "orc_clip_min0", "orc_pointwise_chain"):
            getattr(lib, f).restype = None
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _c32(a):
    return np.ascontiguousarray(a, dtype=F32)


def num_threads() -> int:
    return int(_c().orc_num_threads())


def use_all_host_threads() -> int:
    """Give the OpenMP stages and cv2 every host core (torchrun exports OMP_NUM_THREADS=1)."""
    n = os.cpu_count() or 1
    _c().orc_set_num_threads(n)
    cv.setNumThreads(n)
    return n


# --------------------------------------------------------------------------------------
# a2  apply_2d_lut  (cpu_processor.py:364; third-party; shaders/lut_2d.wgsl:18-108)
# --------------------------------------------------------------------------------------
def apply_2d_lut(image: np.ndarray, lut: np.ndarray) -> np.ndarray:
    image = _c32(image)
    lut = _c32(lut)
    h, w, cin = image.shape
    assert cin in (3, 4) and lut.shape == (lut.shape[0], lut.shape[0], 3)
    out = np.empty((h, w, 3), F32)
    _c().orc_apply_2d_lut(_fp(image), h * w, cin, _fp(lut), lut.shape[0], _fp(out))
    return out


def apply_2d_lut_np(image: np.ndarray, lut: np.ndarray) -> np.ndarray:
    """NumPy twin of the C function (same float32 operation order) used to cross-check it."""
    x, y, z = (image[..., k].astype(F32) for k in range(3))
    n = lut.shape[0]
    s = (x + y) + z
    dark = s < F32(1e-12)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = F32(n - 1) / np.where(dark, F32(1), s)
        r, g = x * inv, y * inv
        rfl, gfl = np.floor(r), np.floor(g)
        ri = np.clip(np.nan_to_num(rfl, nan=0.0), 0, n - 2).astype(np.int64)
        gi = np.clip(np.nan_to_num(gfl, nan=0.0), 0, n - 2).astype(np.int64)
        rf, gf = r - rfl, g - gfl
        fs = rf + gf
        lower = fs <= F32(1.0)
        a, b = lut[ri + 1, gi], lut[ri, gi + 1]
        c = np.where(lower[..., None], lut[ri, gi], lut[ri + 1, gi + 1])
        wa = np.where(lower, rf, F32(1) - gf)[..., None]
        wb = np.where(lower, gf, F32(1) - rf)[..., None]
        wc = np.where(lower, F32(1) - fs, fs - F32(1))[..., None]
        out = ((a * wa + b * wb) + c * wc) * s[..., None]
    out[dark] = 0
    return out.astype(F32)


# --------------------------------------------------------------------------------------
# a4  log_clip  (cpu_processor.py:378; third-party; shaders/lut_1d.wgsl:23-26)
# --------------------------------------------------------------------------------------
def log_clip(image: np.ndarray, eps: float = LOG_CLIP_EPS) -> np.ndarray:
    """In place, like the reference call site (its return value is discarded)."""
    assert image.dtype == F32 and image.flags.c_contiguous
    _c().orc_log_clip(_fp(image), image.size, eps)
    return image


def log_clip_np(image: np.ndarray, eps: float = LOG_CLIP_EPS) -> np.ndarray:
    return np.log10(np.maximum(image, F32(eps)).astype(np.float64)).astype(F32)


# --------------------------------------------------------------------------------------
# a5  multi_channel_interp  (cpu_processor.py:380; third-party; lut_1d.wgsl:43-47,
#     (4,N) layout gpu_processor.py:307-333)
# --------------------------------------------------------------------------------------
def curve_inv_range(curve: np.ndarray) -> np.float32:
    """float32(1/(x_last - x_first)) evaluated in double (gpu_processor.py:322-325)."""
    d = float(curve[0, -1]) - float(curve[0, 0])
    return F32(1.0 / d) if d != 0.0 else F32(0.0)


def abscissa_uniform(xp: np.ndarray) -> bool:
    """Row 0 of a (4, N) table counts as uniform when every sample lies within 1e-3 of a step of the straight
    line between its ends (e.g. a float32 linspace).  Uniform tables take the reference GPU path's normalised
    lookup (lut_1d.wgsl:43-47, gpu_processor.py:322-325); anything else is an np.interp."""
    xp = np.asarray(xp, np.float64)
    n = xp.shape[0]
    step = (xp[-1] - xp[0]) / (n - 1)
    if not step > 0.0:
        return step == 0.0
    return bool(np.all(np.abs(xp - (xp[0] + step * np.arange(n))) <= 1e-3 * step))


def multi_channel_interp_nonuniform(image: np.ndarray, curve: np.ndarray) -> np.ndarray:
    """np.interp semantics per channel on the table's own abscissa: clamped ends, bracketing samples by
    binary search, slope and offset in binary64 (separate multiply and add), one rounding to binary32."""
    image, curve = _c32(image), _c32(curve)
    xp = curve[0].astype(np.float64)
    n = xp.shape[0]
    out = np.empty_like(image)
    for k in range(3):
        fp = curve[k + 1].astype(np.float64)
        v = image[..., k].astype(np.float64)
        j = np.clip(np.searchsorted(xp, v, side="right") - 1, 0, n - 2)
        slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j])
        res = slope * (v - xp[j]) + fp[j]
        res = np.where(~(v > xp[0]), fp[0], res)              # includes NaN -> first sample
        res = np.where(v >= xp[-1], fp[-1], res)
        out[..., k] = res.astype(F32)
    return out


def multi_channel_interp(image: np.ndarray, curve: np.ndarray) -> np.ndarray:
    image, curve = _c32(image), _c32(curve)
    assert curve.shape[0] == 4 and image.shape[-1] == 3
    if not abscissa_uniform(curve[0]):
        return multi_channel_interp_nonuniform(image, curve)
    out = np.empty_like(image)
    _c().orc_curve_interp(_fp(image), image.size // 3, _fp(curve), curve.shape[1], curve_inv_range(curve), _fp(out))
    return out


def multi_channel_interp_np(image: np.ndarray, curve: np.ndarray) -> np.ndarray:
    n = curve.shape[1]
    t = (image.astype(F32) - curve[0, 0]) * curve_inv_range(curve)
    t = np.nan_to_num(np.clip(t, F32(0), F32(1)), nan=0.0).astype(F32)
    p = t * F32(n - 1)
    i = np.minimum(p.astype(np.int64), n - 2)
    f = p - i.astype(F32)
    out = np.empty_like(p)
    for k in range(3):
        row = curve[k + 1]
        out[..., k] = row[i[..., k]] + f[..., k] * (row[i[..., k] + 1] - row[i[..., k]])
    return out


# --------------------------------------------------------------------------------------
# a9  apply_lut_tetrahedral  (utils.py:247-380)  -- PINNED bit-exact
# --------------------------------------------------------------------------------------
def apply_lut_tetrahedral(image: np.ndarray, lut: np.ndarray, scale: float = 1.0) -> np.ndarray:
    image, lut = _c32(image), _c32(lut)
    out = np.empty_like(image)
    _c().orc_tetra(_fp(image), image.size // 3, _fp(lut), lut.shape[0], float(scale), _fp(out))
    return out


def apply_lut_tetrahedral_np(image: np.ndarray, lut: np.ndarray, scale: float = 1.0) -> np.ndarray:
    """Vectorised twin: binary64 coordinates, binary32 vertex differences, binary64 sum."""
    n = lut.shape[0]
    v = image.astype(np.float64) * (np.float64(scale) * (n - 1))
    i0 = np.trunc(np.clip(v, -2e9, 2e9)).astype(np.int64)
    top = i0 >= n - 1
    d = np.where(top, 1.0, v - i0)
    i0 = np.where(top, n - 2, i0)
    i1 = i0 + 1
    i0, i1 = np.where(i0 < 0, np.maximum(i0 + n, 0), i0), np.where(i1 < 0, np.maximum(i1 + n, 0), i1)
    dr, dg, db = d[..., 0], d[..., 1], d[..., 2]
    sel = [i0, i1]

    def vert(r, g, b):
        return lut[sel[r][..., 0], sel[g][..., 1], sel[b][..., 2]]

    a = dr >= dg
    conds = [a & (dg >= db), a & ~(dg >= db) & (dr >= db), a & ~(dg >= db) & ~(dr >= db),
             ~a & (db >= dg), ~a & ~(db >= dg) & (db >= dr), ~a & ~(db >= dg) & ~(db >= dr)]
    # (first vertex, second vertex, ordered fractions) per tetrahedron, utils.py:298-376
    paths = [((1, 0, 0), (1, 1, 0), (dr, dg, db)), ((1, 0, 0), (1, 0, 1), (dr, db, dg)),
             ((0, 0, 1), (1, 0, 1), (db, dr, dg)), ((0, 0, 1), (0, 1, 1), (db, dg, dr)),
             ((0, 1, 0), (0, 1, 1), (dg, db, dr)), ((0, 1, 0), (1, 1, 0), (dg, dr, db))]
    c000, c111 = vert(0, 0, 0), vert(1, 1, 1)
    out = np.zeros(image.shape, np.float64)
    for cond, (v1, v2, (d1, d2, d3)) in zip(conds, paths):
        m1, m2 = vert(*v1), vert(*v2)
        acc = ((c000.astype(np.float64) + d1[..., None] * (m1 - c000).astype(np.float64))
               + d2[..., None] * (m2 - m1).astype(np.float64)) + d3[..., None] * (c111 - m2).astype(np.float64)
        out[cond] = acc[cond]
    return out.astype(F32)


# --------------------------------------------------------------------------------------
# a10 quantise  (cpu_processor.py:407)
# --------------------------------------------------------------------------------------
def quantise_u8(image: np.ndarray) -> np.ndarray:
    image = _c32(image)
    out = np.empty(image.shape, np.uint8)
    _c().orc_quantise_u8(_fp(image), image.size, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out


def clip_min0(image: np.ndarray) -> np.ndarray:
    _c().orc_clip_min0(_fp(image), image.size)
    return image


def pointwise_chain(xyz, lut2d, curve, lut3d, eps: float = LOG_CLIP_EPS, tetra_scale: float = 0.25) -> np.ndarray:
    """cpu_processor.py:364,378,380,405,407 in one multi-threaded pass (halation/MTF/grain off)."""
    xyz, lut2d, curve, lut3d = _c32(xyz), _c32(lut2d), _c32(curve), _c32(lut3d)
    h, w, cin = xyz.shape
    out = np.empty((h, w, 3), np.uint8)
    _c().orc_pointwise_chain(_fp(xyz), h * w, cin, _fp(lut2d), lut2d.shape[0], _fp(curve), curve.shape[1],
                             curve_inv_range(curve), eps, _fp(lut3d), lut3d.shape[0], tetra_scale,
                             out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out


# --------------------------------------------------------------------------------------
# a3  halation  (effects.py:200-217, 239-287)  -- builders PINNED by golden kernels
# --------------------------------------------------------------------------------------
def halation_kernel_size(size: float) -> int:
    return 2 * math.floor(math.ceil(size) / 2) + 1          # effects.py:204


def exponential_blur_kernel(size: float) -> np.ndarray:
    """K = 1/d^2 * max((r-d)/r, 0), centre 1, normalised; float64 (effects.py:200-217)."""
    radius = size / 2
    k = halation_kernel_size(size)
    off = np.arange(k, dtype=np.float64) - (k // 2)
    dist = off[:, None] ** 2 + off[None, :] ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        kern = (1.0 / dist) * np.maximum((radius - np.sqrt(dist)) / radius, 0.0)
    kern[k // 2, k // 2] = 1.0
    return kern / kern.sum()


def compute_halation_kernel(scale, halation_size=1.0, halation_red_factor=1.0, halation_green_factor=0.4,
                            halation_blue_factor=0.0, halation_intensity=1.0, bw=False) -> np.ndarray:
    """Per channel (f_c*K + delta)/(f_c + 1) in float32 (effects.py:239-263)."""
    if bw:
        halation_red_factor = halation_blue_factor = halation_green_factor
    base = exponential_blur_kernel(scale / 4 * halation_size).astype(F32)
    kern = np.repeat(base[:, :, None], 3, axis=2)
    fac = F32(halation_intensity) * np.array([halation_red_factor, halation_green_factor, halation_blue_factor], F32)
    kern = kern * fac
    mid = kern.shape[0] // 2
    kern[mid, mid, :] += F32(1.0)
    kern = kern / (fac + F32(1.0))
    return kern.astype(F32)


def convolve_2d(rgb: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """Per-channel cv2.filter2D, written into the input array (effects.py:146-156)."""
    if kernel.ndim == 2:
        return cv.filter2D(rgb, -1, kernel)
    for c in range(kernel.shape[-1]):
        rgb[..., c] = cv.filter2D(rgb[..., c], -1, kernel[..., c])
    return rgb


def correlate_truth_f64(rgb: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """float64 direct correlation with mirror (= REFLECT_101) borders: the error-budget truth."""
    out = np.empty(rgb.shape, np.float64)
    for c in range(rgb.shape[-1]):
        kc = kernel[..., c] if kernel.ndim == 3 else kernel
        out[..., c] = ndimage.correlate(rgb[..., c].astype(np.float64), kc.astype(np.float64), mode="mirror")
    return out


def halation(rgb, scale, halation_size=1.0, halation_green_factor=0.4, halation_intensity=1.0, bw=False):
    kern = compute_halation_kernel(scale, halation_size, 1.0, halation_green_factor, 0.0, halation_intensity, bw)
    return convolve_2d(rgb, kern)                            # effects.py:266-287


# --------------------------------------------------------------------------------------
# a6  film_sharpness / mtf_kernel  (effects.py:114-197)  -- builders PINNED by golden kernels
# --------------------------------------------------------------------------------------
def mtf_kernel_size(scale: float) -> int:
    k = round(0.1 / (1 / scale))                            # effects.py:125, 161
    return k + 1 if k % 2 == 0 else k


def mtf_kernel_layer(logf, vals, scale) -> np.ndarray:
    """|ifft2| of a radial MTF sampled on the fftfreq grid (effects.py:114-143, 159-162)."""
    pixel = 1 / scale
    k = mtf_kernel_size(scale)
    fr = np.fft.fftfreq(k, d=pixel)
    rad = np.sqrt(fr[None, :] ** 2 + fr[:, None] ** 2)
    resp = np.interp(np.log1p(rad), np.asarray(logf), np.asarray(vals), left=1, right=0)
    kern = np.fft.fftshift(np.abs(np.fft.ifft2(resp)))
    return kern / np.sum(kern)


def mtf_kernel(mtf, scale, sharpening_strength=0.0, sharpening_sigma=1.0) -> np.ndarray:
    """`mtf` is the stock's list of (logf, vals) per channel (effects.py:165-185)."""
    kern = np.stack([mtf_kernel_layer(lf, vs, scale) for lf, vs in mtf], axis=-1, dtype=F32)
    if sharpening_strength:
        blurred = ndimage.gaussian_filter(kern, sigma=sharpening_sigma * scale / 50)  # all 3 axes, :181
        kern += sharpening_strength * (kern - blurred)
    return kern


def film_sharpness(rgb, mtf, scale, sharpening_strength=0.0, sharpening_sigma=1.0):
    return convolve_2d(rgb, mtf_kernel(mtf, scale, sharpening_strength, sharpening_sigma))


# --------------------------------------------------------------------------------------
# a7  grain  (effects.py:220-236; third-party generate_grain / grain_kernel / grain_transform;
#     GPU restatement shaders/grain.wgsl:36-92, noise.wgsl, gpu_processor.py:904-936)
#     PARITY UNPINNED: the grain kernel shape is a documented stand-in.
# --------------------------------------------------------------------------------------
def grain_kernel(pixel_size_mm: float, grain_size_mm: float = 0.01, grain_sigma: float = 0.4):
    """Stand-in for spectral_film_lut.grain_generation.grain_kernel (call: gpu_processor.py:927-929).

    Radial kernel = 3-point quadrature of a log-normal distribution of Gaussian grain blobs
    (median radius grain_size/2, log-std grain_sigma), normalised to unit L2 norm so that
    filtering unit white noise keeps unit variance.  Returns None when the grain is finer
    than the pixel grid (the reference then substitutes a 1x1 kernel, gpu_processor.py:931-932).
    """
    radius_px = 0.5 * grain_size_mm / pixel_size_mm
    if radius_px < 0.2:
        return None
    sig = [radius_px * math.exp(grain_sigma * q) for q in (-1.0, 0.0, 1.0)]
    half = max(1, int(math.ceil(3.0 * sig[-1])))
    off = np.arange(-half, half + 1, dtype=np.float64)
    d2 = off[:, None] ** 2 + off[None, :] ** 2
    kern = sum(w * np.exp(-d2 / (2 * s * s)) / (2 * math.pi * s * s) for w, s in zip((0.25, 0.5, 0.25), sig))
    kern /= math.sqrt(np.sum(kern ** 2))
    return kern.astype(F32)


def white_noise(shape, bw_grain: bool, seed: int) -> np.ndarray:
    h, w = shape[:2]
    return np.random.default_rng(seed).standard_normal((h, w, 1 if bw_grain else 3), dtype=F32)


def generate_grain(shape, scale, grain_size_mm, bw_grain, grain_sigma, noise=None, seed=0) -> np.ndarray:
    """Unit grain field: white N(0,1) noise correlated with grain_kernel (grain.wgsl:50-75)."""
    if noise is None:
        noise = white_noise(shape, bw_grain, seed)
    noise = np.array(noise, dtype=F32, copy=True)
    kern = grain_kernel(1 / scale, grain_size_mm, grain_sigma)
    if kern is not None:
        for c in range(noise.shape[-1]):
            noise[..., c] = cv.filter2D(noise[..., c], -1, kern)
    if noise.shape[-1] == 1:
        noise = np.repeat(noise, 3, axis=-1)
    return noise


def grain_factors(density: np.ndarray, grain_curve: np.ndarray) -> np.ndarray:
    """Per-pixel grain amplitude looked up from density (grain.wgsl:77-86)."""
    return multi_channel_interp(density, grain_curve)


def apply_grain(rgb, grain_curve, scale, grain_size_mm=0.01, grain_sigma=0.4, bw_grain=False, noise=None, seed=0):
    field = generate_grain(rgb.shape, scale, grain_size_mm, bw_grain, grain_sigma, noise, seed)
    rgb += field * grain_factors(rgb, grain_curve)           # effects.py:234-235
    return rgb


# --------------------------------------------------------------------------------------
# a8  burn  (effects.py:360-418)  -- PINNED by golden (library calls identical)
# --------------------------------------------------------------------------------------
def burn_mask(green: np.ndarray, d_ref: float, burn_scale: float) -> np.ndarray:
    """Blurred highlight mask of one channel at full resolution (effects.py:360-389, 404-409)."""
    h, w = green.shape
    step = math.ceil(min(h, w) / burn_scale)
    low = cv.resize(green, (w // step, h // step), interpolation=cv.INTER_AREA)
    low = np.clip(low - d_ref, 0, None)
    low = ndimage.gaussian_filter(low, sigma=3, truncate=2)
    up = ndimage.zoom(low, step, order=1)
    up = np.pad(up, [(0, max(h - up.shape[0], 0)), (0, max(w - up.shape[1], 0))], mode="edge")
    return up[:h, :w]


def burn(image: np.ndarray, d_ref, highlight_burn: float, burn_scale: float) -> np.ndarray:
    ref = float(d_ref[1 if len(d_ref) > 1 else 0])       # python float keeps the arithmetic float32
    mask = burn_mask(np.ascontiguousarray(image[..., 1]), ref, float(burn_scale))
    return np.clip(image - float(highlight_burn) * mask[..., None], 0, None)


# --------------------------------------------------------------------------------------
# chroma NR (effects.py:421-561)  -- "next" row, PINNED by tests/golden/chroma_nr.npz
# --------------------------------------------------------------------------------------
def gaussian_kernel_1d(size: int, sigma: float) -> np.ndarray:
    """float32 taps, normalised by their float32 running sum (effects.py:421-435 under numba)."""
    half = size // 2
    taps = np.array([math.exp(-((i - half) * (i - half)) / (2.0 * sigma * sigma)) for i in range(size)], dtype=F32)
    total = F32(0.0)
    for v in taps:
        total = F32(total + v)
    return (taps / total).astype(F32)


def _blur_axis_clamped(plane: np.ndarray, taps: np.ndarray, axis: int) -> np.ndarray:
    """float32 products accumulated in binary64 in tap order, edge-clamped (effects.py:438-482)."""
    half = taps.shape[0] // 2
    n = plane.shape[axis]
    acc = np.zeros(plane.shape, np.float64)
    for i in range(-half, half + 1):
        idx = np.clip(np.arange(n) + i, 0, n - 1)
        acc += (np.take(plane, idx, axis=axis) * taps[i + half]).astype(np.float64)
    return acc.astype(F32)


def chroma_nr_filter(image: np.ndarray, size: int = 0) -> np.ndarray:
    """XYZ -> xyY, Gaussian blur of the two chromaticity planes, back to XYZ (effects.py:497-561)."""
    image = np.asarray(image, dtype=F32)
    X, Y, Z = image[..., 0], image[..., 1], image[..., 2]
    denom = (X + Y) + Z
    ok = denom > 1e-8
    with np.errstate(divide="ignore", invalid="ignore"):
        cx = np.where(ok, X / denom, F32(0)).astype(F32)
        cy = np.where(ok, Y / denom, F32(0)).astype(F32)
    taps_n = int(size) * 2 + 1
    taps = gaussian_kernel_1d(taps_n, 0.3 * ((taps_n - 1) * 0.5 - 1) + 0.8)
    cx = _blur_axis_clamped(_blur_axis_clamped(cx, taps, 1), taps, 0)
    cy = _blur_axis_clamped(_blur_axis_clamped(cy, taps, 1), taps, 0)
    good = cy > 1e-8
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (Y / cy).astype(F32)
        out = np.zeros(image.shape, F32)
        out[..., 0] = np.where(good, cx * inv, F32(0))
        out[..., 1] = np.where(good, Y, F32(0))
        # `1.0 - cx - cy` is binary64 under numba (float64 literal), times the float32 `inv`
        out[..., 2] = np.where(good, ((1.0 - cx.astype(np.float64) - cy.astype(np.float64)) * inv.astype(np.float64)), 0.0)
    return out


# --------------------------------------------------------------------------------------
# canvas (effects.py:290-357)  -- "next" row
# --------------------------------------------------------------------------------------
def get_canvas_data(shape, canvas_mode: str, canvas_scale: float = 1.0, canvas_ratio: float = 1.0):
    colour = (255,) * 3 if "white" in canvas_mode else (0,) * 3 if "black" in canvas_mode else (128,) * 3
    h, w = shape[:2]
    if "Uniform" in canvas_mode:
        border = int(max(h, w) * (canvas_scale - 1))
        res = (h + border, w + border)
    else:
        if "Proportional" in canvas_mode:
            canvas_ratio = w / h
        if w / h > canvas_ratio:
            res = (int(w / canvas_ratio * canvas_scale), int(w * canvas_scale))
        else:
            res = (int(h * canvas_scale), int(h * canvas_ratio * canvas_scale))
    offset = np.subtract(res, (h, w)) // 2
    return res, colour, offset


def add_canvas(image: np.ndarray, canvas_mode: str, canvas_scale: float = 1.0, canvas_ratio: float = 1.0):
    if canvas_mode == "No":
        return image
    res, colour, off = get_canvas_data(image.shape, canvas_mode, canvas_scale, canvas_ratio)
    canvas = np.empty((res[0], res[1], 3), np.uint8)
    canvas[:] = np.asarray(colour, np.uint8)
    canvas[off[0]:off[0] + image.shape[0], off[1]:off[1] + image.shape[1]] = image
    return canvas

# ---- auto exposure (reference color_processing.py:71-99; applied at raw_conversion.py:51-53) ----------
def calc_exposure(rgb: np.ndarray, ref_exposure: float = 0.18, metadata: dict | None = None) -> float:
    """Exposure compensation in stops: power mean (exponent 1/factor) of the green samples of every second
    row and column, against 18 % grey.  `factor` is 3 without EXIF, else sqrt(N^2 / ISO / t) + 1 with N
    defaulting to f/4 (color_processing.py:78-92).  Same NumPy expressions, same float32 array arithmetic."""
    lum_mat = rgb[::2, ::2, 1]
    factor = 3
    if metadata is not None:
        if "EXIF:FNumber" in metadata and metadata["EXIF:FNumber"] and metadata["EXIF:FNumber"] != "undef":
            factor = metadata["EXIF:FNumber"] ** 2 / metadata["EXIF:ISO"] / metadata["EXIF:ExposureTime"]
        else:
            factor = 4 ** 2 / metadata["EXIF:ISO"] / metadata["EXIF:ExposureTime"]
        factor = math.sqrt(factor) + 1
    log_lum = lum_mat ** (1 / factor)
    average_exposure = log_lum.mean() ** factor
    return math.log2(ref_exposure / average_exposure)



# --------------------------------------------------------------------------------------
# a1  the whole hot path in the reference's order (cpu_processor.py:363-407)
# --------------------------------------------------------------------------------------
def render(xyz: np.ndarray, luts: dict, *, frame_width=36, frame_height=24, halation_on=True,
           halation_size=1.0, halation_green_factor=0.4, halation_intensity=1.0, bw=False,
           sharpness=True, sharpening_strength=0.0, sharpening_sigma=1.0,
           grain=2, grain_size=6.0, grain_sigma=0.4, noise=None, seed=0,
           highlight_burn=0.0, burn_scale=50.0, canvas_mode="No", canvas_scale=1.0, canvas_ratio=1.0,
           eps=LOG_CLIP_EPS, stages: dict | None = None) -> np.ndarray:
    """`luts`: {"lut2d": (n,n,3), "curve": (4,N), "lut3d": (n,n,n,3), "mtf": [(logf, vals)]*3 | None,
    "grain_curve": (4,N) | None, "d_ref": (3,)}.  If `stages` is a dict it receives float32 copies of
    the intermediate images ("exposure", "halation", "density", "mtf", "grain", "burn", "rgb")."""

    def tap(name, img):
        if stages is not None:
            stages[name] = np.array(img, dtype=F32, copy=True)

    image = apply_2d_lut(xyz, luts["lut2d"])                                  # :364
    tap("exposure", image)
    scale = max(image.shape) / max(frame_width, frame_height)                 # :366
    if halation_on:                                                           # :368-376
        image = halation(image, scale, halation_size, halation_green_factor, halation_intensity, bw)
        tap("halation", image)
    log_clip(image, eps)                                                      # :378
    image = multi_channel_interp(image, luts["curve"])                        # :380
    tap("density", image)
    if sharpness and luts.get("mtf") is not None:                             # :382-385
        image = film_sharpness(image, luts["mtf"], scale, sharpening_strength, sharpening_sigma)
        tap("mtf", image)
    if grain and luts.get("grain_curve") is not None:                         # :387-397
        image = apply_grain(image, luts["grain_curve"], scale, grain_size / 1000, grain_sigma,
                            bw_grain=(grain == 1), noise=noise, seed=seed)
        image = clip_min0(image)
        tap("grain", image)
    if highlight_burn:                                                        # :399-403
        image = burn(image, luts["d_ref"], highlight_burn, burn_scale).astype(F32)
        tap("burn", image)
    image = apply_lut_tetrahedral(image, luts["lut3d"], 0.25)                 # :405
    tap("rgb", image)
    out = quantise_u8(image)                                                  # :407
    return add_canvas(out, canvas_mode, canvas_scale, canvas_ratio)           # :409


# --------------------------------------------------------------------------------------
# presentation blit (shaders/copy_to_int.wgsl:18-51)  -- "next" row, PARITY UNPINNED: the
# reference samples with a hardware linear sampler whose weight precision is implementation-defined;
# restated with exact float32 bilinear weights, texel centres at +0.5, clamp to edge, round to nearest.
# --------------------------------------------------------------------------------------
def present(image: np.ndarray, dst_hw, transform, colour=(255, 255, 255)) -> np.ndarray:
    h, w = image.shape[:2]
    dh, dw = dst_hw
    sx, sy, ox, oy, cx0, cy0, cx1, cy1 = (F32(v) for v in transform)
    dx = (np.arange(dw, dtype=F32) + F32(0.5))[None, :]
    dy = (np.arange(dh, dtype=F32) + F32(0.5))[:, None]
    su = ((dx - ox) * sx).astype(F32) + np.zeros((dh, 1), F32)
    sv = ((dy - oy) * sy).astype(F32) + np.zeros((1, dw), F32)
    inside = (su >= 0) & (su <= 1) & (sv >= 0) & (sv <= 1)
    fx = (su * F32(w) - F32(0.5)).astype(F32)
    fy = (sv * F32(h) - F32(0.5)).astype(F32)
    x0f, y0f = np.floor(fx), np.floor(fy)
    ax, ay = (fx - x0f).astype(F32)[..., None], (fy - y0f).astype(F32)[..., None]
    x0 = np.clip(x0f.astype(np.int64), 0, w - 1)
    x1 = np.clip(x0f.astype(np.int64) + 1, 0, w - 1)
    y0 = np.clip(y0f.astype(np.int64), 0, h - 1)
    y1 = np.clip(y0f.astype(np.int64) + 1, 0, h - 1)
    img = image.astype(F32)
    one = F32(1.0)
    top = (img[y0, x0] * (one - ax) + img[y0, x1] * ax).astype(F32)
    bot = (img[y1, x0] * (one - ax) + img[y1, x1] * ax).astype(F32)
    val = np.clip(np.rint((top * (one - ay) + bot * ay).astype(F32)), 0, 255).astype(np.uint8)
    out = np.zeros((dh, dw, 4), np.uint8)
    canvas = (~inside) & (dx >= cx0) & (dx <= cx1) & (dy >= cy0) & (dy <= cy1)
    out[canvas] = (*colour, 255)
    out[inside, :3] = val[inside]
    out[inside, 3] = 255
    return out
