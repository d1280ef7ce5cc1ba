"""CPU restatement of cv2.resize as the reference calls it (TEST INFRASTRUCTURE ONLY, like film_oracle.py).

reference: src/raw2film/utils.py:226-244 `resolution_scaling` -- cv.resize(INTER_AREA) when shrinking,
cv.resize(INTER_LANCZOS4) when enlarging; used before the path on the float32 frame
(cpu_processor.py:122-134, gpu_processor.py:748-758) and after it on the uint8 image (cpu_processor.py:411-412).

cv2 (OpenCV 4.13, imgproc/src/resize.cpp) is a third-party wheel without sources in the reference tree, so the
algorithms are restated from OpenCV's published implementation and PINNED against cv2 itself on this machine
(tests/golden/resize.npz, minted by tests/golden/make_golden.py; tests/test_oracle_golden.py):

* INTER_AREA, non-integer scale (`resizeArea_` + `computeResizeAreaTab`): per destination cell a list of
  (source index, weight) taps, weights float32; horizontal sums `buf += src * alpha` in tap order, vertical
  `sum = beta * buf` / `sum += beta * buf` in row order, all float32, separate multiply and add.  uint8 sources
  accumulate in float32 too and round half to even at the end.  Bit-exact against cv2.
* INTER_AREA, integer scale (`resizeAreaFast_`): float32: block sum in row-major order, four terms at a time
  (`sum += ((s0 + s1) + s2) + s3`), times float32(1/area); uint8: integer block sum, `sum * float32(1/area)` rounded
  half to even, except 2x2 which is `(a + b + c + d + 2) >> 2`.  Bit-exact against cv2.
* INTER_LANCZOS4 (`interpolateLanczos4`, `HResizeLanczos4`, `VResizeLanczos4`): 8 taps per axis, edge-replicated
  source indices.  uint8: coefficients rounded to 1/2048 fixed point, exact integer sums, `(sum + 2^21) >> 22`,
  saturated: bit-exact against cv2.  float32: left-to-right float32 sums; cv2's vertical pass is a SIMD kernel
  whose fused-multiply-add use and tail handling depend on the host CPU's vector ISA, so float32 agrees to a few
  ulp only (documented tolerance 1e-6 of the data range).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def target_size(shape, resolution):
    """(rows, cols) after resolution_scaling(image, resolution), or None when the factor is exactly 1."""
    rows, cols = shape[:2]
    factor = min(resolution[0] / rows, resolution[1] / cols)
    if factor == 1:
        return None
    return round(rows * factor), round(cols * factor)


# ---- INTER_AREA -----------------------------------------------------------------------------------------
def area_tab(ssize: int, dsize: int):
    """computeResizeAreaTab: lists (dst index, src index, float32 weight) in cv2's order."""
    scale = 1.0 / (dsize / ssize)
    di, si, al = [], [], []
    for dx in range(dsize):
        f1 = dx * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = math.ceil(f1), math.floor(f2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            di.append(dx); si.append(s1 - 1); al.append((s1 - f1) / cell)
        for sx in range(s1, s2):
            di.append(dx); si.append(sx); al.append(1.0 / cell)
        if f2 - s2 > 1e-3:
            di.append(dx); si.append(s2); al.append(min(min(f2 - s2, 1.0), cell) / cell)
    return np.asarray(di, np.int64), np.asarray(si, np.int64), np.asarray(al, F32)


def _tab_slots(di, si, al, dsize):
    """Per destination index: taps padded to the longest list (weight 0 = no tap)."""
    start = np.searchsorted(di, np.arange(dsize))
    count = np.searchsorted(di, np.arange(dsize), side="right") - start
    m = int(count.max())
    idx = np.zeros((dsize, m), np.int64)
    wgt = np.zeros((dsize, m), F32)
    for k in range(m):
        ok = count > k
        idx[ok, k] = si[start[ok] + k]
        wgt[ok, k] = al[start[ok] + k]
    return idx, wgt, count


def is_area_fast(ssize, dsize):
    scale = 1.0 / (dsize / ssize)
    return abs(scale - int(scale)) < np.finfo(np.float64).eps, int(scale)


def resize_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    fx, ix = is_area_fast(sw, dw)
    fy, iy = is_area_fast(sh, dh)
    if fx and fy:
        return _resize_area_fast(src, dw, dh, ix, iy)
    s = src.astype(F32)
    xi, xw, xc = _tab_slots(*area_tab(sw, dw), dw)
    yi, yw, yc = _tab_slots(*area_tab(sh, dh), dh)
    buf = np.zeros((sh, dw, s.shape[2]), F32)
    for k in range(xi.shape[1]):
        on = (xc > k)[None, :, None]
        buf = np.where(on, buf + s[:, xi[:, k], :] * xw[:, k][None, :, None], buf).astype(F32)
    out = np.zeros((dh, dw, s.shape[2]), F32)
    for k in range(yi.shape[1]):
        term = (yw[:, k][:, None, None] * buf[yi[:, k]]).astype(F32)
        on = (yc > k)[:, None, None]
        out = np.where(on, term if k == 0 else out + term, out).astype(F32)
    if src.dtype == np.uint8:
        return np.clip(np.rint(out), 0, 255).astype(np.uint8)
    return out


def _resize_area_fast(src, dw, dh, ix, iy):
    area = ix * iy
    scale = F32(1.0 / area)
    blocks = [src[j:dh * iy:iy, i:dw * ix:ix] for j in range(iy) for i in range(ix)]   # ofs order: row-major
    if src.dtype == np.uint8:
        total = sum(b.astype(np.int64) for b in blocks)
        if ix == 2 and iy == 2:
            return ((total + 2) >> 2).astype(np.uint8)
        return np.clip(np.rint(total.astype(F32) * scale), 0, 255).astype(np.uint8)
    total = np.zeros(blocks[0].shape, F32)
    k = 0
    while k + 4 <= area:
        total = (total + (((blocks[k] + blocks[k + 1]) + blocks[k + 2]) + blocks[k + 3])).astype(F32)
        k += 4
    while k < area:
        total = (total + blocks[k]).astype(F32)
        k += 1
    return (total * scale).astype(F32)


# ---- INTER_LANCZOS4 --------------------------------------------------------------------------------------
_S45 = 0.70710678118654752440084436210485
_CS = ((1, 0), (-_S45, -_S45), (0, 1), (_S45, -_S45), (-1, 0), (_S45, _S45), (0, -1), (-_S45, _S45))


def lanczos4_coeffs(x) -> np.ndarray:
    """interpolateLanczos4: eight float32 weights for the fractional offset x in [0, 1)."""
    x = F32(x)
    y0 = -(float(x) + 3) * math.pi * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    co = np.zeros(8, F32)
    total = F32(0)
    for i in range(8):
        d = F32(F32(x + F32(3)) - F32(i))
        if abs(d) >= F32(1e-6):
            y = -float(d) * math.pi * 0.25
            co[i] = F32((_CS[i][0] * s0 + _CS[i][1] * c0) / (y * y))
        else:
            co[i] = F32(1e30)
        total = F32(total + co[i])
    return (co * F32(F32(1.0) / total)).astype(F32)


def lanczos4_tab(ssize: int, dsize: int):
    """Per destination index: first source index - 3 .. + 4 (clamped by the caller) and the eight weights."""
    scale = 1.0 / (dsize / ssize)
    ofs = np.zeros(dsize, np.int64)
    al = np.zeros((dsize, 8), F32)
    for d in range(dsize):
        f = F32((d + 0.5) * scale - 0.5)
        s = math.floor(f)
        ofs[d] = s
        al[d] = lanczos4_coeffs(F32(f - F32(s)))
    return ofs, al


def resize_lanczos4(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw, cn = src.shape
    xo, xa = lanczos4_tab(sw, dw)
    yo, ya = lanczos4_tab(sh, dh)
    xidx = np.clip(xo[:, None] + np.arange(-3, 5)[None, :], 0, sw - 1)
    yidx = np.clip(yo[:, None] + np.arange(-3, 5)[None, :], 0, sh - 1)
    if src.dtype == np.uint8:
        ia = np.clip(np.rint(xa * F32(2048)), -32768, 32767).astype(np.int64)
        ib = np.clip(np.rint(ya * F32(2048)), -32768, 32767).astype(np.int64)
        s = src.astype(np.int64)
        tmp = sum(s[:, xidx[:, j], :] * ia[:, j][None, :, None] for j in range(8))
        out = sum(tmp[yidx[:, j]] * ib[:, j][:, None, None] for j in range(8))
        return np.clip((out + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    s = src.astype(F32)
    tmp = None
    for j in range(8):
        term = (s[:, xidx[:, j], :] * xa[:, j][None, :, None]).astype(F32)
        tmp = term if tmp is None else (tmp + term).astype(F32)
    out = None
    for j in range(8):
        term = (tmp[yidx[:, j]] * ya[:, j][:, None, None]).astype(F32)
        out = term if out is None else (out + term).astype(F32)
    return out


def resolution_scaling(image: np.ndarray, resolution) -> np.ndarray:
    """utils.py:226-244 without cv2."""
    size = target_size(image.shape, resolution)
    if size is None:
        return image
    rows, cols = size
    factor = min(resolution[0] / image.shape[0], resolution[1] / image.shape[1])
    return resize_area(image, cols, rows) if factor < 1 else resize_lanczos4(image, cols, rows)
