"""Host-side steps that sit on either side of the render path and that BOTH reference processors
also run on the CPU (they belong to the reference's "CPU phase", not to the per-pixel hot path):

* `resolution_scaling` -- reference utils.py:226-244: cv2 INTER_AREA when shrinking, INTER_LANCZOS4
  when enlarging, aspect-preserving fit into `resolution` = (rows, cols).  Used before the path
  (cpu_processor.py:127-134, gpu_processor.py:751-758) and after it (cpu_processor.py:411-412).
* `canvas_geometry` -- reference effects.py:290-335 (`get_canvas_data`): canvas size, colour and
  paste offset; the paste itself runs on the device (r2f_canvas_paste).
* `exposure_factor` -- reference color_processing.py:78-92: the EXIF-dependent exponent of the auto-exposure
  power mean (the reduction itself runs on the device, r2f_calc_exposure).
"""
from __future__ import annotations

import math

import cv2 as cv
import numpy as np


def exposure_factor(metadata: dict | None) -> float:
    """Exponent `factor` of calc_exposure (reference color_processing.py:78-92): 3 without metadata, else
    sqrt(N^2 / ISO / t) + 1 with f-number N (f/4 when EXIF has none), ISO and exposure time t."""
    factor = 3
    if metadata is not None:
        fnum = metadata.get("EXIF:FNumber")
        if "EXIF:FNumber" in metadata and fnum and fnum != "undef":
            factor = fnum ** 2 / metadata["EXIF:ISO"] / metadata["EXIF:ExposureTime"]
        else:
            factor = 4 ** 2 / metadata["EXIF:ISO"] / metadata["EXIF:ExposureTime"]
        factor = math.sqrt(factor) + 1
    return float(factor)


def target_size(shape, resolution):
    """(rows, cols) that `resolution_scaling` (reference utils.py:226-244) resizes an image of `shape` to, or None
    when it leaves the image alone (factor exactly 1)."""
    rows, cols = shape[:2]
    factor = min(resolution[0] / rows, resolution[1] / cols)
    if factor == 1:
        return None
    return round(rows * factor), round(cols * factor)


def resolution_scaling(image: np.ndarray, resolution) -> np.ndarray:
    rows, cols = image.shape[:2]
    factor = min(resolution[0] / rows, resolution[1] / cols)
    if factor == 1:
        return image
    size = (round(cols * factor), round(rows * factor))          # cv2 takes (width, height)
    return cv.resize(image, size, interpolation=cv.INTER_AREA if factor < 1 else cv.INTER_LANCZOS4)


def canvas_geometry(shape, canvas_mode: str, canvas_scale: float = 1.0, canvas_ratio: float = 1.0):
    """-> ((canvas_rows, canvas_cols), (r, g, b), (row_offset, col_offset))."""
    rows, cols = shape[:2]
    if "white" in canvas_mode:
        colour = (255, 255, 255)
    elif "black" in canvas_mode:
        colour = (0, 0, 0)
    else:
        colour = (128, 128, 128)
    if "Uniform" in canvas_mode:
        border = int(max(rows, cols) * (canvas_scale - 1))
        size = (rows + border, cols + border)
    elif "Proportional" in canvas_mode or "Fixed" in canvas_mode:
        ratio = cols / rows if "Proportional" in canvas_mode else canvas_ratio
        if cols / rows > ratio:
            size = (int(cols / ratio * canvas_scale), int(cols * canvas_scale))
        else:
            size = (int(rows * canvas_scale), int(rows * ratio * canvas_scale))
    else:
        raise ValueError(f"unknown canvas mode {canvas_mode!r}")
    offset = ((size[0] - rows) // 2, (size[1] - cols) // 2)
    return size, colour, offset


def present_geometry(src_size, dst_size, pipeline_resolution=None, output_resolution=None, canvas_resolution=None):
    """The uniform block of the presentation blit, restating `_bind_copy_to_dst` (reference
    gpu_processor.py:1416-1512): (scale_x, scale_y, offset_x, offset_y, canvas_min_x, canvas_min_y, canvas_max_x,
    canvas_max_y).  All sizes are (width, height) like the reference's attributes."""
    src_w, src_h = src_size
    dst_w, dst_h = dst_size
    has_canvas = canvas_resolution is not None and canvas_resolution[0] > 0
    if has_canvas:
        content_w, content_h = canvas_resolution
    elif output_resolution is not None and output_resolution[0] > 0:
        content_w, content_h = output_resolution
    else:
        content_w, content_h = pipeline_resolution if pipeline_resolution is not None else (src_w, src_h)
    src_aspect, dst_aspect = content_w / content_h, dst_w / dst_h
    if src_aspect > dst_aspect:
        cw, ch, cx, cy = dst_w, dst_w / src_aspect, 0.0, (dst_h - dst_w / src_aspect) / 2.0
    else:
        cw, ch, cx, cy = dst_h * src_aspect, dst_h, (dst_w - dst_h * src_aspect) / 2.0, 0.0
    canvas = (cx, cy, cx + cw, cy + ch) if has_canvas else (0.0, 0.0, 0.0, 0.0)
    if canvas_resolution is not None and output_resolution is not None:
        target_w, target_h = output_resolution
        rw, rh = cw * (target_w / content_w), ch * (target_h / content_h)
        ox, oy = cx + (cw - rw) / 2.0, cy + (ch - rh) / 2.0
    else:
        rw, rh, ox, oy = cw, ch, cx, cy
    return (1.0 / rw, 1.0 / rh, ox, oy) + canvas


def histogram_image(counts: np.ndarray, mix_table: np.ndarray, height: int = 100) -> np.ndarray:
    """RGB histogram widget image (height, 256, 4) uint8 from per-channel 256-bin counts.

    reference utils.py:171-223 (everything after the counting loop, which runs on the device):
    float32 log1p(count / max), 3-tap moving average with replicated ends, scaling to `height`,
    truncation to int, and the (2,2,2,4) colour mix table lookup per column / row.
    """
    f = counts.astype(np.float32)                                   # (3, 256)
    peak = np.float32(max(f.max(), 0))
    if peak == 0:
        peak = np.float32(1)
    f = np.log1p(f / peak).astype(np.float32)
    left = np.concatenate([f[:, :1], f[:, :-1]], axis=1)
    right = np.concatenate([f[:, 1:], f[:, -1:]], axis=1)
    smooth = ((left + f + right) / np.float32(3)).astype(np.float32)
    top = np.float32(max(smooth.max(), 0))
    if top == 0:
        top = np.float32(1)
    heights = ((smooth * np.float32(height)) / top).astype(np.int32)   # (3, 256)
    rows = np.arange(height)[:, None]                                   # y
    active = rows >= (height - heights[:, None, :])                     # (3, height, 256)
    return mix_table[active[0].astype(np.intp), active[1].astype(np.intp), active[2].astype(np.intp)]
