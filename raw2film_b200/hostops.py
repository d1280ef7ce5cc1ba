"""Host-side steps that sit on either side of the render path and that BOTH reference processors
also run on the CPU (they belong to the reference's "CPU phase", not to the per-pixel hot path):

* `resolution_scaling` -- reference utils.py:226-244: cv2 INTER_AREA when shrinking, INTER_LANCZOS4
  when enlarging, aspect-preserving fit into `resolution` = (rows, cols).  Used before the path
  (cpu_processor.py:127-134, gpu_processor.py:751-758) and after it (cpu_processor.py:411-412).
* `canvas_geometry` -- reference effects.py:290-335 (`get_canvas_data`): canvas size, colour and
  paste offset; the paste itself runs on the device (r2f_canvas_paste).
"""
from __future__ import annotations

import cv2 as cv
import numpy as np


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
