"""Mint golden vectors from the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (imported through oracle/ref_loader.py with third-party stubs).
Outputs small .npz fixtures next to this file; they are committed and are what pins the
oracle (tests/test_oracle_golden.py) on machines where the reference tree is absent.

Every array named `ref_*` was computed by a reference function:
  raw2film.utils.apply_lut_tetrahedral      (utils.py:247-380)
  raw2film.effects.compute_halation_kernel  (effects.py:239-263)
  raw2film.effects.mtf_kernel               (effects.py:165-185)
  raw2film.effects.convolve_2d              (effects.py:146-156)
  raw2film.effects.burn                     (effects.py:392-418)
  raw2film.effects.add_canvas               (effects.py:338-357)
  raw2film.utils.resolution_scaling         (utils.py:226-244)
  raw2film.effects.chroma_nr_filter         (effects.py:421-561)
  raw2film.utils.generate_histogram         (utils.py:145-223)
  raw2film.color_processing.calc_exposure   (color_processing.py:71-99)   [`make_golden.py calc_exposure`]
  raw2film.utils.resolution_scaling         (utils.py:226-244), float32 + uint8 battery [`make_golden.py resize`]
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_loader import load_reference  # noqa: E402

MTF_POLYLINES = [
    # (log1p(cycles/mm), response) per channel -- synthetic stand-in stock data
    (np.log1p([0.0, 5.0, 10.0, 20.0, 50.0, 100.0, 200.0]), [1.0, 1.03, 1.05, 0.92, 0.55, 0.22, 0.05]),
    (np.log1p([0.0, 5.0, 10.0, 20.0, 50.0, 100.0, 200.0]), [1.0, 1.04, 1.08, 0.98, 0.62, 0.27, 0.06]),
    (np.log1p([0.0, 5.0, 10.0, 20.0, 50.0, 100.0, 200.0]), [1.0, 1.02, 1.03, 0.85, 0.45, 0.15, 0.03]),
]


class _Stock:
    """Hashable stand-in exposing what effects.mtf_kernel / effects.burn read."""

    def __init__(self, mtf, d_ref):
        self.mtf = mtf
        self.d_ref = d_ref

    def __hash__(self):
        return id(self)


def main():
    effects, utils = load_reference()
    rng = np.random.default_rng(20261017)

    # ---- tetrahedral LUT -------------------------------------------------------------
    tet = {}
    for n in (9, 17):
        lut = rng.random((n, n, n, 3), dtype=np.float32)
        img = (rng.random((48, 64, 3), dtype=np.float32) * 4.4).astype(np.float32)
        img[0, :8] = 0.0
        img[1, :8] = 4.0          # exactly the top of the table -> clamp branch
        img[2, :8] = 5.0          # above the table
        img[3, :8, 0] = 4.0 * np.arange(8) / (n - 1)   # exactly on lattice planes
        img[4, :8] = np.float32(1e-7)
        img[5, :8, 1] = img[5, :8, 0]                   # ties dr == dg
        img[6, :8, 2] = img[6, :8, 1]                   # ties dg == db
        img[7, :8] = img[7, :8, :1]                     # grey axis
        tet[f"lut{n}"] = lut
        tet[f"img{n}"] = img
        tet[f"ref_out{n}"] = utils.apply_lut_tetrahedral(img, lut, 0.25)
        tet[f"ref_out{n}_s1"] = utils.apply_lut_tetrahedral((img / 4.4).astype(np.float32), lut, 1.0)
    np.savez_compressed(os.path.join(HERE, "tetra.npz"), **tet)

    # ---- halation kernels --------------------------------------------------------------
    hal = {}
    cases = [(6000 / 36, 1.0, 0.4, 1.0, False), (6000 / 36, 1.0, 0.3, 1.0, False),
             (9504 / 36, 2.0, 0.3, 1.0, False), (1920 / 36, 1.0, 0.3, 1.0, False),
             (20.0, 1.0, 0.4, 1.0, False), (50.0, 1.5, 0.3, 0.7, True), (33.3, 0.7, 0.25, 2.0, False)]
    hal["cases"] = np.array([[c[0], c[1], c[2], c[3], float(c[4])] for c in cases], np.float64)
    for i, (scale, size, gf, inten, bw) in enumerate(cases):
        hal[f"ref_kernel{i}"] = effects.compute_halation_kernel(
            scale, halation_size=size, halation_green_factor=gf, halation_intensity=inten, bw=bw)
    np.savez_compressed(os.path.join(HERE, "halation_kernels.npz"), **hal)

    # ---- MTF kernels -------------------------------------------------------------------
    mtf = {"logf": np.stack([np.asarray(p[0]) for p in MTF_POLYLINES]),
           "vals": np.stack([np.asarray(p[1], np.float64) for p in MTF_POLYLINES])}
    stock = _Stock([(tuple(lf), tuple(vs)) for lf, vs in MTF_POLYLINES], (0.5, 0.6, 0.7))
    mcases = [(6000 / 36, 0.0, 1.0), (9504 / 36, 0.0, 1.0), (1920 / 36, 0.0, 1.0), (6000 / 36, 0.5, 1.0),
              (100.0, 1.0, 2.0), (400.0, 0.0, 1.0)]
    mtf["cases"] = np.array(mcases, np.float64)
    for i, (scale, strength, sigma) in enumerate(mcases):
        mtf[f"ref_kernel{i}"] = np.array(effects.mtf_kernel(stock, scale, strength, sigma), copy=True)
    np.savez_compressed(os.path.join(HERE, "mtf_kernels.npz"), **mtf)

    # ---- convolve_2d (orientation + border) ----------------------------------------------
    conv = {}
    img = rng.random((40, 56, 3), dtype=np.float32)
    k_small = rng.random((7, 7, 3), dtype=np.float32)           # asymmetric: proves correlation
    k_small /= k_small.sum(axis=(0, 1), keepdims=True)
    k_big = rng.random((15, 15, 3), dtype=np.float32)           # > 11x11: cv2 DFT path
    k_big /= k_big.sum(axis=(0, 1), keepdims=True)
    conv["img"], conv["k_small"], conv["k_big"] = img, k_small, k_big
    conv["ref_small"] = effects.convolve_2d(img.copy(), k_small)
    conv["ref_big"] = effects.convolve_2d(img.copy(), k_big)
    np.savez_compressed(os.path.join(HERE, "convolve.npz"), **conv)

    # ---- highlight burn ------------------------------------------------------------------
    b = {}
    dens = (rng.random((72, 108, 3), dtype=np.float32) * 3.0).astype(np.float32)
    dens = np.ascontiguousarray(dens)
    b["density"] = dens
    b["params"] = np.array([0.5, 10.0])
    b["d_ref"] = np.array(stock.d_ref)
    b["ref_out"] = effects.burn(dens.copy(), stock, 0.5, 10.0).astype(np.float32)
    dens2 = np.ascontiguousarray((rng.random((57, 83, 3), dtype=np.float32) * 3.0).astype(np.float32))
    b["density2"] = dens2
    b["params2"] = np.array([0.8, 7.0])
    b["ref_out2"] = effects.burn(dens2.copy(), stock, 0.8, 7.0).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "burn.npz"), **b)

    # ---- canvas + resize -----------------------------------------------------------------
    cv_ = {}
    im8 = rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)
    cv_["img"] = im8
    modes = ["Proportional white", "Proportional black", "Uniform white", "Uniform black", "Fixed white",
             "Fixed black"]
    for i, m in enumerate(modes):
        cv_[f"ref_{i}"] = effects.add_canvas(im8, m, 1.2, 0.8)
    im8b = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    cv_["img_resize"] = im8b
    cv_["ref_down"] = utils.resolution_scaling(im8b, (32, 32))
    cv_["ref_up"] = utils.resolution_scaling(im8b, (128, 400))
    np.savez_compressed(os.path.join(HERE, "canvas_resize.npz"), **cv_)
    # ---- chroma NR (effects.py:421-561) ---------------------------------------------------
    cn = {}
    xyz = (rng.random((61, 83, 3), dtype=np.float32) * np.float32(1.5)).astype(np.float32)
    xyz[0, :4] = 0.0                      # denom <= eps branch
    xyz[1, :4, 1] = 0.0                   # y chromaticity 0 -> xyY_to_XYZ zero branch
    cn["xyz"] = xyz
    for size in (1, 3, 8):
        cn[f"ref_out{size}"] = np.asarray(effects.chroma_nr_filter(xyz.copy(), size))
        sz = int(size) * 2 + 1
        cn[f"ref_kernel{size}"] = np.asarray(effects.gaussian_kernel_1d(sz, 0.3 * ((sz - 1) * 0.5 - 1) + 0.8))
    np.savez_compressed(os.path.join(HERE, "chroma_nr.npz"), **cn)
    # ---- RGB histogram widget (utils.py:145-223) --------------------------------------------------
    hg = {}
    img = rng.integers(0, 256, (90, 130, 3), dtype=np.uint8)
    img[:30] = (img[:30] // 4 + 100).astype(np.uint8)          # a peak, so the log/height scaling matters
    mix = rng.integers(0, 256, (2, 2, 2, 4), dtype=np.uint8)
    mix[0, 0, 0] = 0
    hg["img"], hg["mix"] = img, mix
    hg["ref_hist100"] = np.asarray(utils.generate_histogram(img, mix, 100))
    hg["ref_hist64"] = np.asarray(utils.generate_histogram(img, mix, 64))
    np.savez_compressed(os.path.join(HERE, "histogram.npz"), **hg)
    print("golden vectors written to", HERE)


def make_calc_exposure():
    """calc_exposure (color_processing.py:71-99) on frames ingested like raw_conversion.py:51
    (uint16 / 65535 in float32), odd and even sizes, with and without EXIF metadata."""
    load_reference()
    from raw2film import color_processing  # noqa: E402

    rng = np.random.default_rng(7151)
    out = {}
    metas = [None, {"EXIF:FNumber": 2.8, "EXIF:ISO": 400, "EXIF:ExposureTime": 1 / 250},
             {"EXIF:FNumber": "undef", "EXIF:ISO": 100, "EXIF:ExposureTime": 1 / 30}]
    out["meta_fnumber"] = np.array([0.0, 2.8, -1.0])       # 0: no metadata, -1: "undef"
    out["meta_iso"] = np.array([0.0, 400.0, 100.0])
    out["meta_time"] = np.array([0.0, 1 / 250, 1 / 30])
    for i, shape in enumerate([(97, 131), (64, 96), (101, 150)]):
        lum = np.exp(rng.normal(-2.5, 1.2, shape)).clip(0, 1)
        u16 = (rng.random((*shape, 3)) * 0.5 + 0.5) * lum[..., None] * 65535
        u16 = u16.astype(np.uint16)
        u16[:3, :5] = 0                                                 # zeros: 0 ** (1/factor) = 0
        u16[5, :7] = 65535
        out[f"u16_{i}"] = u16
        rgb = u16.astype(np.float32) / 65535.0
        out[f"ref_exp_{i}"] = np.array([color_processing.calc_exposure(rgb, metadata=m) for m in metas], np.float64)
    np.savez_compressed(os.path.join(HERE, "calc_exposure.npz"), **out)
    print("calc_exposure.npz written:", {k: v for k, v in out.items() if k.startswith("ref_")})


def make_resize():
    """resolution_scaling (utils.py:226-244) -- i.e. cv2.resize INTER_AREA / INTER_LANCZOS4 -- on float32 frames
    (pre-path use, cpu_processor.py:122-134) and uint8 images (post-path use, :411-412): integer and non-integer
    shrink factors, enlargements, 3- and 4-channel float frames."""
    _, utils = load_reference()
    rng = np.random.default_rng(41017)
    out = {}
    f = (rng.random((72, 96, 3), dtype=np.float32) * np.float32(2.5)).astype(np.float32)
    f[5:9, 7:12] = 40.0                                   # a highlight: large dynamic range
    u = rng.integers(0, 256, (72, 96, 3), dtype=np.uint8)
    out["f32"], out["u8"] = f, u
    # (rows, cols) boxes handed to resolution_scaling; the aspect-preserving fit decides the real size
    boxes = {"half": (36, 48), "third": (24, 32), "quarter": (18, 24), "sixth": (12, 16), "odd": (29, 41),
             "near": (71, 95), "tiny": (5, 9), "tall": (50, 400), "up15": (108, 144), "up2": (144, 192),
             "up_odd": (173, 240), "up_one": (73, 98)}
    for name, box in boxes.items():
        out[f"box_{name}"] = np.asarray(box)
        out[f"ref_f32_{name}"] = utils.resolution_scaling(f.copy(), box)
        out[f"ref_u8_{name}"] = utils.resolution_scaling(u.copy(), box)
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **out)
    print("resize.npz written:", {k: v.shape for k, v in out.items() if k.startswith("ref_f32")})


def make_third_party():
    """Known-answer vectors of the third-party stages from the REAL spectral_film_lut (only when it is installed):
    the day this runs, tests/test_oracle_golden.py::test_third_party_goldens pins the WGSL restatements."""
    from oracle import third_party

    fns = third_party.load()
    if not fns:
        print("spectral_film_lut is not installed: nothing minted (the stages stay PARITY UNPINNED)")
        return
    rng = np.random.default_rng(808)
    out = {}
    xyz = (rng.random((64, 96, 3), dtype=np.float32) * np.float32(2.0)).astype(np.float32)
    xyz[0, :4] = 0.0
    lut2d = rng.random((64, 64, 3), dtype=np.float32)
    out["xyz"], out["lut2d"] = xyz, lut2d
    if "apply_2d_lut" in fns:
        out["ref_apply_2d_lut"] = np.asarray(fns["apply_2d_lut"](xyz.copy(), lut2d))
    expo = (rng.random((64, 96, 3), dtype=np.float32) * np.float32(4.0)).astype(np.float32)
    expo[1, :4] = 0.0
    out["exposure"] = expo
    if "log_clip" in fns:
        img = expo.copy()
        res = fns["log_clip"](img)
        out["ref_log_clip"] = np.asarray(img if res is None else res)
    curve = np.stack([np.linspace(-4, 2, 256)] + [np.cumsum(rng.random(256) * 0.02) for _ in range(3)]).astype(np.float32)
    warped = curve.copy()
    warped[0] = (-1 + 3 * (0.35 * np.linspace(-1, 1, 256) + 0.65 * np.linspace(-1, 1, 256) ** 3)).astype(np.float32)
    logs = (rng.random((64, 96, 3), dtype=np.float32) * np.float32(7.0) - np.float32(4.5)).astype(np.float32)
    out["curve"], out["curve_warped"], out["logs"] = curve, warped, logs
    if "multi_channel_interp" in fns:
        out["ref_interp"] = np.asarray(fns["multi_channel_interp"](logs.copy(), curve))
        out["ref_interp_warped"] = np.asarray(fns["multi_channel_interp"](logs.copy(), warped))
    if "grain_kernel" in fns:
        for i, (px, size, sigma) in enumerate([(1 / 166.67, 0.006, 0.4), (1 / 264.0, 0.006, 0.4), (1 / 53.3, 0.01, 0.3)]):
            k = fns["grain_kernel"](px, grain_size_mm=size, grain_sigma=sigma)
            out[f"grain_args_{i}"] = np.array([px, size, sigma])
            out[f"ref_grain_kernel_{i}"] = np.zeros((0, 0), np.float32) if k is None else np.asarray(k, np.float32)
    np.savez_compressed(os.path.join(HERE, "third_party.npz"), **out)
    print("third_party.npz written:", sorted(k for k in out if k.startswith("ref_")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "third_party":
        make_third_party()
    elif len(sys.argv) > 1 and sys.argv[1] == "calc_exposure":
        make_calc_exposure()
    elif len(sys.argv) > 1 and sys.argv[1] == "resize":
        make_resize()
    else:
        main()
