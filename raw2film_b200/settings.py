"""Pure host logic shared by the processor, the batch exporter and the tests (no CUDA needed):
the reference's settings defaults, stage gating and frame sharding."""
from __future__ import annotations

from . import flags as F

# signature defaults of CpuProcessor.process (reference cpu_processor.py:269-322)
DEFAULTS = dict(
    lens_correction=True, print_film=None, exp_comp=0.0, red_light=0.0, green_light=0.0, blue_light=0.0,
    projector_kelvin=6500, shadow_comp=0.0, sat_adjust=1.0, gamma_func="sRGB", exp_kelvin=6500, tint=0.0,
    inversion_gamma=4.0, idealized_curve=False, inversion=False, push_pull=0.0, white_balance=False,
    white_clip=False, icc_transform=None, resolution=None, frame_width=36, frame_height=24, rotation=0.0,
    zoom=1.0, rotate_times=0, flip=False, cam=None, lens=None, canvas_mode="No", canvas_scale=1.0,
    canvas_ratio=1.0, halation_intensity=1.0, halation=True, halation_size=1.0, halation_green_factor=0.4,
    sharpness=True, sharpening_strength=0.0, sharpening_sigma=1.0, chroma_nr=0, grain=2, highlight_burn=0.0,
    burn_scale=50.0, half_size=True, cache=True, color_masking=None, max_scale=400.0)

# the GUI's profile / image defaults (reference gui.py:486-531), i.e. what a default export passes
GUI_PROFILE_DEFAULTS = dict(
    red_light=0, green_light=0, blue_light=0, halation=True, sharpness=True, grain=2, film_format="135",
    frame_width=36, frame_height=24, grain_size=6, halation_size=1.0, halation_green_factor=0.3,
    projector_kelvin=6500, inversion_gamma=4.0, idealized_curve=False, halation_intensity=1, shadow_comp=0,
    white_clip=False, white_balance=False, sat_adjust=1, grain_sigma=0.4, gamma_func="sRGB", push_pull=0.0,
    sharpening_strength=0.0, sharpening_sigma=1.0, color_masking=1.0)
GUI_IMAGE_DEFAULTS = dict(
    exp_comp=0, zoom=1, rotate_times=0, rotation=0, exp_kelvin=6000, profile="Default", canvas_mode="No",
    canvas_scale=1.0, canvas_ratio=0.8, highlight_burn=0, burn_scale=50, flip=False, tint=0, chroma_nr=0)


def merged(settings: dict) -> dict:
    """Defaults overlaid by the caller's flat dict; unknown keys are kept and ignored downstream,
    like the reference's `**_` (cpu_processor.py:321)."""
    s = dict(DEFAULTS)
    s.update(settings)
    return s


def pixels_per_mm(h: int, w: int, frame_width, frame_height) -> float:
    """cpu_processor.py:366."""
    return max(h, w) / max(frame_width, frame_height)


def stage_flags(s: dict, negative_film) -> int:
    """The `if` chain of cpu_processor.py:368-403 as R2F_* stage flags."""
    flags = 0
    if s["halation"]:
        flags |= F.HALATION
    if s["sharpness"] and negative_film.mtf is not None:
        flags |= F.MTF
    if s["grain"] and negative_film.rms_density is not None:
        flags |= F.GRAIN
        if s["grain"] == 1:
            flags |= F.GRAIN_BW
    if s["highlight_burn"] and (s["print_film"] is not None
                                or negative_film.density_measure in ["status_m", "bw"]):
        flags |= F.BURN
    return flags


def shard_frames(n_frames: int, world_size: int, rank: int) -> list[int]:
    """Whole-frame round-robin partition for batch export: frame i -> rank i mod world_size
    (generalises the reference's single consumer, gui_objects.py:65-115, to one consumer per GPU)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    return list(range(rank, n_frames, world_size))
