"""Swap point for the PARITY-UNPINNED stages (TEST INFRASTRUCTURE ONLY).

`apply_2d_lut`, `log_clip`, `multi_channel_interp`, `generate_grain` and `grain_kernel` belong to
spectral-film-lut >= 0.8.0 (reference pyproject.toml:28), which is neither in the reference tree nor installable
offline; oracle/film_oracle.py restates them from the WGSL shaders the reference ships.

The day the package is importable, `bind(film_oracle)` replaces those restatements with the package's own
functions, and `python tests/golden/make_golden.py third_party` mints `tests/golden/third_party.npz` from them;
tests/test_oracle_golden.py::test_third_party_goldens then pins the restatements (it is skipped while the file
does not exist).  Nothing else has to change: the CUDA path is compared with whatever film_oracle exposes.

Call sites in the reference: cpu_processor.py:364 (apply_2d_lut), :378 (log_clip), :380 (multi_channel_interp),
effects.py:230-233 (generate_grain, FilmSpectral.grain_transform), gpu_processor.py:927-929 (grain_kernel).
"""
from __future__ import annotations

NAMES = ("apply_2d_lut", "log_clip", "multi_channel_interp", "generate_grain", "grain_kernel")


def load():
    """{name: function} of the real package's functions, or {} when it is not installed."""
    found = {}
    try:
        import spectral_film_lut.utils as sfl_utils  # type: ignore
    except Exception:  # noqa: BLE001 - absent offline
        return found
    for name in ("apply_2d_lut", "log_clip", "multi_channel_interp"):
        fn = getattr(sfl_utils, name, None)
        if callable(fn):
            found[name] = fn
    try:
        import spectral_film_lut.grain_generation as sfl_grain  # type: ignore

        for name in ("generate_grain", "grain_kernel"):
            fn = getattr(sfl_grain, name, None)
            if callable(fn):
                found[name] = fn
    except Exception:  # noqa: BLE001
        pass
    return found


def available() -> bool:
    return bool(load())


def bind(film_oracle_module) -> list[str]:
    """Rebind the restated stages of `film_oracle_module` to the package's functions; returns what was bound."""
    bound = []
    for name, fn in load().items():
        if name in ("generate_grain",):     # different signature (shape, scale, grain_size_mm, bw, cached, sigma)
            continue
        setattr(film_oracle_module, name, fn)
        bound.append(name)
    return bound
