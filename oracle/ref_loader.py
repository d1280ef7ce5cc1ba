"""Container-only loader for the UNMODIFIED reference modules (test infrastructure).

TEST INFRASTRUCTURE -- never imported by the product path (raw2film_b200/).

`/root/reference` exists only in the build container; nothing that runs on the
GPU box may call this.  It is used by `tests/golden/make_golden.py` (to mint the
committed golden vectors from the reference's own functions) and by the
container-only cross-checks in `tests/test_oracle_vs_reference.py`.

The reference imports several packages that are not installable here
(spectral_film_lut, lensfunpy, rawpy, exiftool, colour, wgpu; see SURVEY.md
appendix A).  None of them is executed by the functions we borrow
(`raw2film.effects`, `raw2film.utils`), so each is replaced by an empty stub
module carrying only the names the reference touches at import time.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "raw2film"))


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load_reference():
    """Return (effects, utils) modules of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present (GPU box?) - golden fixtures only")
    os.environ.setdefault("NUMBA_THREADING_LAYER", "workqueue")  # __main__.py:12
    if "raw2film.effects" in sys.modules:
        return sys.modules["raw2film.effects"], sys.modules["raw2film.utils"]

    def _missing(*_a, **_k):
        raise NotImplementedError("third-party spectral_film_lut is not available")

    _stub("spectral_film_lut", BASE_DIR="")
    _stub("spectral_film_lut.config", DEFAULT_DTYPE=np.float32)
    _stub("spectral_film_lut.film_spectral", FilmSpectral=object)
    _stub("spectral_film_lut.grain_generation", generate_grain=_missing, grain_kernel=_missing)
    _stub("spectral_film_lut.utils", create_lut=_missing, log_clip=_missing,
          multi_channel_interp=_missing)
    _stub("spectral_film_lut.xy_lut", apply_2d_lut=_missing)
    _stub("spectral_film_lut.color_space", GAMMA_KEYS=str)
    lf = _stub("lensfunpy")
    lf.util = _stub("lensfunpy.util")
    _stub("rawpy")
    _stub("exiftool")
    _stub("colour", convert=lambda v, a, b: np.array([0.5, 0.5, 0.5]))
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    from raw2film import effects, utils  # noqa: E402

    return effects, utils
