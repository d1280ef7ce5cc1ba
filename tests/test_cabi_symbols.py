"""The C-ABI library loads on a machine without a GPU, exports every symbol include/r2f_b200.h
declares, and reports errors through codes + r2f_last_error (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    text = open(os.path.join(ROOT, "include", "r2f_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(r2f_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from raw2film_b200 import _cabi

    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/r2f_b200.h but not exported"
    assert sorted(_cabi.EXPORTS) == names, "python binding list out of sync with the header"


def test_abi_version_and_constants_match_header():
    from raw2film_b200 import _cabi, flags

    text = open(os.path.join(ROOT, "include", "r2f_b200.h")).read()
    defs = {k: int(v.rstrip("u"), 0) for k, v in re.findall(r"#define\s+(R2F_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+u?|\d+)", text)}
    assert _cabi.lib.r2f_abi_version() == defs["R2F_ABI_VERSION"] == _cabi.ABI_VERSION
    assert (flags.HALATION, flags.MTF, flags.GRAIN, flags.GRAIN_BW, flags.BURN) == (
        defs["R2F_HALATION"], defs["R2F_MTF"], defs["R2F_GRAIN"], defs["R2F_GRAIN_BW"], defs["R2F_BURN"])
    for name, tap in flags.TAPS.items():
        assert defs["R2F_TAP_" + name.upper()] == tap
    assert defs["R2F_PROF_COUNT"] == len(_cabi.PROF_NAMES)


def test_errors_are_codes_with_messages():
    from raw2film_b200 import _cabi

    lib = _cabi.lib
    assert lib.r2f_create(0, None) != 0 and b"null" in lib.r2f_last_error()
    assert lib.r2f_set_lut2d(None, None, 0) != 0
    assert lib.r2f_workspace_bytes(0, 10, 0) == 0
    assert lib.r2f_workspace_bytes(4000, 6000, 0x7) >= 4000 * 6000 * 4 * 9
    assert lib.r2f_launch_count(None) == 0
    with pytest.raises(_cabi.R2FError):
        _cabi.check(lib.r2f_set_burn(None, 0.0, 0.0, 50.0))


def test_processor_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from raw2film_b200 import B200Processor

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200Processor()
