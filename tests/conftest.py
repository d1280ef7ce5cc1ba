import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

os.environ.setdefault("NUMBA_THREADING_LAYER", "workqueue")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # Build the native pieces once if a toolchain is present (the GPU box uses the prebuilt .so files
    # that travelled with the snapshot; nvcc is there too, so a stale library is rebuilt as well).
    from oracle import film_oracle

    film_oracle.build_c_oracle()
    try:
        from raw2film_b200 import build as r2f_build

        r2f_build.build()
    except Exception as exc:  # noqa: BLE001
        if not os.path.exists(os.path.join(ROOT, "raw2film_b200", "libr2f_b200.so")):
            raise RuntimeError(f"libr2f_b200.so missing and could not be built: {exc}") from exc


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
