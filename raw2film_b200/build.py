"""Build libr2f_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libr2f_b200.so")
SOURCES = ["r2f_api.cu", "r2f_kernels.cu", "r2f_fft.cu", "r2f_conv_sym.cu", "r2f_grain_sym.cu", "r2f_resize.cu"]
HEADERS = ["device_math.cuh", "fast_chain.cuh", "conv_tile.cuh", "noise.cuh", "sym_conv.cuh", "r2f_kernels.h", "r2f_fft.h", "r2f_resize.h", os.path.join("..", "..", "include", "r2f_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",          # pointwise chain must round every float op separately (bit-exact vs oracle)
    "-Xcompiler", "-fPIC", "-shared", "--threads", "4",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources if needed; returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libr2f_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:/bin:" + env.get("PATH", "")     # host compiler: the system gcc
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True))
