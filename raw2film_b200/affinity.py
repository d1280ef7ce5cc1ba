"""Bind a rank to the CPU cores / NUMA node closest to its GPU (NVML), so that pinned staging
buffers are first-touched next to the PCIe root the GPU hangs off.  Host-side plumbing for the
multi-GPU batch export ("near-linear scaling to 8 GPUs additionally depends on host
memory bandwidth feeding 8 PCIe links")."""
from __future__ import annotations

import os


def bind_to_gpu(device_index: int) -> dict:
    """Restrict this process to the CPUs NVML reports as local to `device_index`.
    Returns a small report; never raises (affinity is an optimisation)."""
    info = {"bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        # honour CUDA_VISIBLE_DEVICES remapping: NVML enumerates physical devices
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1 and w * 64 + b < ncpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed), first=allowed[0], last=allowed[-1])
        try:
            info["numa"] = pynvml.nvmlDeviceGetNumaNodeId(handle)
        except Exception:  # noqa: BLE001
            pass
    except Exception as exc:  # noqa: BLE001
        info["error"] = str(exc)[:120]
    return info
