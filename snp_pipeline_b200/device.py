"""One lazily created libsnpgpu context per process (the reference runs each subcommand as its own process; a
device is chosen with --device / $SNPGPU_DEVICE / CUDA_VISIBLE_DEVICES).  There is no CPU fallback: if the
library or a GPU is missing this raises, and the subcommand's excepthook turns that into exit 100 / 98."""
from __future__ import annotations

import os

from . import _lib

_ctx = None


def context():
    global _ctx
    if _ctx is None:
        _ctx = _lib.Context(int(os.environ.get("SNPGPU_DEVICE", "0")))
    return _ctx


def release():
    global _ctx
    if _ctx is not None:
        _ctx.close()
        _ctx = None


def bind_to_gpu_cpus(device_index=0):
    """Pins this process to the CPU cores next to the GPU (NVML's CPU affinity for the device), so that pinned host
    buffers allocated afterwards land on the GPU's NUMA node and host<->device copies do not cross sockets.
    Returns the CPU set, or None when NVML (nvidia-ml-py) is unavailable or the call fails: a hint, never an error."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = device_index
            if visible:
                tok = visible.split(",")[device_index].strip()
                idx = int(tok) if tok.isdigit() else None
            h = pynvml.nvmlDeviceGetHandleByIndex(idx) if idx is not None else pynvml.nvmlDeviceGetHandleByUUID(tok)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                return cpus
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        pass
    return None
