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
