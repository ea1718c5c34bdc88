"""pslite_b200 — a Blackwell-native parameter-server framework.

Capabilities follow bytedance/ps-lite (scheduler / server / worker roles, DMLC_* env
discovery, KVWorker push/pull, KVServer request handlers, barriers) re-designed for an
8xB200 NVSwitch box: values travel as one-sided writes into peer HBM issued by sm_100a
kernels (the "nvl" van), the server-side handler is a fused dequant+sum+AdamW kernel, and
a plain TCP van carries the control plane and CPU-only jobs.

Layout
    pslite_b200._C        native runtime + torch binding (built in-tree by pslite_b200.build)
    pslite_b200.ops       thin wrappers over the raw kernels (+ PyTorch references for tests)
    pslite_b200.parallel  launcher glue (torchrun -> DMLC roles) and the PS trainer
    pslite_b200.models    Llama-3 and ResNet-50 definitions used by the benchmarks
    pslite_b200.utils     env, timing, clocks sampling
"""
from __future__ import annotations

import importlib
import os

__version__ = "0.1.0"

_C = None


def _load_native():
    """Import the native extension; build it in-tree on first use if it is missing."""
    global _C
    if _C is not None:
        return _C
    # the extension hands torch tensors to Python (alloc_exportable, ...): the torch Python
    # package must be initialised first, or wrapping the first tensor dereferences a null type
    import torch  # noqa: F401
    try:
        _C = importlib.import_module("pslite_b200._C")
    except ImportError:
        if os.environ.get("PSLITE_NO_AUTOBUILD"):
            raise
        from . import build as _build

        _build.build()
        _C = importlib.import_module("pslite_b200._C")
    return _C


def native():
    """The native module (`pslite_b200._C`)."""
    return _load_native()
