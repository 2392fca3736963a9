"""cadr_b200 — B200-native (sm_100a) backend for CADR's per-frame drawable processing and data upload.

Layout (only what the hot path needs):
  csrc/     hand-written CUDA kernels + the C ABI (include/cadr_b200.h) -> lib/libcadr_b200.so
  host/     C++ facade with the reference's class names (CadR::Renderer, DataStorage, ...) over the C ABI
  _capi.py  ctypes binding of the C ABI (what a Python host binds)
  synth.py  synthetic scene generators for BASELINE.json's configs (inputs for tests and bench)
  frame.py  per-frame driver over the C ABI (drawable processing, culling, multi-GPU exchange)

Importing the package does not load the shared library; `cadr_b200.Context(...)` does and raises if it was
not built.  There is no CPU/PyTorch fallback anywhere in this package.
"""
from ._capi import (CadrError, Context, CopyRegion, CullParams, ExchangeSync, HandlePatch, LogicError, NoDevice,  # noqa: F401
                    OutOfResources, Timeout, LIB_PATH, SYMBOLS, lib)

__all__ = ["CadrError", "Context", "CopyRegion", "CullParams", "ExchangeSync", "HandlePatch", "LogicError", "NoDevice",
           "OutOfResources", "Timeout", "LIB_PATH", "SYMBOLS", "lib"]
