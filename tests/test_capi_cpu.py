"""CPU suite: the C-ABI library loads and exports every symbol include/cadr_b200.h declares, the error
taxonomy works, and nothing computes without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import cadr_b200
from cadr_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cadr_b200.h")).read()
    return sorted(set(re.findall(r"CADR_API\s+[\w\s\*]+?\b(cadr_b200_\w+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    assert declared_symbols() == sorted(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    l = ctypes.CDLL(cadr_b200.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(l, name), name
    assert cadr_b200.lib().cadr_b200_abi_version() == 6


def test_struct_sizes_match_header():
    assert ctypes.sizeof(_capi.CopyRegion) == 24
    assert ctypes.sizeof(_capi.HandlePatch) == 16
    # cadr_cull_params: 6 u64-ish head + planes/eye + tail (kept in sync by hand; checked against offsets)
    assert _capi.CullParams.planes.offset == 48
    assert _capi.CullParams.eye.offset == 48 + 96
    assert _capi.CullParams.numStateSets.offset == 160
    assert _capi.CullParams.stateSetRegions.offset == 168
    assert _capi.CullParams.exchangeWorld.offset == 232
    assert _capi.CullParams.exchangeTag.offset == 232 + 16 + 2 * 64
    assert _capi.CullParams.drawableBounds.offset == 232 + 16 + 3 * 64
    assert ctypes.sizeof(_capi.CullParams) == 232 + 16 + 3 * 64 + 16
    assert ctypes.sizeof(_capi.ExchangeSync) == 32 + 2 * 64
    assert _capi.CullParams.addressDelta.offset == 232 + 16 + 3 * 64 + 8          # the consumer-side pointer translation (ABI 6)
    assert ctypes.sizeof(_capi.ExchangePull) == 16 + 3 * 8 + 8 + 2 * 64 and _capi.ExchangePull.regions.offset == 48
    assert cadr_b200.lib().cadr_b200_cull_counters_bytes(64) == 64 + 64 * 8


def test_address_space_only_context_refuses_compute(address_ctx):
    c = address_ctx
    assert c.device == -1
    a = c.arena_alloc(1000)
    b = c.arena_alloc(16)
    assert a % 256 == 0 and b % 256 == 0 and b >= a + 1000
    for call in (lambda: c.process_drawables(a, 1, a, a, a, 4),
                 lambda: c.cull_compact(_capi.CullParams()),
                 lambda: c.upload([(a, 0, 16)], np.zeros(16, np.uint8)),
                 lambda: c.scatter_copy([(a, 0, 16)], b),
                 lambda: c.patch_handles(a, 1, [(1, 2)]),
                 lambda: c.upload_stage([(a, 0, 16)], np.zeros(16, np.uint8)),
                 lambda: c.upload_commit(0x181),
                 lambda: c.exchange_pull_instances(_capi.ExchangePull()),
                 lambda: c.ipc_export_range(a),
                 lambda: c.memcpy_h2d(a, np.zeros(16, np.uint8)),
                 lambda: c.sync()):
        with pytest.raises(cadr_b200.NoDevice):
            call()
    c.arena_free(a)
    with pytest.raises(cadr_b200.LogicError):
        c.arena_free(a)  # double free is API misuse
    c.arena_free(b)


def test_no_gpu_means_no_context_not_a_fallback():
    """On a box without a GPU, creating a real context fails loudly; on a GPU box it succeeds."""
    try:
        c = cadr_b200.Context(0)
    except cadr_b200.NoDevice as e:
        assert "no CPU fallback" in str(e)
    else:
        assert c.sm_count > 0
        c.close()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under cadr_b200/ or include/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cadr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                t = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\boracle\b", t) and not f == "build.py":
                    for line in t.splitlines():
                        if re.search(r"\boracle\b", line) and re.search(r"import|include|dlopen|CDLL|-l", line):
                            bad.append((f, line.strip()))
    assert not bad, bad
