"""Shared helpers of the parity tests: run a Scene through the oracle (CPU) and through the C ABI (GPU)."""
from __future__ import annotations

import numpy as np

from cadr_b200 import synth
from cadr_b200.frame import DeviceScene, canon_equal, canonicalise
from oracle import binding as ob

FAKE_BASE = 0x7F1200000000
FAKE_LIST = 0x7F2000000000


def oracle_tier_r(scene: synth.Scene, arena_base: int = FAKE_BASE, list_base: int = FAKE_LIST, img=None, threads=1):
    img = scene.image(arena_base) if img is None else img
    dl = np.ascontiguousarray(scene.drawables)
    mem = ob.Memory([(arena_base, img), (list_base, dl)])
    ind, ptr = ob.process_drawables(mem, arena_base + scene.root_off, scene.handle_level, list_base, scene.n, threads)
    return mem, ind, ptr


def oracle_tier_x(scene: synth.Scene, planes, eye, arena_base: int = FAKE_BASE, list_base: int = FAKE_LIST, img=None):
    mem, ind, ptr = oracle_tier_r(scene, arena_base, list_base, img)
    res = ob.cull_compact(mem, arena_base + scene.root_off, scene.handle_level, list_base, scene.n, ind, ptr,
                          scene.cull, planes, eye, scene.regions)
    return ind, ptr, res


def gpu_frame(ctx, scene: synth.Scene, planes=None, eye=None):
    """-> (DeviceScene, indirect, pointers, tier-x result or None); caller closes the DeviceScene."""
    ds = DeviceScene(ctx, scene)
    ds.record_drawable_processing()
    res = None
    if planes is not None:
        ds.cull(planes, eye)
    ctx.sync(ds.stream)
    ind, ptr = ds.read_tier_r()
    if planes is not None:
        res = ds.read_tier_x()
    return ds, ind, ptr, res


def assert_tier_x_equal(gpu: dict, ref: dict):
    assert gpu["status"] == 0, f"GPU reported overflow status {gpu['status']}"
    assert ref["status"] == 0
    assert np.array_equal(gpu["inst_count"], ref["inst_count"]), "per-StateSet survivor counts differ"
    assert gpu["near_band"] == ref["near_band"], "near-band counts differ"
    ok, why = canon_equal(canonicalise(gpu), canonicalise(ref))
    assert ok, why
