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


def fold_by_drawable_lod(result: dict, n: int):
    """The per-(drawable, lod) summary of an emitted Tier X result (numpy): survivors, index sum, index sum of squares
    - what oracle/binding.cull_summary computes straight from the matrices.  Commands of one (drawable, lod) that were
    emitted per work item are folded together."""
    k = np.zeros((n, 3), np.uint64); sm = np.zeros((n, 3), np.uint64); sq = np.zeros((n, 3), np.uint64)
    regions = result["regions"]
    for s in range(regions.shape[0]):
        base = int(regions[s, 0])
        for ci in range(base, base + int(result["cmd_count"][s])):
            d, lod = int(result["tag"][ci, 0]), int(result["tag"][ci, 1])
            cnt, first = int(result["cmd"][ci, 1]), int(result["cmd"][ci, 4])
            run = result["inst"][first:first + cnt].astype(np.uint64)
            k[d, lod] += np.uint64(cnt); sm[d, lod] += run.sum(dtype=np.uint64); sq[d, lod] += (run * run).sum(dtype=np.uint64)
    return k, sm, sq
