"""GPU: the optional per-drawable bounds pre-test (cadr_b200_compute_drawable_bounds + cadr_cull_params.drawableBounds).
The contract is that the table never changes a frame's result — the oracle knows nothing about it — while long lists
outside the frustum are dropped before their matrices are read."""
import numpy as np
import pytest
import torch

from cadr_b200 import synth
from cadr_b200.frame import DeviceScene
from helpers import assert_tier_x_equal, oracle_tier_x
from test_fullsize_gpu import build, gpu_summary

pytestmark = pytest.mark.gpu


def frame(ctx, ds, planes, eye, fused=True):
    if fused:
        ds.upload_drawable_list()
        ds.process_and_cull(planes, eye)
    else:
        ds.record_drawable_processing()
        ds.cull(planes, eye)
    ctx.sync(ds.stream)
    return ds.read_tier_x()


def clusterise(sc, seed, cube=400.0, sigma=6.0):
    """random_scene spreads every list over the whole cube; give each list a centre so that bounds mean something"""
    rng = np.random.default_rng(seed)
    start = 0
    for cnt in sc.ml_count:
        cnt = int(cnt)
        centre = (rng.random(3) - 0.5) * cube
        sc.matrices[start:start + cnt, 12:15] = (centre + rng.normal(0, sigma, (cnt, 3))).astype(np.float32)
        start += cnt


@pytest.mark.parametrize("kw", [dict(seed=71, n=500, num_lists=60, max_count=300, state_sets=4, big_lists=4),
                                dict(seed=72, n=300, list_counts=[33, 64, 100, 700, 1500, 2500, 40, 5, 0], state_sets=3, first_handle=2040),
                                dict(seed=76, n=900, list_counts=[4, 8, 16, 32, 5, 31, 3, 1, 2, 0], state_sets=3)],
                         ids=["ragged", "boundaries", "short-lists-only"])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "two-calls"])
def test_result_is_identical_with_bounds_and_work_items_are_dropped(ctx, kw, fused):
    sc = synth.random_scene(**kw)
    clusterise(sc, kw["seed"])
    ds = DeviceScene(ctx, sc)
    try:
        ds.record_drawable_processing()
        ctx.sync(ds.stream)
        queued_without = {}
        for f in (0, 90, 200):
            planes, eye = synth.orbit_camera(f, 250.0, far=500.0)
            queued_without[f] = frame(ctx, ds, planes, eye, fused)["chunk_count"]
        ds.compute_bounds()
        dropped = 0
        for f in (0, 90, 200):
            planes, eye = synth.orbit_camera(f, 250.0, far=500.0)
            got = frame(ctx, ds, planes, eye, fused)
            _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
            assert_tier_x_equal(got, ref)
            assert got["chunk_count"] <= queued_without[f]
            dropped += queued_without[f] - got["chunk_count"]
        assert dropped > 0 or not (sc.ml_count > 32).any()      # short lists are dropped by their own thread, not from the queue
    finally:
        ds.close()


def test_lists_hugging_a_plane_are_never_dropped_wrongly(ctx):
    """Lists of 40 unit spheres whose instances sit at x = -1 - delta for deltas around the 1e-5 near band, against the
    plane x >= 0: whatever the pre-test decides, visible set and near-band count must equal the oracle's."""
    deltas = [0.0, 1e-7, 1e-6, 5e-6, 9e-6, 1e-5, 1.1e-5, 2e-5, 5e-5, 1e-4, 1e-3, 1e-2, -1e-6, -1e-5, -1e-3, 0.5, 30.0]
    sc = synth.random_scene(73, n=len(deltas), list_counts=[40] * len(deltas), state_sets=1, with_drawable_data=False)
    m = np.zeros((40 * len(deltas), 16), np.float32)
    m[:, 0] = m[:, 5] = m[:, 10] = m[:, 15] = 1.0
    rng = np.random.default_rng(3)
    for k, dlt in enumerate(deltas):
        rows = slice(40 * k, 40 * k + 40)
        m[rows, 12] = np.float32(-1.0) - np.float32(dlt) - (rng.random(40, dtype=np.float32) * np.float32(abs(dlt) * 0.5))
        m[rows, 13] = (rng.random(40, dtype=np.float32) - 0.5) * 100
        m[rows, 14] = (rng.random(40, dtype=np.float32) - 0.5) * 100
    sc.matrices[:] = m
    sc.cull[:, 0:3] = 0
    sc.cull[:, 3] = np.float32(1.0).view(np.uint32)
    sc.cull[:, 4] = 1                                        # one LOD
    big = 1e9
    planes = np.array([[1, 0, 0, 0], [-1, 0, 0, big], [0, 1, 0, big], [0, -1, 0, big], [0, 0, 1, big], [0, 0, -1, big]], np.float32)
    eye = np.zeros(3, np.float32)
    ds = DeviceScene(ctx, sc)
    try:
        ds.record_drawable_processing(); ctx.sync(ds.stream)
        ds.compute_bounds()
        got = frame(ctx, ds, planes, eye)
        _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        assert ref["near_band"] > 0 and 0 < ref["inst_count"].sum() < sc.total_instances
        assert_tier_x_equal(got, ref)
        # margin here = 2e-5 + 2^-18 * (|x| + reach) ~ 2.8e-5: every list further out than that is dropped, every list
        # closer (the near-band ones among them) is kept and evaluated
        assert got["chunk_count"] == sum(1 for d in deltas if d < 2.9e-5)
    finally:
        ds.close()


def test_bounds_follow_rewritten_matrices(ctx):
    """Bounds are stale after a MatrixList was rewritten; recomputing them for the affected drawables (index list)
    restores the contract."""
    sc = synth.random_scene(74, n=60, list_counts=[200] * 60, state_sets=2, with_drawable_data=False)
    clusterise(sc, 74)
    ds = DeviceScene(ctx, sc)
    try:
        ds.record_drawable_processing(); ctx.sync(ds.stream)
        ds.compute_bounds()
        planes, eye = synth.orbit_camera(10, 250.0, far=500.0)
        # move three lists from wherever they are to right in front of the camera
        moved = np.array([3, 17, 44])
        target = eye + (np.zeros(3) - eye) / np.linalg.norm(eye) * 100.0
        for l in moved:
            rows = slice(int(sc.ml_count[:l].sum()), int(sc.ml_count[:l + 1].sum()))
            sc.matrices[rows, 12:15] = target.astype(np.float32) + np.random.default_rng(int(l)).normal(0, 3, (200, 3)).astype(np.float32)
            blk = np.ascontiguousarray(sc.matrices[rows]).view(np.uint8).reshape(-1)
            ctx.memcpy_h2d(ds.arena + int(sc.ml_off[l]) + 64, blk, stream=ds.stream)
        drawables = np.flatnonzero(np.isin(sc.drawable_ml, moved)).astype(np.uint32)
        idx = ctx.arena_alloc(max(drawables.nbytes, 16))
        ctx.memcpy_h2d(idx, drawables, stream=ds.stream)
        ds.compute_bounds(indices=idx, count=len(drawables))
        got = frame(ctx, ds, planes, eye)
        _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        assert_tier_x_equal(got, ref)
        vis = {int(d) for s in range(sc.num_state_sets) for d in got["tag"][int(sc.regions[s, 0]):int(sc.regions[s, 0]) + int(got["cmd_count"][s]), 0]}
        assert set(drawables.tolist()) <= vis                   # the moved lists are now seen
        ctx.arena_free(idx)
    finally:
        ds.close()


def test_c3_full_size_with_bounds_equals_without(ctx):
    """100 M instances: the frame with the pre-test has the same per-StateSet counts and index checksums as without,
    and queues a fraction of the work items."""
    scene = synth.config3(100_000, 1000, state_sets=64, host_matrices=False)
    ds, arena, stream = build(ctx, scene)
    try:
        with torch.cuda.stream(stream):
            ds.record_drawable_processing()
            planes, eye = synth.orbit_camera(40, 1500.0, far=3000.0)
            ds.process_and_cull(planes, eye); stream.synchronize()
            a = gpu_summary(ds, arena, scene); qa = ds.read_counters()["chunk_count"]
            ds.compute_bounds(); stream.synchronize()
            ds.process_and_cull(planes, eye); stream.synchronize()
            b = gpu_summary(ds, arena, scene); qb = ds.read_counters()["chunk_count"]
        assert a["status"] == 0 and b["status"] == 0
        assert np.array_equal(a["inst_count"], b["inst_count"]) and np.array_equal(a["cmd_count"], b["cmd_count"])
        assert (a["sum_k"], a["key_digest"], a["idx_sum"], a["idx_sq"]) == (b["sum_k"], b["key_digest"], b["idx_sum"], b["idx_sq"])
        assert qa == 100_000 and qb < 0.6 * qa
    finally:
        ds.close()


def test_bounds_api_misuse_is_a_logic_error(ctx):
    import cadr_b200
    sc = synth.random_scene(75, n=50, list_counts=[100] * 5, state_sets=1)
    ds = DeviceScene(ctx, sc)
    try:
        ds.record_drawable_processing(); ctx.sync(ds.stream)
        p = ds.cull_params(np.zeros((6, 4), np.float32), np.zeros(3, np.float32))
        buf = ctx.arena_alloc(50 * 32 + 64)
        with pytest.raises(cadr_b200.LogicError):
            ctx.compute_drawable_bounds(p, buf + 16, 50)            # 32-byte alignment
        with pytest.raises(cadr_b200.LogicError):
            ctx.compute_drawable_bounds(p, buf, 51)                 # more than numDrawables without an index list
        with pytest.raises(cadr_b200.LogicError):
            ctx.compute_drawable_bounds(p, 0, 50)
        p.drawableBounds = buf + 16
        with pytest.raises(cadr_b200.LogicError):
            ctx.cull_compact(p)
        ctx.compute_drawable_bounds(p, buf, 0)                      # nothing to do is fine
        ctx.arena_free(buf)
    finally:
        ds.close()
