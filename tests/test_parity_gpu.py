"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle on identical seeded inputs.
Bit-exact for everything: Tier R is integer; Tier X is compared after canonical sorting within each StateSet
(emission order depends on atomic arrival), including the near-band instances, because oracle and kernel
perform the same fp32 operations in the same order without FMA contraction."""
import numpy as np
import pytest

import cadr_b200
from cadr_b200 import synth
from cadr_b200.frame import DeviceScene, canonicalise
from oracle import binding as ob
from helpers import assert_tier_x_equal, gpu_frame, oracle_tier_r, oracle_tier_x

pytestmark = pytest.mark.gpu

R_CASES = [
    dict(seed=1), dict(seed=2, first_handle=1990), dict(seed=3, first_handle=3000, force_level=3),
    dict(seed=4, first_handle=4_194_250, big_lists=2), dict(seed=5, n=1, num_lists=3),
    dict(seed=6, n=257, with_drawable_data=False), dict(seed=7, n=4099, num_lists=900, num_geometries=40),
]


def _oracle_r(ds: DeviceScene):
    sc = ds.scene
    return oracle_tier_r(sc, arena_base=ds.arena, list_base=ds.drawable_list)


@pytest.mark.parametrize("kw", R_CASES, ids=lambda k: f"seed{k['seed']}")
def test_tier_r_random_scenes(ctx, kw):
    sc = synth.random_scene(**kw)
    ds, ind, ptr, _ = gpu_frame(ctx, sc)
    try:
        _, e_ind, e_ptr = _oracle_r(ds)
        assert np.array_equal(ind, e_ind)
        assert np.array_equal(ptr, e_ptr)
    finally:
        ds.close()


@pytest.mark.parametrize("builder", [lambda: synth.config1(12), lambda: synth.config2(70_001),
                                     lambda: synth.config3(300, 100, state_sets=7)], ids=["C1", "C2", "C3"])
def test_tier_r_baseline_shapes(ctx, builder):
    sc = builder()
    ds, ind, ptr, _ = gpu_frame(ctx, sc)
    try:
        _, e_ind, e_ptr = _oracle_r(ds)
        assert np.array_equal(ind, e_ind) and np.array_equal(ptr, e_ptr)
    finally:
        ds.close()


def test_zero_drawables_is_a_no_op(ctx):
    """Renderer.cpp:600-620: nothing is dispatched for an empty scene."""
    before = ctx.launch_count
    ctx.process_drawables(0, 1, 0, 0, 0, 0)
    assert ctx.launch_count == before


def test_api_misuse_is_a_logic_error(ctx):
    a = ctx.arena_alloc(4096)
    try:
        with pytest.raises(cadr_b200.LogicError):
            ctx.process_drawables(a, 4, a, a, a, 10)           # no such handle level
        with pytest.raises(cadr_b200.LogicError):
            ctx.process_drawables(a, 1, a, a, a, 1 << 30)      # Renderer.cpp:687 limit
        with pytest.raises(cadr_b200.LogicError):
            ctx.process_drawables(a, 1, a + 8, a, a, 10)       # misaligned list
        with pytest.raises(cadr_b200.LogicError):
            ctx.patch_handles(a, 1, [(5000, 1)])               # handle outside a level-1 table
    finally:
        ctx.arena_free(a)


X_CASES = [
    (dict(seed=11), 0), (dict(seed=12, big_lists=3), 40), (dict(seed=13, first_handle=4_194_250, big_lists=1), 200),
    (dict(seed=14, n=2000, num_lists=400, max_count=40, state_sets=17), 120),
    (dict(seed=15, n=700, num_lists=60, max_count=300, state_sets=3, big_lists=5), 300),
    (dict(seed=16, n=1, num_lists=2), 0),
]


@pytest.mark.parametrize("kw,frame", X_CASES, ids=lambda v: f"seed{v['seed']}" if isinstance(v, dict) else f"f{v}")
def test_tier_x_random_scenes(ctx, kw, frame):
    sc = synth.random_scene(**kw)
    planes, eye = synth.orbit_camera(frame, 250.0, far=500.0)
    ds, ind, ptr, got = gpu_frame(ctx, sc, planes, eye)
    try:
        _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        assert_tier_x_equal(got, ref)
    finally:
        ds.close()


@pytest.mark.parametrize("builder,radius,far", [(lambda: synth.config2(200_003), 1500.0, 1500.0),
                                                (lambda: synth.config3(600, 1000, state_sets=64), 1500.0, 3000.0),
                                                (lambda: synth.config3(64, 2500, state_sets=5), 1500.0, 3000.0)],
                         ids=["C2", "C3", "C3-multi-item"])
def test_tier_x_baseline_shapes(ctx, builder, radius, far):
    sc = builder()
    ds = DeviceScene(ctx, sc)
    try:
        for frame in (0, 77):
            planes, eye = synth.orbit_camera(frame, radius, far=far)
            ds.record_drawable_processing()
            ds.cull(planes, eye)
            ctx.sync(ds.stream)
            got = ds.read_tier_x()
            _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
            assert_tier_x_equal(got, ref)
            p = got["inst_count"].sum() / sc.total_instances
            assert 0.02 < p < 0.9, f"survivor fraction {p}"
    finally:
        ds.close()


BOUNDARY_COUNTS = [0, 1, 2, 31, 32, 33, 34, 63, 64, 65, 95, 96, 97, 127, 128, 129, 255, 256, 257, 511, 512, 513, 514,
                   1023, 1024, 1025, 2047, 2048, 2049, 3072]


@pytest.mark.parametrize("fused", [False, True], ids=["two-calls", "fused"])
def test_tier_x_list_length_boundaries(ctx, fused):
    """Every hand-over point between the kernels: thread-per-list (<= 32 matrices), flat batches of medium lists
    (33..64), warp-per-item (> 64), work items of 1024, full and ragged last steps of 32, full and ragged last work
    items.  (The A/B kernels of csrc/experiments are not part of the product library; scripts/fuzz_parity.py runs
    them from libcadr_b200_exp.so.)"""
    sc = synth.random_scene(51, n=4 * len(BOUNDARY_COUNTS) + 3, list_counts=BOUNDARY_COUNTS, state_sets=4, first_handle=2030)
    ds = DeviceScene(ctx, sc)
    try:
        for frame in (10, 250):
            planes, eye = synth.orbit_camera(frame, 250.0, far=500.0)
            if fused:
                ds.upload_drawable_list()
                ds.process_and_cull(planes, eye)
            else:
                ds.record_drawable_processing()
                ds.cull(planes, eye)
            ctx.sync(ds.stream)
            got = ds.read_tier_x()
            _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
            assert_tier_x_equal(got, ref)
            assert 0 < got["inst_count"].sum() < sc.total_instances
    finally:
        ds.close()


@pytest.mark.parametrize("fused", [False, True], ids=["two-calls", "fused"])
@pytest.mark.parametrize("shape", ["all-medium", "mixed", "one-partial-batch", "exactly-32"])
def test_tier_x_medium_lists_in_flat_batches(ctx, shape, fused):
    """Lists of 33..64 matrices are consumed 32 items per warp as one flat run of instances (cullMediumBatches): several
    batches with a ragged last one, items of every length in the range, batches that mix StateSets, medium items queued
    next to long items (the two ends of one workspace) and short lists, under a culling camera, with everything
    visible (all 64 bits of the masks set) and with nothing visible."""
    rng = np.random.default_rng(97)
    if shape == "all-medium":
        counts, n = rng.integers(33, 65, 333).tolist(), 333
    elif shape == "mixed":
        counts = rng.integers(33, 65, 150).tolist() + [1, 0, 5, 32, 65, 100, 1024, 1500, 2100] * 6
        rng.shuffle(counts)
        n = len(counts) + 40                     # some lists are shared by two drawables
    elif shape == "one-partial-batch":
        counts, n = [64, 33, 50, 34, 63], 5
    else:
        counts, n = [33 + (k % 32) for k in range(32)], 32
    sc = synth.random_scene(52, n=n, list_counts=counts, state_sets=min(6, n), first_handle=2000)
    big = 1e9
    all_in = np.array([[1, 0, 0, big], [-1, 0, 0, big], [0, 1, 0, big], [0, -1, 0, big], [0, 0, 1, big], [0, 0, -1, big]], np.float32)
    none_in = all_in.copy(); none_in[0, 3] = -big
    cams = [synth.orbit_camera(10, 250.0, far=500.0), synth.orbit_camera(200, 250.0, far=500.0),
            (all_in, np.zeros(3, np.float32)), (none_in, np.zeros(3, np.float32))]
    ds = DeviceScene(ctx, sc)
    try:
        for planes, eye in cams:
            if fused:
                ds.upload_drawable_list()
                ds.process_and_cull(planes, eye)
            else:
                ds.record_drawable_processing()
                ds.cull(planes, eye)
            ctx.sync(ds.stream)
            got = ds.read_tier_x()
            _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
            assert_tier_x_equal(got, ref)
            cnt = sc.ml_count[sc.drawable_ml]
            assert got["medium_count"] == int(((cnt > 32) & (cnt <= 64)).sum())
            assert got["status"] == 0
    finally:
        ds.close()


@pytest.mark.parametrize("side", [36, 100], ids=["46k", "1M-as-in-the-reference"])
def test_reference_instanced_boxes_scene_and_camera(ctx, side):
    """RenderingPerformance InstancedBoxesScene (one drawable, one list of side^3 matrices; 100^3 in the reference)
    under the example's own orthographic camera: Tier R record + culled result against the oracle; the list spans
    46 / 977 work items."""
    sc = synth.config1_instanced(side)
    ds = DeviceScene(ctx, sc)
    try:
        for frame in (0, 1):
            planes, eye = synth.reference_camera(frame)
            ds.record_drawable_processing()
            ds.cull(planes, eye)
            ctx.sync(ds.stream)
            ind, ptr = ds.read_tier_r()
            got = ds.read_tier_x()
            e_ind, e_ptr, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
            assert np.array_equal(ind, e_ind) and np.array_equal(ptr, e_ptr)
            assert ind[0, 1] == side ** 3 and ind[0, 0] == 36                # instanceCount, vertexCount of the box
            assert_tier_x_equal(got, ref)
            # the ortho volume is 100 deep and the grid 340 wide: a slab of the boxes survives
            assert 0 < got["inst_count"].sum() < sc.total_instances
    finally:
        ds.close()


def test_tier_x_all_visible_none_visible_and_idempotent(ctx):
    sc = synth.random_scene(21, big_lists=2, n=900, num_lists=80)
    big = 1e9
    all_in = np.array([[1, 0, 0, big], [-1, 0, 0, big], [0, 1, 0, big], [0, -1, 0, big], [0, 0, 1, big], [0, 0, -1, big]], np.float32)
    none = all_in.copy(); none[0, 3] = -big
    eye = np.zeros(3, np.float32)
    ds = DeviceScene(ctx, sc)
    try:
        ds.record_drawable_processing()
        ds.cull(all_in, eye); ctx.sync(ds.stream)
        a1 = ds.read_tier_x()
        ds.cull(all_in, eye); ctx.sync(ds.stream)
        a2 = ds.read_tier_x()
        cnt = sc.ml_count[sc.drawable_ml].astype(np.int64)
        nonempty = sc.cull[:, 3].view(np.float32) >= 0
        assert a1["inst_count"].sum() == int(cnt[nonempty].sum())
        assert_tier_x_equal(a1, a2)                        # same frame twice: same canonical result
        ds.cull(none, eye); ctx.sync(ds.stream)
        z = ds.read_tier_x()
        assert z["inst_count"].sum() == 0 and z["cmd_count"].sum() == 0 and z["status"] == 0
    finally:
        ds.close()


def test_tier_x_region_overflow_is_reported(ctx):
    sc = synth.random_scene(22, n=500, num_lists=40)
    sc.regions = sc.regions.copy()
    sc.regions[:, 3] = np.minimum(sc.regions[:, 3], 3)     # instance regions far too small
    big = 1e9
    all_in = np.array([[1, 0, 0, big], [-1, 0, 0, big], [0, 1, 0, big], [0, -1, 0, big], [0, 0, 1, big], [0, 0, -1, big]], np.float32)
    ds = DeviceScene(ctx, sc)
    try:
        ds.record_drawable_processing()
        ds.cull(all_in, np.zeros(3, np.float32)); ctx.sync(ds.stream)
        assert ds.read_counters()["status"] & 1
    finally:
        ds.close()


def test_tier_x_work_item_queue_overflow_is_reported(ctx):
    """A work-item workspace smaller than the scene needs: status bit 1, the items that did not fit are dropped, the
    ones that did are still complete and correct (never a write past the workspace)."""
    sc = synth.random_scene(24, n=300, list_counts=[40, 700, 1500, 2500, 90, 33], state_sets=3)
    big = 1e9
    all_in = np.array([[1, 0, 0, big], [-1, 0, 0, big], [0, 1, 0, big], [0, -1, 0, big], [0, 0, 1, big], [0, 0, -1, big]], np.float32)
    ds = DeviceScene(ctx, sc)
    try:
        need = ds.chunk_cap
        assert need > 40
        guard = ctx.arena_alloc(128)                       # poison right behind a deliberately small workspace
        ds.chunk_cap = need // 3
        small_ws = ctx.arena_alloc(ds.chunk_cap * 128 + 256)
        ctx.memset(small_ws, 0xAB, ds.chunk_cap * 128 + 256)
        ds.chunk_ws = small_ws
        ds.record_drawable_processing()
        ds.cull(all_in, np.zeros(3, np.float32)); ctx.sync(ds.stream)
        c = ds.read_counters()
        assert c["status"] == 2
        tail = np.empty(256, np.uint8); ctx.memcpy_d2h(tail, small_ws + ds.chunk_cap * 128); ctx.sync()
        assert (tail == 0xAB).all()
        cnt = sc.ml_count[sc.drawable_ml].astype(np.int64)
        nonempty = sc.cull[:, 3].view(np.float32) >= 0
        assert 0 < c["inst_count"].sum() < int(cnt[nonempty].sum())     # some items were dropped, the rest were emitted
        ctx.arena_free(small_ws); ctx.arena_free(guard)
    finally:
        ds.close()


def test_tier_x_bad_range_index_is_reported_not_followed(ctx):
    sc = synth.random_scene(23, n=400, num_lists=40)
    sc.cull = sc.cull.copy()
    bad = np.arange(0, sc.n, 7)
    sc.cull[bad, 10] = 0x7FFFFFF0                      # far outside the region table
    planes, eye = synth.orbit_camera(5, 250.0, far=500.0)
    ds = DeviceScene(ctx, sc)
    try:
        ds.record_drawable_processing()
        ds.cull(planes, eye); ctx.sync(ds.stream)
        got = ds.read_tier_x()
        assert got["status"] == 4
        # everything else is processed as if the bad drawables had empty lists
        ok = synth.random_scene(23, n=400, num_lists=40)
        ok.ml_count = ok.ml_count.copy()
        keep = np.ones(sc.n, bool); keep[bad] = False
        emitted = np.unique(got["tag"][np.concatenate([np.arange(int(sc.regions[s, 0]), int(sc.regions[s, 0]) + int(got["cmd_count"][s]))
                                                       for s in range(sc.num_state_sets)]).astype(int), 0]) if got["cmd_count"].sum() else np.array([])
        assert not np.intersect1d(emitted, bad).size
        _, _, ref = oracle_tier_x(ok, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        ref_keep = sum(len(e[6]) for lst in canonicalise(ref).values() for e in lst if keep[e[0]])
        assert int(got["inst_count"].sum()) == ref_keep
    finally:
        ds.close()


def test_upload_scatter_and_patch(ctx):
    rng = np.random.default_rng(9)
    size = 3 << 20
    arena = ctx.arena_alloc(size)
    host = np.zeros(size, np.uint8)
    staging = rng.integers(0, 256, 2 << 20, dtype=np.uint8)
    try:
        ctx.memset(arena, 0, size)
        # ragged regions: 16-B aligned starts (allocator guarantee), arbitrary sizes, one > 256 KiB (own DMA),
        # one misaligned pair (generic path), one empty
        regs, dst, src = [], 64, 0
        for b in [1, 15, 16, 17, 100, 4096, 33000, 70001, 300_000, 0, 5]:
            regs.append((arena + dst, src, b))
            dst = (dst + b + 15) & ~15
            src = (src + b + 15) & ~15
        regs.append((arena + dst + 3, src + 5, 1000))
        regions = np.array(regs, np.uint64)
        ctx.upload(regions, staging)
        ctx.sync()
        got = np.empty(size, np.uint8)
        ctx.memcpy_d2h(got, arena); ctx.sync()
        ob.upload(ob.Memory([(arena, host)]), regions, staging)
        assert np.array_equal(got, host)
        # device-resident staging: the scatter kernel alone
        stage_dev = ctx.arena_alloc(staging.nbytes)
        ctx.memcpy_h2d(stage_dev, staging[::-1].copy())
        ctx.scatter_copy(regions, stage_dev)
        ctx.memcpy_d2h(got, arena); ctx.sync()
        host2 = host.copy()
        ob.upload(ob.Memory([(arena, host2)]), regions, staging[::-1].copy())
        assert np.array_equal(got, host2)
        ctx.arena_free(stage_dev)
        # sparse sources (host-packed path) and a region above 1 MiB (its own DMA) in one call
        big = rng.integers(0, 256, 24 << 20, dtype=np.uint8)
        sparse = np.array([[arena + 4096 * i, (1 << 20) * i + 16 * i, 700 + i] for i in range(20)] +
                          [[arena + (1 << 20), 5 << 20, (1 << 20) + 12345]], np.uint64)
        ctx.upload(sparse, big)
        ctx.memcpy_d2h(got, arena); ctx.sync()
        ob.upload(ob.Memory([(arena, host2)]), sparse, big)
        assert np.array_equal(got, host2)
    finally:
        ctx.arena_free(arena)


def test_upload_with_absolute_sources_in_several_staging_blocks(ctx):
    """stagingBase == NULL: every region names its source by absolute host address (what the facade's DataMemory does,
    DataMemory.cpp:417-425 with mapped StagingMemory blocks).  Sources that are dense inside EACH block but lie in
    different pinned blocks must not be shipped as one span across the gap between the blocks (regression: invalid
    argument from cudaMemcpyAsync after the facade had grown a second staging block); a source outside every known block
    is packed."""
    import ctypes
    nblk, blk = 3, 1 << 20
    ptrs = [ctx.host_alloc(blk) for _ in range(nblk)]
    loose = np.random.default_rng(1).integers(0, 256, 4096, dtype=np.uint8)       # not from host_alloc
    arena = ctx.arena_alloc(8 << 20)
    try:
        ctx.memset(arena, 0, 8 << 20)
        rng = np.random.default_rng(2)
        host = np.zeros(8 << 20, np.uint8)
        regs, dst = [], 0
        for k, p in enumerate(ptrs):
            buf = np.ctypeslib.as_array((ctypes.c_uint8 * blk).from_address(p))
            buf[:] = rng.integers(0, 256, blk, dtype=np.uint8)
            for j in range(40):                                # dense: 40 x 20 000 B of 1 MiB, contiguous
                regs.append((arena + dst, p + 20_000 * j + 16, 19_000 + 16 * j))
                host[dst:dst + 19_000 + 16 * j] = buf[20_000 * j + 16:20_000 * j + 16 + 19_000 + 16 * j]
                dst += 20_480
        regs.append((arena + dst, loose.ctypes.data + 100, 3000))
        host[dst:dst + 3000] = loose[100:3100]
        ctx.upload(np.array(regs, np.uint64), 0)
        got = np.empty(8 << 20, np.uint8)
        ctx.memcpy_d2h(got, arena); ctx.sync()
        assert np.array_equal(got, host)
    finally:
        ctx.arena_free(arena)
        for p in ptrs:
            ctx.host_free(p)


@pytest.mark.parametrize("kw", [dict(seed=31), dict(seed=32, first_handle=1990), dict(seed=33, first_handle=4_194_250)],
                         ids=["L1", "L2", "L3"])
def test_patch_handles_matches_oracle(ctx, kw):
    sc = synth.random_scene(**kw)
    ds = DeviceScene(ctx, sc)
    try:
        img = sc.image(ds.arena)
        hs = np.unique(sc.drawables[:, 2])[:5]
        patches = np.stack([hs, np.uint64(ds.arena) + sc.ml_off[::-1][:len(hs)]], axis=1).astype(np.uint64)
        ctx.patch_handles(ds.root, sc.handle_level, [(int(h), int(a)) for h, a in patches])
        ds.record_drawable_processing(); ctx.sync(ds.stream)
        ind, ptr = ds.read_tier_r()
        mem = ob.Memory([(ds.arena, img), (ds.drawable_list, np.ascontiguousarray(sc.drawables))])
        ob.patch_handles(mem, ds.root, sc.handle_level, patches)
        e_ind, e_ptr = ob.process_drawables(mem, ds.root, sc.handle_level, ds.drawable_list, sc.n)
        assert np.array_equal(ind, e_ind) and np.array_equal(ptr, e_ptr)
    finally:
        ds.close()


@pytest.mark.parametrize("kw,frame", X_CASES[:4], ids=lambda v: f"seed{v['seed']}" if isinstance(v, dict) else f"f{v}")
def test_fused_process_and_cull_equals_two_calls(ctx, kw, frame):
    """cadr_b200_process_and_cull == cadr_b200_process_drawables + cadr_b200_cull_compact, output for output."""
    sc = synth.random_scene(**kw)
    planes, eye = synth.orbit_camera(frame, 250.0, far=500.0)
    ds = DeviceScene(ctx, sc)
    try:
        ds.upload_drawable_list()
        ctx.memset(ds.indirect, 0xEE, sc.n * 16); ctx.memset(ds.pointers, 0xEE, sc.n * 32)
        ds.process_and_cull(planes, eye)
        ctx.sync(ds.stream)
        ind, ptr = ds.read_tier_r()
        got = ds.read_tier_x()
        _, e_ind, e_ptr = _oracle_r(ds)
        assert np.array_equal(ind, e_ind) and np.array_equal(ptr, e_ptr)      # Tier R records written by the fused pass
        _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        assert_tier_x_equal(got, ref)
    finally:
        ds.close()


def test_fused_baseline_shapes(ctx):
    for sc, far in ((synth.config2(150_001), 1500.0), (synth.config3(300, 1000, state_sets=16), 3000.0)):
        planes, eye = synth.orbit_camera(40, 1500.0, far=far)
        ds = DeviceScene(ctx, sc)
        try:
            ds.upload_drawable_list()
            ds.process_and_cull(planes, eye)
            ctx.sync(ds.stream)
            ind, ptr = ds.read_tier_r()
            got = ds.read_tier_x()
            _, e_ind, e_ptr = _oracle_r(ds)
            assert np.array_equal(ind, e_ind) and np.array_equal(ptr, e_ptr)
            _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
            assert_tier_x_equal(got, ref)
        finally:
            ds.close()


@pytest.mark.parametrize("kw", [dict(seed=41), dict(seed=42, first_handle=4_194_250, big_lists=2, n=350)], ids=["L1", "L3"])
def test_consumer_side_walk_matches_oracle(ctx, kw):
    """the emitted buffers are consumable — walk them like the reference's vertex shader does."""
    sc = synth.random_scene(valid_geometry=True, max_count=30, **kw)
    planes, eye = synth.orbit_camera(15, 250.0, far=500.0)
    ds = DeviceScene(ctx, sc)
    digest = ctx.arena_alloc(16)
    try:
        ds.record_drawable_processing()
        ds.cull(planes, eye)
        out = np.zeros(2, np.uint64)
        img = sc.image(ds.arena)
        mem = ob.Memory([(ds.arena, img), (ds.drawable_list, np.ascontiguousarray(sc.drawables))])
        # Tier R: every drawable drawn with vkCmdDrawIndirect
        ctx.consume_check(ds.indirect, ds.pointers, 0, sc.n, digest)
        ctx.memcpy_d2h(out, digest); ctx.sync(ds.stream); ctx.sync()
        ind, ptr = ds.read_tier_r()
        exp = ob.consume_check(mem, ind, ptr, 0, sc.n)
        assert (int(out[0]), int(out[1])) == exp and exp[1] > 10_000
        # Tier X: every draw range drawn with vkCmdDrawIndexedIndirectCount; the GPU result is walked on the GPU, the
        # oracle's own (differently ordered, unsplit) result on the CPU: the digests must agree
        _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        p = ds.cull_params(planes, eye)
        total = 0
        for s in range(sc.num_state_sets):
            ctx.consume_check_culled(p, s, int(sc.regions[s, 1]), digest)
            ctx.memcpy_d2h(out, digest); ctx.sync()
            e = ob.consume_check_culled(mem, ref, s)
            assert (int(out[0]), int(out[1])) == e, f"range {s}"
            total += e[1]
        assert total > 1000
    finally:
        ctx.arena_free(digest)
        ds.close()


@pytest.mark.parametrize("shape", ["c3", "c3-shard", "c2", "c1"])
def test_consumer_side_walk_on_the_baseline_shapes(ctx, shape):
    """The BASELINE shapes carry REAL box geometry (8 corners, 36 / 24 / 12 indices per LOD), so what bench.py's
    multi-GPU verification walks at full size is walkable: consumer walk over every range of a small instance of each
    shape == the oracle's digest; with a non-zero addressDelta on pointers moved by the opposite amount the digest is
    unchanged (the translation a renderer GPU applies to another rank's records)."""
    if shape == "c3":
        sc, cam = synth.config3(300, 40, state_sets=8), synth.orbit_camera(30, 1500.0, far=3000.0)
    elif shape == "c3-shard":
        sc, cam = synth.config3_shard(900, 300, 300, 40, state_sets=8), synth.orbit_camera(30, 1500.0, far=3000.0)
    elif shape == "c2":
        sc, cam = synth.config2(5000), synth.orbit_camera(100, 1500.0, far=1500.0)
    else:
        sc, cam = synth.config1(12), synth.reference_camera(1)
    planes, eye = cam
    ds = DeviceScene(ctx, sc)
    digest = ctx.arena_alloc(16)
    try:
        ds.upload_drawable_list()
        ds.process_and_cull(planes, eye)
        ctx.sync(ds.stream)
        mem = ob.Memory([(ds.arena, sc.image(ds.arena)), (ds.drawable_list, np.ascontiguousarray(sc.drawables))])
        _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        p = ds.cull_params(planes, eye)
        out, total = np.zeros(2, np.uint64), 0
        for s in range(sc.num_state_sets):
            if not int(sc.regions[s, 1]):
                continue
            ctx.consume_check_culled(p, s, int(sc.regions[s, 1]), digest)
            ctx.memcpy_d2h(out, digest); ctx.sync()
            e = ob.consume_check_culled(mem, ref, s)
            assert (int(out[0]), int(out[1])) == e, f"range {s}"
            total += e[1]
        assert total > 1000
        # address translation: shift every forwarded pointer by -delta, walk with +delta
        n_cmd = ds.cmd_cap
        ptr = ds._read(ds.ptr_out, n_cmd * 32, np.uint64).copy()
        delta = 0x10000
        live = ptr != 0
        ptr[live] -= np.uint64(delta)
        ctx.memcpy_h2d(ds.ptr_out, ptr); ctx.sync()
        p.addressDelta = delta
        s = int(np.nonzero(sc.regions[:, 1])[0][0])
        ctx.consume_check_culled(p, s, int(sc.regions[s, 1]), digest)
        ctx.memcpy_d2h(out, digest); ctx.sync()
        assert (int(out[0]), int(out[1])) == ob.consume_check_culled(mem, ref, s)
    finally:
        ctx.arena_free(digest)
        ds.close()


def test_two_phase_upload_stages_on_a_copy_stream_and_commits_in_order(ctx):
    """cadr_b200_upload_stage / _commit against the oracle's copy regions (DataMemory::recordUploads, DataMemory.cpp:400-446):
    nothing reaches a destination before the commit; two staged uploads may be outstanding (two slots), a third is refused;
    commits apply in the order they are issued (the later one wins where regions overlap); regions above 1 MiB go
    through the staging slot as well; a ticket cannot be committed twice; other calls of the upload family keep working
    while one upload is staged."""
    import torch
    rng = np.random.default_rng(19)
    size = 6 << 20
    arena = ctx.arena_alloc(size)
    copy_stream = torch.cuda.Stream()
    try:
        ctx.memset(arena, 0, size); ctx.sync()
        host = np.zeros(size, np.uint8)
        st_a = rng.integers(0, 256, 4 << 20, dtype=np.uint8)
        st_b = rng.integers(0, 256, 4 << 20, dtype=np.uint8)
        regs_a = np.array([[arena + 64 + 4096 * i, 3000 * i, 2500 + i] for i in range(300)] + [[arena + (3 << 20), 1 << 20, (1 << 20) + 777]], np.uint64)
        regs_b = np.array([[arena + 64 + 4096 * i + 1024, 5000 * i + 16, 3000] for i in range(200)], np.uint64)     # overlaps regs_a's ranges
        ta = ctx.upload_stage(regs_a, st_a, copy_stream.cuda_stream)
        tb = ctx.upload_stage(regs_b, st_b, copy_stream.cuda_stream)
        assert ta and tb and ta != tb
        with pytest.raises(cadr_b200.LogicError):
            ctx.upload_stage(regs_b, st_b, copy_stream.cuda_stream)          # both slots hold staged uploads
        copy_stream.synchronize(); ctx.sync()
        got = np.empty(size, np.uint8)
        ctx.memcpy_d2h(got, arena); ctx.sync()
        assert not got.any(), "staging touched a destination"
        ctx.upload_commit(ta)
        # one slot is free again: the single-call forms work while tb is still staged
        small = np.array([[arena + (5 << 20), 64, 4000]], np.uint64)
        ctx.upload(small, st_a)
        ctx.upload_commit(tb)
        ctx.sync()
        ctx.memcpy_d2h(got, arena); ctx.sync()
        mem = ob.Memory([(arena, host)])
        ob.upload(mem, regs_a, st_a); ob.upload(mem, small, st_a); ob.upload(mem, regs_b, st_b)
        assert np.array_equal(got, host)
        with pytest.raises(cadr_b200.LogicError):
            ctx.upload_commit(tb)
        assert ctx.upload_stage(np.zeros((0, 3), np.uint64), st_a, copy_stream.cuda_stream) == 0
        ctx.upload_commit(0)
    finally:
        ctx.sync()
        ctx.arena_free(arena)


@pytest.mark.parametrize("fused", [False, True], ids=["two-calls", "fused"])
def test_primitive_set_offsets_need_only_four_byte_alignment(ctx, fused):
    """PrimitiveSetRef is declared buffer_reference_align = 4 (processDrawables.comp:29-33): a primitiveSetOffset /
    lodPrimitiveSetOffset that is a multiple of 4 but not of 8 is legal.  Round 1 read the PrimitiveSets of QUEUED lists
    (more than 32 matrices) with one 8-byte load and would have faulted here; short, medium and long lists all take
    such offsets now (advisor finding)."""
    counts = [1, 7, 32, 33, 50, 64, 65, 200, 1024, 1500]
    sc = synth.random_scene(71, n=60, num_geometries=5, list_counts=counts, state_sets=3)
    P = sc.gen["build"]["geometries"]
    per_geom = np.array([g["primitive_sets"].shape[0] for g in P])
    room = per_geom[sc.drawable_geom] * 8                              # bytes of each drawable's PrimitiveSet array
    # odd multiples of 4 wherever 8 bytes still fit inside the array
    off = sc.cull[:, 5:8].astype(np.int64)
    new = np.where(off + 4 + 8 <= room[:, None], off + 4, off)
    sc.cull[:, 5:8] = new.astype(np.uint32)
    rec_off = (sc.drawables[:, 5] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    rec_new = np.where(rec_off + 4 + 8 <= room, rec_off + 4, rec_off)
    sc.drawables[:, 5] = rec_new.astype(np.uint64)
    assert (new % 8 == 4).any() and (rec_new % 8 == 4).any()
    ds = DeviceScene(ctx, sc)
    try:
        planes, eye = synth.orbit_camera(40, 250.0, far=500.0)
        if fused:
            ds.upload_drawable_list(); ds.process_and_cull(planes, eye)
        else:
            ds.record_drawable_processing(); ds.cull(planes, eye)
        ctx.sync(ds.stream)
        got = ds.read_tier_x()
        ind, ptr = ds.read_tier_r()
        e_ind, e_ptr, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        assert np.array_equal(ind, e_ind) and np.array_equal(ptr, e_ptr)
        assert_tier_x_equal(got, ref)
        assert got["inst_count"].sum() > 0
    finally:
        ds.close()


def test_back_to_back_frames_read_complete_counters(ctx):
    """Stream order across the programmatic dependent launch: cullMediumKernel is launched behind cullListWarpKernel with
    programmatic stream serialisation and leaves at once when there are no medium lists (the C3 shape) - without
    griddepcontrol.wait in it, its completion would not imply the list kernel's, and the counters D2H / the next frame's
    counters memset queued behind it could overtake a list kernel that is still running (round-1 verdict and advisor).
    400 frames queued back to back without host synchronisation, cameras alternating so that consecutive frames differ,
    every frame's counters copied to the host right behind its kernels: each copy must hold the complete totals of its
    frame (known from synchronised runs of the two cameras)."""
    sc = synth.config3(3000, 1000, state_sets=16)             # long lists only: the medium kernel has nothing to do
    ds = DeviceScene(ctx, sc)
    try:
        cams = [synth.orbit_camera(30, 1500.0, far=3000.0), synth.orbit_camera(200, 1500.0, far=3000.0)]
        ds.upload_drawable_list()
        expect = []
        for planes, eye in cams:
            ds.process_and_cull(planes, eye)
            ctx.sync(ds.stream)
            c = ds.read_counters()
            assert c["status"] == 0 and c["medium_count"] == 0 and c["inst_count"].sum() > 0
            expect.append(ds._read(ds.counters, ds.counters_bytes, np.uint8).copy())
        assert not np.array_equal(expect[0][64:], expect[1][64:])
        frames = 400
        host = ctx.host_alloc(frames * ds.counters_bytes)
        try:
            for f in range(frames):
                planes, eye = cams[f & 1]
                ds.process_and_cull(planes, eye)
                ctx.memcpy_d2h(host + f * ds.counters_bytes, ds.counters, ds.counters_bytes, stream=ds.stream)
            ctx.sync(ds.stream)
            import ctypes
            got = np.ctypeslib.as_array((ctypes.c_uint8 * (frames * ds.counters_bytes)).from_address(host)).reshape(frames, -1)
            keep = np.r_[0:12, 16:20, 64:ds.counters_bytes]      # status, near-band count, queued items, queued medium lists, per-range totals
            for f in range(frames):
                assert np.array_equal(got[f][keep], expect[f & 1][keep]), f"frame {f}: counters were read before the frame's kernels had finished"
        finally:
            ctx.host_free(host)
    finally:
        ds.close()
