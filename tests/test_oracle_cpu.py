"""CPU suite: the oracle (oracle/cadr_oracle.c) against independent numpy expectations derived from the
scene DESCRIPTION (not from the tables the oracle walks), at all three handle levels, across the level
transitions (2047->2048 and 4194303->4194304 handles) and on the edge cases the domain has."""
import numpy as np
import pytest

from cadr_b200 import synth
from cadr_b200.frame import canonicalise
from oracle import binding as ob
from helpers import FAKE_BASE, FAKE_LIST, fold_by_drawable_lod, oracle_tier_r, oracle_tier_x

CASES = [
    dict(seed=1),                                           # level 1
    dict(seed=2, first_handle=1990),                        # straddles 2047 -> 2048: level 2
    dict(seed=3, first_handle=3000, force_level=3),         # few handles, three levels
    dict(seed=4, first_handle=4_194_250, big_lists=2),      # straddles 4194303 -> 4194304: level 3
    dict(seed=5, n=1, num_lists=3),
    dict(seed=6, n=257, with_drawable_data=False),
]


def expected_tier_r(sc: synth.Scene, base: int, img: np.ndarray):
    """What processDrawables.comp must emit, computed from the scene DESCRIPTION (block offsets per geometry /
    matrix list / drawable) without walking any handle table."""
    n = sc.n
    dg = sc.drawable_geom
    ps_addr = sc.geo_off[dg, 2].astype(np.int64) + (sc.drawables[:, 5] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    u32 = img.view(np.uint32)
    ind = np.zeros((n, 4), np.uint32)
    ind[:, 0] = u32[ps_addr // 4]
    ind[:, 1] = sc.ml_count[sc.drawable_ml]
    ind[:, 2] = u32[ps_addr // 4 + 1]
    ptr = np.zeros((n, 4), np.uint64)
    ptr[:, 0] = np.uint64(base) + sc.geo_off[dg, 0]
    ptr[:, 1] = np.uint64(base) + sc.geo_off[dg, 1]
    ptr[:, 2] = np.uint64(base) + sc.ml_off[sc.drawable_ml]
    ptr[:, 3] = np.where(sc.dd_off != 0, np.uint64(base) + sc.dd_off, np.uint64(0))
    return ind, ptr


@pytest.mark.parametrize("kw", CASES, ids=lambda k: f"seed{k['seed']}")
def test_tier_r_against_scene_description(kw):
    sc = synth.random_scene(**kw)
    img = sc.image(FAKE_BASE)
    _, ind, ptr = oracle_tier_r(sc, img=img)
    e_ind, e_ptr = expected_tier_r(sc, FAKE_BASE, img)
    assert np.array_equal(ind, e_ind)
    assert np.array_equal(ptr, e_ptr)
    assert (ptr[sc.drawables[:, 3] == 0, 3] == 0).all()    # handle 0 resolves to the zero slot


def test_levels_match_handle_ranges():
    assert synth.random_scene(1).handle_level == 1
    assert synth.random_scene(2, first_handle=1990).handle_level == 2
    assert synth.random_scene(4, first_handle=4_194_250).handle_level == 3


def test_primitive_set_values_follow_offset():
    geos = [dict(vertices=np.arange(96, dtype=np.uint8), indices=np.arange(144, dtype=np.uint8),
                 primitive_sets=np.array([[36, 0], [24, 36], [12, 60]], np.uint32))]
    n = 3
    sc = synth.build_scene("ps", geometries=geos, ml_count=np.array([2, 5], np.uint32), drawable_geom=np.zeros(n, int),
                           drawable_ml=np.array([0, 1, 1]), drawable_ps_offset=np.array([0, 8, 16], np.uint64),
                           state_set=np.zeros(n, np.uint32), sphere=np.zeros((n, 4), np.float32),
                           lod_count=np.ones(n, np.uint32), lod_ps_offset=np.zeros((n, 3), np.uint32),
                           lod_threshold=np.zeros((n, 2), np.float32),
                           matrices=np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (7, 1)))
    _, ind, ptr = oracle_tier_r(sc)
    assert ind.tolist() == [[36, 2, 0, 0], [24, 5, 36, 0], [12, 5, 60, 0]]
    assert ptr[1, 2] == ptr[2, 2] != ptr[0, 2]


def test_oracle_reports_faults_instead_of_emulating_ub():
    sc = synth.random_scene(7)
    img = sc.image(FAKE_BASE)
    dl = np.ascontiguousarray(sc.drawables).copy()
    dl[0, 2] = 1_000_000  # matrixListHandle far outside a level-1 table
    mem = ob.Memory([(FAKE_BASE, img), (FAKE_LIST, dl)])
    with pytest.raises(RuntimeError):
        ob.process_drawables(mem, FAKE_BASE + sc.root_off, sc.handle_level, FAKE_LIST, sc.n)


def _fma(a, b, c):
    """fusedMultiplyAdd in numpy: the float32 x float32 product is exact in float64; the sum is rounded to 53 bits
    and then to 24 (a double rounding that can differ from a true fma only on exact float32 midpoints)."""
    return (a.astype(np.float64) * np.float64(b) + np.asarray(c, dtype=np.float64)).astype(np.float32) if np.isscalar(b) or np.ndim(b) == 0 \
        else (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, dtype=np.float64)).astype(np.float32)


def numpy_lod(sc: synth.Scene, planes, eye):
    """Independent evaluation of the Tier X spec (vectorised numpy, fused multiply-adds emulated in float64).
    -> per-drawable list of int8 arrays: lod per instance or -1."""
    starts = np.concatenate([[0], np.cumsum(sc.ml_count.astype(np.int64))])
    f = np.float32
    out = []
    for d in range(sc.n):
        k = int(sc.drawable_ml[d])
        M = sc.matrices[starts[k]:starts[k + 1]]
        cd = sc.cull[d]
        bs = cd[0:4].view(np.float32)
        lodc = min(max(int(cd[4]), 1), 3)
        thr = cd[8:10].view(np.float32)
        c = [_fma(M[:, 8 + a], bs[2], _fma(M[:, 4 + a], bs[1], _fma(M[:, 0 + a], bs[0], M[:, 12 + a]))) for a in range(3)]
        s = [_fma(M[:, 4 * q + 2], M[:, 4 * q + 2], _fma(M[:, 4 * q + 1], M[:, 4 * q + 1], M[:, 4 * q] * M[:, 4 * q])) for q in range(3)]
        smax = np.maximum(np.maximum(s[0], s[1]), s[2])
        with np.errstate(invalid="ignore"):
            r = np.sqrt(smax) * bs[3]
            vis = np.full(M.shape[0], bs[3] >= 0)
            for p in planes:
                dot = _fma(c[2], p[2], _fma(c[1], p[1], _fma(c[0], p[0], np.full(M.shape[0], p[3], np.float32))))
                vis &= dot >= -r
        dx, dy, dz = c[0] - f(eye[0]), c[1] - f(eye[1]), c[2] - f(eye[2])
        dist = np.sqrt(_fma(dz, dz, _fma(dy, dy, dx * dx)))
        lod = np.zeros(M.shape[0], np.int8)
        if lodc > 1:
            lod += thr[0] <= dist
        if lodc > 2:
            lod += thr[1] <= dist
        out.append(np.where(vis, lod, -1).astype(np.int8))
    return out


@pytest.mark.parametrize("kw,frame", [(dict(seed=11), 0), (dict(seed=12, big_lists=3), 40),
                                      (dict(seed=13, first_handle=4_194_250, big_lists=1), 200)],
                         ids=["small", "big", "level3"])
def test_tier_x_oracle_against_numpy(kw, frame):
    sc = synth.random_scene(**kw)
    planes, eye = synth.orbit_camera(frame, 250.0, far=500.0)
    ind, ptr, res = oracle_tier_x(sc, planes, eye)
    lods = numpy_lod(sc, planes, eye)
    canon = canonicalise(res)
    got = {}
    for s, lst in canon.items():
        for (d, lod, count, first, voff, p, inst) in lst:
            got[(d, lod)] = inst
            assert int(sc.cull[d, 10]) == s and voff == 0
            assert p == tuple(int(x) for x in ptr[d])
    exp = {}
    for d, l in enumerate(lods):
        for lod in range(3):
            idx = np.nonzero(l == lod)[0].astype(np.uint32)
            if idx.size:
                exp[(d, lod)] = idx
    assert got.keys() == exp.keys()
    for key in exp:
        assert np.array_equal(got[key], exp[key]), key
    assert res["num_instances"] == sum(v.size for v in exp.values()) > 0
    assert res["num_instances"] < sc.total_instances  # the camera really culls something


def test_near_band_counts_only_instances_whose_classification_could_flip():
    """The near-band definition (Tier X): |min_k(dot_k + r)| < 1e-5, or a visible instance on an LOD threshold.
    Unit spheres against the planes x >= 0 and z <= 10 (the others far away): touching x = 0 counts, touching x = 0
    while far beyond z = 10 does not, clearly in / out do not; a visible instance exactly at threshold distance counts."""
    pos = np.array([[-1.0, 0, 0],            # touches x >= 0 from outside, inside everything else      -> near, visible (dot == -r)
                    [-1.0 - 4e-6, 0, 0],     # 4e-6 outside                                               -> near, invisible
                    [-1.0 + 4e-6, 5, 0],     # 4e-6 inside, 5.1 from the eye                              -> near, visible, lod 1
                    [-1.0, 0, 500.0],        # touches x = 0 but 489 units beyond z <= 10                 -> NOT near
                    [-1.5, 0, 0],            # clearly outside                                            -> not near
                    [3.0, 0, 0],             # clearly inside, at distance 3 = threshold                  -> near (threshold), lod 1
                    [3.0, 0, 500.0]],        # at the threshold distance from nothing visible             -> not near
                   np.float32)
    sc = synth.random_scene(90, n=1, list_counts=[len(pos)], state_sets=1, with_drawable_data=False)
    m = np.zeros((len(pos), 16), np.float32)
    m[:, 0] = m[:, 5] = m[:, 10] = m[:, 15] = 1.0
    m[:, 12:15] = pos
    sc.matrices[:] = m
    sc.cull[:, 0:3] = 0
    sc.cull[:, 3] = np.float32(1.0).view(np.uint32)
    sc.cull[:, 4] = 2                                                     # two LODs, threshold 3.0
    sc.cull[:, 8] = np.float32(3.0).view(np.uint32)
    big = 1e9
    planes = np.array([[1, 0, 0, 0], [-1, 0, 0, big], [0, 1, 0, big], [0, -1, 0, big], [0, 0, 1, big], [0, 0, -1, 10]], np.float32)
    eye = np.zeros(3, np.float32)
    _, _, res = oracle_tier_x(sc, planes, eye)
    assert res["near_band"] == 4
    canon = canonicalise(res)[0]
    assert {(d, lod): inst.tolist() for (d, lod, *_x, inst) in canon} == {(0, 0): [0], (0, 1): [2, 5]}


def test_tier_x_everything_and_nothing_visible():
    sc = synth.random_scene(21, big_lists=1)
    big = 1e9
    all_in = np.array([[1, 0, 0, big], [-1, 0, 0, big], [0, 1, 0, big], [0, -1, 0, big], [0, 0, 1, big], [0, 0, -1, big]], np.float32)
    none = all_in.copy(); none[0, 3] = -big
    eye = np.zeros(3, np.float32)
    _, _, r_all = oracle_tier_x(sc, all_in, eye)
    _, _, r_none = oracle_tier_x(sc, none, eye)
    cnt = sc.ml_count[sc.drawable_ml].astype(np.int64)
    nonempty = sc.cull[:, 3].view(np.float32) >= 0
    assert r_all["num_instances"] == int(cnt[nonempty].sum())   # empty spheres never survive
    assert r_none["num_instances"] == 0 and r_none["num_commands"] == 0


def test_upload_and_patch_restatements():
    rng = np.random.default_rng(5)
    arena = np.zeros(1 << 16, np.uint8)
    staging = rng.integers(0, 256, 1 << 15, dtype=np.uint8)
    regions = np.array([[FAKE_BASE + 64, 0, 100], [FAKE_BASE + 4096, 512, 1], [FAKE_BASE + 8192, 1024, 20000], [FAKE_BASE, 0, 0]], np.uint64)
    mem = ob.Memory([(FAKE_BASE, arena)])
    ob.upload(mem, regions, staging)
    got = mem.arrays[0]
    assert np.array_equal(got[64:164], staging[:100]) and got[4096] == staging[512]
    assert np.array_equal(got[8192:28192], staging[1024:21024]) and got[164:4096].sum() == 0
    with pytest.raises(RuntimeError):
        ob.upload(mem, np.array([[FAKE_BASE + (1 << 16) - 8, 0, 16]], np.uint64), staging)
    sc = synth.random_scene(31, first_handle=4_194_250)
    img = sc.image(FAKE_BASE)
    mem = ob.Memory([(FAKE_BASE, img), (FAKE_LIST, np.ascontiguousarray(sc.drawables))])
    h = int(sc.drawables[0, 2])
    ob.patch_handles(mem, FAKE_BASE + sc.root_off, sc.handle_level, np.array([[h, FAKE_BASE + 4096]], np.uint64))
    _, ptr = ob.process_drawables(mem, FAKE_BASE + sc.root_off, sc.handle_level, FAKE_LIST, sc.n)
    assert (ptr[sc.drawables[:, 2] == np.uint64(h), 2] == FAKE_BASE + 4096).all()


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_parallel_summary_equals_the_emitted_result(seed):
    """oracle_cull_summary (the whole-scene parity entry: OpenMP, per (drawable, lod) count / index sum / sum of
    squares) against the sequential oracle_cull_compact folded the same way, on ragged scenes with long lists, in one
    call and chunk by chunk over arbitrary subsets of the drawables (how the full-size GPU tests feed it)."""
    sc = synth.random_scene(seed, n=700, num_lists=120, max_count=90, state_sets=6, big_lists=3)
    planes, eye = synth.orbit_camera(40 * seed, 250.0, far=500.0)
    img = sc.image(FAKE_BASE)
    mem, ind, ptr = oracle_tier_r(sc, img=img)
    ref = ob.cull_compact(mem, FAKE_BASE + sc.root_off, sc.handle_level, FAKE_LIST, sc.n, ind, ptr, sc.cull, planes, eye, sc.regions)
    k, sm, sq = fold_by_drawable_lod(ref, sc.n)
    for threads in (1, 4):
        got = ob.cull_summary(mem, ind, ptr, sc.cull, planes, eye, threads)
        assert np.array_equal(got["k"].astype(np.uint64), k) and np.array_equal(got["sum"], sm) and np.array_equal(got["sq"], sq)
        assert got["near_band"] == ref["near_band"] and int(got["k"].sum()) == ref["num_instances"]
    rng = np.random.default_rng(seed)
    order = rng.permutation(sc.n)
    nb = 0
    for part in np.array_split(order, 5):
        got = ob.cull_summary(mem, ind[part], ptr[part], sc.cull[part], planes, eye, 3)
        assert np.array_equal(got["k"].astype(np.uint64), k[part]) and np.array_equal(got["sum"], sm[part]) and np.array_equal(got["sq"], sq[part])
        nb += got["near_band"]
    assert nb == ref["near_band"]
