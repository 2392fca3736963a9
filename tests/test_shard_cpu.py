"""CPU suite for the N>1 path: drawable partitioning and the command-list exchange, world_size 2 over gloo.
Each rank evaluates its shard with the oracle (test infrastructure), the product's Exchange gathers the compacted
lists, and the merged result must equal the single-process result on the whole scene."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cadr_b200 import shard, synth  # noqa: E402


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 8):
        for counts in (rng.integers(0, 2000, 5000), np.ones(1000, int), np.array([10**6] + [1] * 999), np.zeros(17, int)):
            parts = shard.partition(counts, world)
            assert len(parts) == world and parts[0][0] == 0
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert parts[-1][0] + parts[-1][1] == len(counts)
            w = np.maximum(counts, 1)
            loads = [int(w[f:f + c].sum()) for f, c in parts]
            assert max(loads) - w.sum() / world <= w.max()      # within one drawable of the ideal share


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_scene(sc: synth.Scene, first: int, count: int):
    """The slice [first, first+count) of a scene's flattened drawable list, with its own per-range regions."""
    import copy
    sub = copy.copy(sc)
    sub.drawables = sc.drawables[first:first + count].copy()
    sub.cull = sc.cull[first:first + count].copy()
    sub.drawable_ml = sc.drawable_ml[first:first + count].copy()
    return sub


def _worker(rank, world, port, q):
    from helpers import oracle_tier_x
    from cadr_b200.frame import canonicalise
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = synth.random_scene(77, n=600, num_lists=120, max_count=50, state_sets=6, big_lists=3)
        planes, eye = synth.orbit_camera(25, 250.0, far=500.0)
        counts = sc.ml_count[sc.drawable_ml]
        first, count = shard.partition(counts, world)[rank]
        sub = _shard_scene(sc, first, count)
        _, _, res = oracle_tier_x(sub, planes, eye)        # this rank's compacted lists (regions of the full scene: ample)
        ex = shard.Exchange(res["cmd"].shape[0], sc.num_state_sets, torch.device("cpu"))
        counters = np.zeros(64 + 8 * sc.num_state_sets, np.uint8)
        counters[64:] = (res["cmd_count"].astype(np.uint64) | (res["inst_count"].astype(np.uint64) << np.uint64(32))).view(np.uint8)
        ex.run(*(torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)) for a in (res["cmd"], res["ptr"], res["tag"])),
               torch.from_numpy(counters))
        directory = ex.directory([sc.regions] * world)
        g = ex.commands()
        firsts = [p[0] for p in shard.partition(counts, world)]
        # merged view: per StateSet, (global drawable, lod) -> (indexCount, instanceCount, firstIndex, pointers)
        merged = {}
        for e in directory:
            for ci in range(e["first_command"], e["first_command"] + e["count"]):
                key = (e["range"], int(g["tag"][ci, 0]) + firsts[e["rank"]], int(g["tag"][ci, 1]))
                merged[key] = (int(g["cmd"][ci, 0]), int(g["cmd"][ci, 1]), int(g["cmd"][ci, 2]), tuple(int(x) for x in g["ptr"][ci]))
        if rank == 0:
            _, _, whole = oracle_tier_x(sc, planes, eye)
            exp = {}
            for s, lst in canonicalise(whole).items():
                for (d, lod, ic, fi, vo, p, inst) in lst:
                    exp[(s, d, lod)] = (ic, len(inst), fi, p)
            q.put(("ok" if merged == exp and len(exp) > 50 else f"mismatch: {len(merged)} vs {len(exp)}", ex.bytes_per_rank))
    finally:
        dist.destroy_process_group()


def test_exchange_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    status, nbytes = q.get(timeout=5)
    assert status == "ok", status
    assert nbytes > 0


def _tier_r_worker(rank, world, port, q):
    from helpers import oracle_tier_r
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = synth.random_scene(78, n=700, num_lists=90, max_count=40, state_sets=5, first_handle=1990)
        counts = sc.ml_count[sc.drawable_ml]
        slices = shard.partition(counts, world)
        first, count = slices[rank]
        _, ind, ptr = oracle_tier_r(_shard_scene(sc, first, count))      # this rank's slice (stand-in for its GPU pass)
        g = shard.TierRGather(slices, torch.device("cpu"))
        g.run(torch.from_numpy(np.ascontiguousarray(ind).view(np.uint8).reshape(-1)), torch.from_numpy(np.ascontiguousarray(ptr).view(np.uint8).reshape(-1)))
        got_ind, got_ptr = g.records()
        _, whole_ind, whole_ptr = oracle_tier_r(sc)
        ok = np.array_equal(got_ind, whole_ind) and np.array_equal(got_ptr, whole_ptr) and g.n == sc.n
        ok = ok and [g.owner(f) for f, c in slices if c] == [r for r, (f, c) in enumerate(slices) if c]
        q.put((rank, "ok" if ok else "mismatch"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tier_r_gather_equals_single_pass_gloo(world):
    """Every rank resolves its slice of the drawable list; after TierRGather every rank holds the indirect / pointers
    arrays of the WHOLE list, bit-identical to one pass over it (slices of unequal length, no padding)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tier_r_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(r, "ok") for r in range(world)]


def test_tier_r_gather_rejects_bad_slices():
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        with pytest.raises(ValueError):
            shard.TierRGather([(0, 5), (6, 2)], torch.device("cpu"))          # two slices for one rank
        with pytest.raises(ValueError):
            shard.TierRGather([(1, 5)], torch.device("cpu"))                  # does not start at 0
        g = shard.TierRGather([(0, 3)], torch.device("cpu"))
        with pytest.raises(ValueError):
            g.run(torch.zeros(16, dtype=torch.uint8), torch.zeros(96, dtype=torch.uint8))   # indirect too short
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_slices_of_one_scene_add_up_to_the_whole_scene(world):
    """synth.slice_scene (what every rank of the one-scene multi-GPU path lays out): per-rank slices cut by
    shard.partition, each with its own arena and handle table over local objects only, global StateSet indices.  Resolved
    and culled by the oracle slice by slice, the results put back at their global drawable positions are exactly the
    whole scene's: Tier R indirect records, per-(drawable, lod) survivor sets, per-StateSet totals, near-band count."""
    from helpers import fold_by_drawable_lod, oracle_tier_r, oracle_tier_x
    whole = synth.random_scene(4242, n=1100, num_geometries=9, num_lists=140, max_count=80, state_sets=7, big_lists=3, valid_geometry=True)
    planes, eye = synth.orbit_camera(75, 250.0, far=500.0)
    w_ind, _, ref = oracle_tier_x(whole, planes, eye)
    K, S, Q = fold_by_drawable_lod(ref, whole.n)
    slices = shard.partition(whole.ml_count[whole.drawable_ml], world)
    inst = np.zeros(whole.num_state_sets, np.int64)
    near = 0
    for f, c in slices:
        sc = synth.slice_scene(whole, f, c)
        assert sc.n == c and sc.num_state_sets == whole.num_state_sets and sc.gen["first"] == f
        assert sc.arena_bytes < whole.arena_bytes or world == 1
        ind, _, r = oracle_tier_x(sc, planes, eye)
        assert np.array_equal(ind, w_ind[f:f + c])
        k, s, q = fold_by_drawable_lod(r, c)
        assert np.array_equal(k, K[f:f + c]) and np.array_equal(s, S[f:f + c]) and np.array_equal(q, Q[f:f + c])
        assert r["status"] == 0
        absent = np.setdiff1d(np.arange(whole.num_state_sets), np.unique(sc.cull[:, 10]))
        assert not sc.regions[absent, 1].any() and not sc.regions[absent, 3].any()        # StateSets without local drawables: empty regions
        inst += r["inst_count"]; near += r["near_band"]
    assert np.array_equal(inst, ref["inst_count"]) and near == ref["near_band"]


def test_config3_shards_are_parts_of_one_config3_scene():
    """synth.config3_shard (the big device-synthesised shape of BASELINE configs[4]): matrices, StateSets and culling
    records of every shard are those of config3() at the same global positions, whichever way the list is cut."""
    n, inst, S = 1900, 24, 16
    whole = synth.config3(n, inst, state_sets=S)
    planes, eye = synth.orbit_camera(30, 1500.0, far=3000.0)
    from helpers import fold_by_drawable_lod, oracle_tier_x
    _, _, ref = oracle_tier_x(whole, planes, eye)
    K, _, Q = fold_by_drawable_lod(ref, n)
    for world in (2, 5):
        tot = 0
        for f, c in shard.partition(np.full(n, inst), world):
            sc = synth.config3_shard(n, f, c, inst, state_sets=S)
            assert np.array_equal(sc.matrices.reshape(c, inst, 16), whole.matrices.reshape(n, inst, 16)[whole.drawable_ml[f:f + c]])
            assert np.array_equal(sc.cull, whole.cull[f:f + c]) and sc.num_state_sets == S
            _, _, r = oracle_tier_x(sc, planes, eye)
            k, _, q = fold_by_drawable_lod(r, c)
            assert np.array_equal(k, K[f:f + c]) and np.array_equal(q, Q[f:f + c])
            tot += r["num_instances"]
        assert tot == ref["num_instances"] and tot > 0
