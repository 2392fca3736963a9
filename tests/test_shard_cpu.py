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
