"""GPU (ONE device is enough): the multi-GPU exchange path with both "ranks" living on the same GPU.

The kernels of the fused exchange only see device addresses — whether exchangeCmd[r] is a peer mapping over NVLink or
plain local memory makes no difference to them.  Two slices of ONE scene are culled one after the other on the same
device, each storing its records into the gathered arrays of BOTH ranks (writeCommandRecord's exchange branch),
publishing its counters and flag (publishKernel); waitPeersKernel then passes on both flag arrays; the survivors' index
runs are pulled into a gathered index buffer (pullInstancesKernel) and every rank's result is walked from the other
rank's view (cadr_b200_consume_check_culled through the gathered arrays).  Everything is compared with the oracle: per
slice (pointers, commands), merged against the WHOLE scene, and through the consumer digest.  tests/multigpu_check.py
runs the same checks with one process per GPU over real peer mappings; this file is what a one-GPU box can verify."""
import numpy as np
import pytest

import cadr_b200
from cadr_b200 import _capi, shard, synth
from cadr_b200.frame import DeviceScene
from oracle import binding as ob
from helpers import fold_by_drawable_lod, oracle_tier_r, oracle_tier_x

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_fused_exchange_with_all_ranks_on_one_gpu(ctx, world):
    whole = synth.random_scene(5151, n=1400, num_geometries=8, num_lists=150, max_count=90, state_sets=6, big_lists=3, valid_geometry=True)
    slices = shard.partition(whole.ml_count[whole.drawable_ml], world)
    scenes = [synth.slice_scene(whole, f, c) for f, c in slices]
    dss = [DeviceScene(ctx, sc) for sc in scenes]
    S = whole.num_state_sets
    cmd_cap = max(ds.cmd_cap for ds in dss)
    inst_cap = (max(ds.inst_cap for ds in dss) + 3) & ~3
    cb = ctx.cull_counters_bytes(S)
    owned = []

    def new(n):
        a = ctx.arena_alloc(max(n, 256)); owned.append(a); return a

    # per "rank": gathered arrays, flags, a device copy of every rank's region table
    gathered = [dict(cmd=new(world * cmd_cap * 20), ptr=new(world * cmd_cap * 32), tag=new(world * cmd_cap * 8), counters=new(world * cb)) for _ in range(world)]
    flags = [new(256) for _ in range(world)]
    regions_dev = []
    for r in range(world):
        reg = np.zeros((S, 4), np.uint32); reg[:scenes[r].regions.shape[0]] = scenes[r].regions
        a = new(reg.nbytes); ctx.memcpy_h2d(a, reg); regions_dev.append(a)
    gathered_inst = new(world * inst_cap * 4)
    digests = new(16 * S)
    for f in flags:
        ctx.memset(f, 0, 256)
    ctx.sync()
    try:
        planes, eye = synth.orbit_camera(130, 250.0, far=500.0)
        syncs = []
        for r, ds in enumerate(dss):
            ds.upload_drawable_list()
            p = ds.cull_params(planes, eye)
            p.exchangeWorld, p.exchangeRank, p.exchangeCmdCapacity = world, r, cmd_cap
            for q in range(world):
                p.exchangeCmd[q], p.exchangePtr[q], p.exchangeTag[q] = gathered[q]["cmd"], gathered[q]["ptr"], gathered[q]["tag"]
            ctx.process_and_cull(p, stream=ds.stream)
            s = _capi.ExchangeSync()
            s.world, s.rank, s.frameSeq, s.localCounters, s.countersBytes = world, r, 1, ds.counters, cb
            for q in range(world):
                s.peerCounters[q], s.peerFlags[q] = gathered[q]["counters"], flags[q]
            ctx.exchange_publish(s, ds.stream)
            syncs.append(s)
            ctx.sync(ds.stream)
        for r, ds in enumerate(dss):              # every rank has published: the stream-side wait passes on every flag array
            ctx.exchange_wait(syncs[r], ds.stream)
            ctx.sync(ds.stream)

        refs = [oracle_tier_x(scenes[r], planes, eye, arena_base=dss[r].arena, list_base=dss[r].drawable_list) for r in range(world)]
        _, _, ref_whole = oracle_tier_x(whole, planes, eye)
        K, Sm, Q = fold_by_drawable_lod(ref_whole, whole.n)
        for viewer in range(world):
            g = {}
            for name, dt, width in (("cmd", np.uint32, 5), ("ptr", np.uint64, 4), ("tag", np.uint32, 2)):
                buf = np.empty(world * cmd_cap * width * np.dtype(dt).itemsize, np.uint8)
                ctx.memcpy_d2h(buf, gathered[viewer][name]); ctx.sync()
                g[name] = buf.view(dt).reshape(-1, width)
            craw = np.empty(world * cb, np.uint8)
            ctx.memcpy_d2h(craw, gathered[viewer]["counters"]); ctx.sync()
            counts = craw.reshape(world, cb)[:, 64:].copy().view(np.uint64)
            assert not craw.reshape(world, cb)[:, :4].any(), "status"
            k = np.zeros((whole.n, 3), np.uint64); sm = np.zeros_like(k); q2 = np.zeros_like(k)
            for r in range(world):
                ref = refs[r][2]
                assert np.array_equal((counts[r] >> np.uint64(32)).astype(np.int64), ref["inst_count"])
                inst = dss[r]._read(dss[r].inst_out, dss[r].inst_cap * 4, np.uint32)
                exp = {}
                for s_ in range(S):
                    b = int(scenes[r].regions[s_, 0])
                    for ci in range(b, b + int(ref["cmd_count"][s_])):
                        exp[(s_, int(ref["tag"][ci, 0]), int(ref["tag"][ci, 1]))] = (int(ref["cmd"][ci, 0]), int(ref["cmd"][ci, 1]), int(ref["cmd"][ci, 2]), tuple(int(x) for x in ref["ptr"][ci]))
                got = {}
                for s_ in range(S):
                    b = r * cmd_cap + int(scenes[r].regions[s_, 0])
                    for ci in range(b, b + int(counts[r][s_] & np.uint64(0xFFFFFFFF))):
                        key = (s_, int(g["tag"][ci, 0]), int(g["tag"][ci, 1]))
                        prev = got.get(key)
                        cnt, first = int(g["cmd"][ci, 1]), int(g["cmd"][ci, 4])
                        got[key] = (int(g["cmd"][ci, 0]), cnt + (prev[1] if prev else 0), int(g["cmd"][ci, 2]), tuple(int(x) for x in g["ptr"][ci]))
                        d, lod = slices[r][0] + key[1], key[2]
                        run = inst[first:first + cnt].astype(np.uint64)
                        k[d, lod] += np.uint64(cnt); sm[d, lod] += run.sum(dtype=np.uint64); q2[d, lod] += (run * run).sum(dtype=np.uint64)
                assert got == exp, f"viewer {viewer}: commands of rank {r}"
            assert np.array_equal(k, K) and np.array_equal(sm, Sm) and np.array_equal(q2, Q), "merged result != whole-scene oracle"

        # pull every rank's survivor runs into one index buffer, then walk every rank's result through the gathered views
        pull = _capi.ExchangePull()
        pull.world, pull.rank, pull.numRanges, pull.countersBytes = world, 0, S, cb
        pull.gatheredCounters, pull.gatheredInst, pull.instCapacity, pull.includeLocal = gathered[0]["counters"], gathered_inst, inst_cap, 1
        for r in range(world):
            pull.regions[r], pull.peerInst[r] = regions_dev[r], dss[r].inst_out
        ctx.exchange_pull_instances(pull)
        ctx.sync()
        for r in range(world):
            ref = refs[r][2]
            pulled = np.empty(inst_cap, np.uint32)
            ctx.memcpy_d2h(pulled, gathered_inst + 4 * r * inst_cap); ctx.sync()
            own = dss[r]._read(dss[r].inst_out, dss[r].inst_cap * 4, np.uint32)
            for s_ in range(S):
                ib, ni = int(scenes[r].regions[s_, 2]), int(ref["inst_count"][s_])
                assert np.array_equal(pulled[ib:ib + ni], own[ib:ib + ni]), f"pulled runs of rank {r} StateSet {s_}"
            mem = ob.Memory([(dss[r].arena, scenes[r].image(dss[r].arena)), (dss[r].drawable_list, np.ascontiguousarray(scenes[r].drawables))])
            for inst_src in (dss[r].inst_out, gathered_inst + 4 * r * inst_cap):
                for viewer in range(world):
                    p = _capi.CullParams()
                    p.numStateSets = S
                    p.cmdOut, p.ptrOut, p.tagOut = (gathered[viewer]["cmd"] + 20 * r * cmd_cap, gathered[viewer]["ptr"] + 32 * r * cmd_cap,
                                                    gathered[viewer]["tag"] + 8 * r * cmd_cap)
                    p.counters, p.stateSetRegions, p.instOut, p.addressDelta = gathered[viewer]["counters"] + r * cb, regions_dev[r], inst_src, 0
                    for s_ in range(S):
                        if not int(scenes[r].regions[s_, 1]):
                            continue
                        ctx.consume_check_culled(p, s_, int(scenes[r].regions[s_, 1]), digests + 16 * s_)
                        out = np.zeros(2, np.uint64)
                        ctx.memcpy_d2h(out, digests + 16 * s_); ctx.sync()
                        assert (int(out[0]), int(out[1])) == ob.consume_check_culled(mem, ref, s_), f"rank {r} range {s_} seen from {viewer}"
    finally:
        ctx.sync()
        for a in owned:
            ctx.arena_free(a)
        for ds in dss:
            ds.close()


def test_pull_rejects_bad_arguments(ctx):
    pull = _capi.ExchangePull()
    pull.world, pull.rank, pull.numRanges, pull.countersBytes = 2, 5, 4, 96
    with pytest.raises(cadr_b200.LogicError):
        ctx.exchange_pull_instances(pull)
    pull.rank = 0
    with pytest.raises(cadr_b200.LogicError):
        ctx.exchange_pull_instances(pull)          # buffers missing


def test_publish_and_wait_as_one_launch(ctx):
    """cadr_b200_exchange_publish_and_wait (the one-launch frame close PeerExchange uses with one process per GPU): with a
    world of one the kernel publishes to its own gathered slot and its wait passes on its own flag, so a one-GPU box can
    check the launch, the copy and the flag without a second process.  (With more ranks the kernel spins until every peer
    has published - exercised over real peer mappings by tests/multigpu_check.py and bench.py --gpus N.)"""
    sc = synth.random_scene(77, n=300, num_lists=40, max_count=70, state_sets=3, big_lists=1)
    ds = DeviceScene(ctx, sc)
    cb = ctx.cull_counters_bytes(sc.num_state_sets)
    gathered, flags = ctx.arena_alloc(max(cb, 256)), ctx.arena_alloc(256)
    try:
        ctx.memset(flags, 0, 256); ctx.memset(gathered, 0xEE, cb); ctx.sync()
        ds.upload_drawable_list()
        for seq in (1, 2):
            planes, eye = synth.orbit_camera(40 * seq, 250.0, far=500.0)
            ds.process_and_cull(planes, eye)
            s = _capi.ExchangeSync()
            s.world, s.rank, s.frameSeq, s.localCounters, s.countersBytes = 1, 0, seq, ds.counters, cb
            s.peerCounters[0], s.peerFlags[0] = gathered, flags
            ctx.exchange_publish_and_wait(s, ds.stream)
            ctx.sync(ds.stream)
            local, got, flag = np.empty(cb, np.uint8), np.empty(cb, np.uint8), np.zeros(1, np.uint64)
            ctx.memcpy_d2h(local, ds.counters); ctx.memcpy_d2h(got, gathered); ctx.memcpy_d2h(flag, flags); ctx.sync()
            assert np.array_equal(local, got) and int(flag[0]) == seq
            _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
            assert np.array_equal((got[64:].view(np.uint64) >> np.uint64(32)).astype(np.int64), ref["inst_count"])
        bad = _capi.ExchangeSync()
        bad.world, bad.rank, bad.frameSeq, bad.localCounters, bad.countersBytes = 1, 0, 3, 0, cb
        bad.peerCounters[0], bad.peerFlags[0] = gathered, flags
        with pytest.raises(cadr_b200.LogicError):
            ctx.exchange_publish_and_wait(bad, ds.stream)
    finally:
        ctx.sync()
        ds.close()
        ctx.arena_free(gathered); ctx.arena_free(flags)


def test_a_peer_that_never_publishes_ends_a_bounded_wait(ctx):
    """cadr_exchange_sync.timeoutMs: the stream-side waits give up when their budget is spent and report it in the status
    word of the rank's own counters (the device-side form of the reference's fence wait with a timeout -> CadR::Timeout,
    Renderer.cpp:982-993).  World of two on one device; "rank 1" never runs."""
    import time
    cb = ctx.cull_counters_bytes(2)
    local, gathered0, gathered1, flags0, flags1 = (ctx.arena_alloc(256) for _ in range(5))
    try:
        for a in (local, gathered0, gathered1, flags0, flags1):
            ctx.memset(a, 0, 256)
        ctx.sync()
        for seq, call in ((1, ctx.exchange_publish_and_wait), (2, ctx.exchange_wait)):
            s = _capi.ExchangeSync()
            s.world, s.rank, s.frameSeq, s.localCounters, s.countersBytes, s.timeoutMs = 2, 0, seq, local, cb, 40
            s.peerCounters[0], s.peerCounters[1], s.peerFlags[0], s.peerFlags[1] = gathered0, gathered1, flags0, flags1
            ctx.memset(local, 0, 256); ctx.sync()
            t0 = time.monotonic()
            call(s)
            ctx.sync()
            dt = time.monotonic() - t0
            status = np.zeros(1, np.uint32)
            ctx.memcpy_d2h(status, local); ctx.sync()
            assert int(status[0]) & _capi.CULL_STATUS_EXCHANGE_TIMEOUT, "the expiry is reported in the status word"
            assert 0.03 < dt < 5.0, f"waited {dt:.3f} s for a 40 ms budget"
        # the peer's slot of rank 0's flags was published by the first call (rank 0 -> every rank), its own wait is what expired
        f = np.zeros(2, np.uint64)
        ctx.memcpy_d2h(f, flags1); ctx.sync()
        assert int(f[0]) == 1
        # a wait that is satisfied does not touch the status
        ctx.memset(local, 0, 256)
        one = np.array([5, 5], np.uint64)
        ctx.memcpy_h2d(flags0, one); ctx.sync()
        s = _capi.ExchangeSync()
        s.world, s.rank, s.frameSeq, s.localCounters, s.countersBytes, s.timeoutMs = 2, 0, 5, local, cb, 40
        s.peerCounters[0], s.peerCounters[1], s.peerFlags[0], s.peerFlags[1] = gathered0, gathered1, flags0, flags1
        ctx.exchange_wait(s); ctx.sync()
        status = np.zeros(1, np.uint32)
        ctx.memcpy_d2h(status, local); ctx.sync()
        assert int(status[0]) == 0
    finally:
        ctx.sync()
        for a in (local, gathered0, gathered1, flags0, flags1):
            ctx.arena_free(a)
