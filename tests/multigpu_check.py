"""Run under torchrun with >= 2 ranks (one per GPU): the multi-GPU path against the oracle.

  A  independent scenes    every rank culls a scene of its own (different sizes, StateSet counts, capacities); the kernels
                           store the compacted commands into every rank's gathered arrays over NVLink; each rank checks
                           ALL ranks' slices of its gathered arrays against the oracle run on those scenes.
  B  ONE scene, partitioned (BASELINE configs[4])   a ragged scene with real geometry is cut by
                           shard.partition into per-rank slices (synth.slice_scene: global StateSet indices, local handle
                           tables).  Every rank then holds the whole frame in a form it can CONSUME:
                             * merged over the ranks, the gathered commands + the instance indices read through the
                               peer-mapped index buffers equal the oracle's result for the WHOLE scene, StateSet by StateSet;
                             * the walk of the reference's vertex shader (shader.vert:99-123, cadr_b200_consume_check_culled)
                               over every rank's ranges - indices and matrices fetched from the owning GPU through the peer
                               mappings, addresses translated by addressDelta - gives the oracle's digest;
                             * the NCCL cross-check of the fused exchange (PeerExchange.verify) agrees;
                             * the optional second stage (instance-index runs pulled to the renderer GPU) delivers the
                               owners' runs, and the consumer walk over the pulled copies gives the same digests;
                             * a renderer issues at most S + world - 1 indirect-count draws (directory()).
  D  deferred wait         PeerExchange(sets=4, deferred_wait=True): a rank waits for the peers' PREVIOUS frame before it
                           publishes its own, frames queued back to back with a different camera each; after every
                           end_frame the gathered result of frame k-1 - and after finish() the last frame's - equals the oracle.
  C  Tier R gather (NCCL)  the fixed-size records of the processing pass, every rank's slice broadcast into whole-list
                           arrays (shard.TierRGather): equal to the oracle's records of the slices, in list order.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cadr_b200  # noqa: E402
from cadr_b200 import shard, synth  # noqa: E402
from cadr_b200.frame import DeviceScene  # noqa: E402
from helpers import fold_by_drawable_lod, oracle_tier_r, oracle_tier_x  # noqa: E402
from oracle import binding as ob  # noqa: E402

problems = []


def fail(msg):
    problems.append(msg)
    print(f"rank {os.environ.get('RANK')}: {msg}", flush=True)


def check_gathered_commands(rank, g, px, scenes, refs, frame):
    """Every rank's slot of the gathered arrays against that rank's oracle result (pointers included)."""
    for r, (sc, ref) in enumerate(zip(scenes, refs)):
        cnt = g["counts"][r][:sc.num_state_sets]
        # instance counts must match exactly; command counts may be higher on the GPU (a list longer than 1024 matrices is
        # emitted as one command per work item) and are compared after merging below
        if not np.array_equal((cnt >> np.uint64(32)).astype(np.int64), ref["inst_count"]):
            fail(f"counters of rank {r} differ in frame {frame}")
            continue
        exp = {}
        for s in range(sc.num_state_sets):
            b = int(sc.regions[s, 0])
            for ci in range(b, b + int(ref["cmd_count"][s])):
                key = (s, int(ref["tag"][ci, 0]), int(ref["tag"][ci, 1]))
                exp[key] = (int(ref["cmd"][ci, 0]), int(ref["cmd"][ci, 1]), int(ref["cmd"][ci, 2]), tuple(int(x) for x in ref["ptr"][ci]))
        got = {}
        for s in range(sc.num_state_sets):
            b = r * px.cmd_cap + int(sc.regions[s, 0])
            for ci in range(b, b + int(cnt[s] & np.uint64(0xFFFFFFFF))):
                key = (s, int(g["tag"][ci, 0]), int(g["tag"][ci, 1]))
                prev = got.get(key)
                k = int(g["cmd"][ci, 1]) + (prev[1] if prev else 0)      # long lists: one command per work item
                got[key] = (int(g["cmd"][ci, 0]), k, int(g["cmd"][ci, 2]), tuple(int(x) for x in g["ptr"][ci]))
        if got != exp:
            fail(f"commands of rank {r} differ in frame {frame} ({len(got)} vs {len(exp)})")


def part_a(ctx, rank, world, local):
    scenes = [synth.random_scene(900 + r, n=500 + 37 * r, num_lists=90, max_count=60, state_sets=4 + r, big_lists=2) for r in range(world)]
    sc = scenes[rank]
    ds = DeviceScene(ctx, sc)
    bases = [None] * world
    dist.all_gather_object(bases, (ds.arena, ds.drawable_list))
    px = shard.PeerExchange(ctx, ds.cmd_cap, sc.num_state_sets)
    # MG_FRAMES frames; with MG_ASYNC=1 they are queued back to back without any host synchronisation (ranks run ahead
    # of each other as far as the stream-side wait allows) and only the last one is checked
    frames, run_async = int(os.environ.get("MG_FRAMES", "5")), os.environ.get("MG_ASYNC") == "1"
    for frame in range(frames):
        planes, eye = synth.orbit_camera(20 * frame, 250.0, far=500.0)
        ds.upload_drawable_list()
        p = ds.cull_params(planes, eye)
        px.begin_frame(p)
        ctx.process_and_cull(p, stream=ds.stream)
        px.end_frame(ds.counters, stream=ds.stream)
        if run_async and frame + 1 < frames:
            continue
        ctx.sync(ds.stream)
        g = px.read()
        if not (g["status"] == 0).all():
            fail(f"status {g['status']} in frame {frame}")
        refs = [oracle_tier_x(scenes[r], planes, eye, arena_base=bases[r][0], list_base=bases[r][1])[2] for r in range(world)]
        check_gathered_commands(rank, g, px, scenes, refs, frame)
        dist.barrier()
    px.close()
    ds.close()
    return frames


def part_b_and_c(ctx, rank, world, local):
    dev = torch.device("cuda", local)
    whole = synth.random_scene(4242, n=900 + 150 * world, num_geometries=9, num_lists=140, max_count=80, state_sets=7, big_lists=3,
                               valid_geometry=True)
    slices = shard.partition(whole.ml_count[whole.drawable_ml], world)
    scenes = [synth.slice_scene(whole, f, c) for f, c in slices]
    sc = scenes[rank]
    ds = DeviceScene(ctx, sc)
    bases = [None] * world
    dist.all_gather_object(bases, (ds.arena, ds.drawable_list))
    inst2 = ctx.arena_alloc(max(ds.inst_cap, 1) * 4)       # instance indices alternate between two buffers by frame parity
    px = shard.PeerExchange(ctx, ds.cmd_cap, sc.num_state_sets, regions=sc.regions, inst_out=[ds.inst_out, inst2], arena=ds.arena,
                            first_drawable=slices[rank][0])
    px.enable_pull(ds.inst_cap)
    S = whole.num_state_sets
    for frame in (3, 11):
        planes, eye = synth.orbit_camera(25 * frame, 250.0, far=500.0)
        ds.upload_drawable_list()
        p = ds.cull_params(planes, eye)
        px.begin_frame(p)
        ctx.process_and_cull(p, stream=ds.stream)
        px.end_frame(ds.counters, stream=ds.stream)
        px.pull_instances(stream=ds.stream, include_local=True)
        ctx.sync(ds.stream)
        dist.barrier()                       # nobody starts the next frame (and rewrites its index buffer) while peers still read
        g = px.read()
        if not (g["status"] == 0).all():
            fail(f"B: status {g['status']}")
        # per-rank oracle (pointers are addresses of the owning GPU) and the oracle on the WHOLE scene
        per_rank = [oracle_tier_x(scenes[r], planes, eye, arena_base=bases[r][0], list_base=bases[r][1]) for r in range(world)]
        check_gathered_commands(rank, g, px, scenes, [x[2] for x in per_rank], f"B{frame}")
        _, _, ref_whole = oracle_tier_x(whole, planes, eye)
        K, Sm, Q = fold_by_drawable_lod(ref_whole, whole.n)
        # merged view: commands from the local gathered arrays, instance indices through the peer mappings and from the pulled copies
        k = np.zeros((whole.n, 3), np.uint64); sm = np.zeros_like(k); q = np.zeros_like(k)
        inst_total = np.zeros(S, np.int64)
        for r in range(world):
            cap_r = int(scenes[r].inst_capacity)
            inst_peer = np.zeros(max(cap_r, 1), np.uint32)
            inst_pulled = np.zeros(max(cap_r, 1), np.uint32)
            if cap_r:
                ctx.memcpy_d2h(inst_peer, px.inst_of(r), cap_r * 4, stream=ds.stream)
                ctx.memcpy_d2h(inst_pulled, px.gathered_inst + 4 * r * px.inst_cap, cap_r * 4, stream=ds.stream)
                ctx.sync(ds.stream)
            for s in range(S):
                c = int(g["counts"][r][s] & np.uint64(0xFFFFFFFF))
                ni = int(g["counts"][r][s] >> np.uint64(32))
                inst_total[s] += ni
                ib = int(scenes[r].regions[s, 2])
                if not np.array_equal(inst_peer[ib:ib + ni], inst_pulled[ib:ib + ni]):
                    fail(f"B{frame}: pulled instance indices of rank {r} StateSet {s} differ from the owner's")
                b = r * px.cmd_cap + int(scenes[r].regions[s, 0])
                for ci in range(b, b + c):
                    d = px.first_drawable[r] + int(g["tag"][ci, 0])
                    lod, cnt, first = int(g["tag"][ci, 1]), int(g["cmd"][ci, 1]), int(g["cmd"][ci, 4])
                    if whole.cull[d, 10] != s:
                        fail(f"B{frame}: command of drawable {d} sits in StateSet {s}")
                    run = inst_peer[first:first + cnt].astype(np.uint64)
                    k[d, lod] += np.uint64(cnt); sm[d, lod] += run.sum(dtype=np.uint64); q[d, lod] += (run * run).sum(dtype=np.uint64)
        if not (np.array_equal(k, K) and np.array_equal(sm, Sm) and np.array_equal(q, Q)):
            fail(f"B{frame}: merged result differs from the oracle's result for the whole scene")
        if not np.array_equal(inst_total, ref_whole["inst_count"]):
            fail(f"B{frame}: per-StateSet survivor totals differ from the whole scene's")
        # consumer walk over EVERY rank's ranges from this GPU: peer-mapped indices + matrices, then the pulled index copies
        for r in range(world):
            mem_r, ind_r, ptr_r = oracle_tier_r(scenes[r], arena_base=bases[r][0], list_base=bases[r][1])
            ref_r = per_rank[r][2]
            e_dig, e_n = 0, 0
            for s in range(S):
                if int(ref_r["cmd_count"][s]):
                    dg, n = ob.consume_check_culled(mem_r, ref_r, s)
                    e_dig = (e_dig + dg) & 0xFFFFFFFFFFFFFFFF; e_n += n
            got = px.consume(r, stream=ds.stream)
            if got != (e_dig, e_n):
                fail(f"B{frame}: consumer walk over rank {r}'s result gives {got}, oracle {(e_dig, e_n)}")
            got2 = px.consume(r, stream=ds.stream, pulled=True)
            if got2 != (e_dig, e_n):
                fail(f"B{frame}: consumer walk over the PULLED indices of rank {r} gives {got2}, oracle {(e_dig, e_n)}")
        v = px.verify(dev)
        if not v["ok"]:
            fail(f"B{frame}: NCCL cross-check of the fused exchange failed: {v['problems']}")
        directory = px.directory()
        if len(directory) > S + world - 1:
            fail(f"B{frame}: {len(directory)} draws for {S} StateSets over {world} ranks")
        dist.barrier()

    # C: Tier R gather over NCCL
    ind, ptr = ds.read_tier_r()
    tg = shard.TierRGather(slices, dev)
    tg.run(torch.from_numpy(np.ascontiguousarray(ind).view(np.uint8).reshape(-1)).to(dev),
           torch.from_numpy(np.ascontiguousarray(ptr).view(np.uint8).reshape(-1)).to(dev))
    g_ind, g_ptr = tg.records()
    for r, (f0, c) in enumerate(slices):
        _, e_ind, e_ptr = oracle_tier_r(scenes[r], arena_base=bases[r][0], list_base=bases[r][1])
        if not (np.array_equal(g_ind[f0:f0 + c], e_ind) and np.array_equal(g_ptr[f0:f0 + c], e_ptr)):
            fail(f"C: gathered Tier R records of rank {r} differ")
    # ... and the indirect records are those of the whole scene (pointers differ: per-rank arenas)
    _, w_ind, _ = oracle_tier_r(whole)
    if not np.array_equal(g_ind, w_ind):
        fail("C: gathered indirect records differ from the whole scene's")
    px.close()
    ctx.arena_free(inst2)
    ds.close()
    return len(directory)


def part_d(ctx, rank, world, local):
    scenes = [synth.random_scene(1300 + r, n=400 + 29 * r, num_lists=70, max_count=60, state_sets=3 + r % 3, big_lists=1) for r in range(world)]
    sc = scenes[rank]
    ds = DeviceScene(ctx, sc)
    bases = [None] * world
    dist.all_gather_object(bases, (ds.arena, ds.drawable_list))
    px = shard.PeerExchange(ctx, ds.cmd_cap, sc.num_state_sets, sets=4, deferred_wait=True)
    cams = [synth.orbit_camera(37 * f, 250.0, far=500.0) for f in range(9)]

    def check(frame_no):          # frame numbers start at 1
        planes, eye = cams[frame_no - 1]
        g = px.read(frame=frame_no)
        refs = [oracle_tier_x(scenes[r], planes, eye, arena_base=bases[r][0], list_base=bases[r][1])[2] for r in range(world)]
        check_gathered_commands(rank, g, px, scenes, refs, f"D{frame_no}")

    ds.upload_drawable_list()
    for f, (planes, eye) in enumerate(cams, start=1):
        p = ds.cull_params(planes, eye)
        px.begin_frame(p)
        ctx.process_and_cull(p, stream=ds.stream)
        px.end_frame(ds.counters, stream=ds.stream)
        if f % 3 == 0:            # every third frame: look at what is complete now (frame f-1); otherwise keep queueing
            ctx.sync(ds.stream)
            if px.complete_frame != f - 1:
                fail("D: complete_frame")
            check(f - 1)
    px.finish(ds.stream)
    ctx.sync(ds.stream)
    check(len(cams))
    dist.barrier()
    px.close()
    ds.close()
    return len(cams)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cadr_b200.Context(local)
    frames = part_a(ctx, rank, world, local)
    draws = part_b_and_c(ctx, rank, world, local)
    deferred_frames = part_d(ctx, rank, world, local)
    t = torch.tensor([0 if problems else 1], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    ok = int(t.item()) == 1
    if rank == 0:
        print("multigpu_check", "ok" if ok else "FAILED",
              f"({world} ranks; A: {frames} frames of independent scenes{', queued without host sync' if os.environ.get('MG_ASYNC') == '1' else ''}; "
              f"B: one scene partitioned, merged == whole-scene oracle, consumer walk through peer mappings == oracle, NCCL cross-check, "
              f"pulled instance runs, {draws} draws; C: Tier R gather over NCCL; D: {deferred_frames} frames with the deferred wait)")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
