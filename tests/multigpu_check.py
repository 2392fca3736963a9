"""Run under torchrun with >= 2 ranks (one per GPU):  the fused peer-memory exchange against the oracle.
Every rank culls its own shard; the kernels store the compacted commands into every rank's gathered arrays over
NVLink; each rank then checks ALL ranks' slices of its gathered arrays against the oracle run on those shards."""
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
from helpers import oracle_tier_r, oracle_tier_x  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cadr_b200.Context(local)
    scenes = [synth.random_scene(900 + r, n=500 + 37 * r, num_lists=90, max_count=60, state_sets=4 + r, big_lists=2) for r in range(world)]
    sc = scenes[rank]
    ds = DeviceScene(ctx, sc)
    bases = [None] * world
    dist.all_gather_object(bases, (ds.arena, ds.drawable_list))
    px = shard.PeerExchange(ctx, ds.cmd_cap, sc.num_state_sets)
    ok = True
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
        assert (g["status"] == 0).all()
        for r in range(world):
            _, _, ref = oracle_tier_x(scenes[r], planes, eye, arena_base=bases[r][0], list_base=bases[r][1])
            cnt = g["counts"][r][:scenes[r].num_state_sets]
            # instance counts must match exactly; command counts may be higher on the GPU (a list longer than 1024
            # matrices is emitted as one command per work item) and are compared after merging below
            if not np.array_equal((cnt >> np.uint64(32)).astype(np.int64), ref["inst_count"]):
                ok = False
                print(f"rank {rank}: counters of rank {r} differ in frame {frame}")
                continue
            exp = {}
            for s in range(scenes[r].num_state_sets):
                b = int(scenes[r].regions[s, 0])
                for ci in range(b, b + int(ref["cmd_count"][s])):
                    key = (s, int(ref["tag"][ci, 0]), int(ref["tag"][ci, 1]))
                    exp[key] = (int(ref["cmd"][ci, 0]), int(ref["cmd"][ci, 1]), int(ref["cmd"][ci, 2]), tuple(int(x) for x in ref["ptr"][ci]))
            got = {}
            for s in range(scenes[r].num_state_sets):
                b = r * px.cmd_cap + int(scenes[r].regions[s, 0])
                for ci in range(b, b + int(cnt[s] & np.uint64(0xFFFFFFFF))):
                    key = (s, int(g["tag"][ci, 0]), int(g["tag"][ci, 1]))
                    prev = got.get(key)
                    k = int(g["cmd"][ci, 1]) + (prev[1] if prev else 0)      # long lists: one command per work item
                    got[key] = (int(g["cmd"][ci, 0]), k, int(g["cmd"][ci, 2]), tuple(int(x) for x in g["ptr"][ci]))
            if got != exp:
                ok = False
                print(f"rank {rank}: commands of rank {r} differ in frame {frame} ({len(got)} vs {len(exp)})")
        dist.barrier()
    if os.environ.get("MG_TIER_R") == "1":
        # opt-in until it has run on NCCL once: the fixed-size Tier R records of every rank's slice gathered into whole-list
        # arrays (shard.TierRGather); here every rank's "slice" is its own scene, the oracle resolves each of them
        slices, first = [], 0
        for r in range(world):
            slices.append((first, scenes[r].n)); first += scenes[r].n
        ind, ptr = ds.read_tier_r()
        tg = shard.TierRGather(slices, torch.device("cuda", local))
        tg.run(torch.from_numpy(np.ascontiguousarray(ind).view(np.uint8).reshape(-1)).cuda(), torch.from_numpy(np.ascontiguousarray(ptr).view(np.uint8).reshape(-1)).cuda())
        g_ind, g_ptr = tg.records()
        for r, (f0, c) in enumerate(slices):
            _, e_ind, e_ptr = oracle_tier_r(scenes[r], arena_base=bases[r][0], list_base=bases[r][1])
            if not (np.array_equal(g_ind[f0:f0 + c], e_ind) and np.array_equal(g_ptr[f0:f0 + c], e_ptr)):
                ok = False
                print(f"rank {rank}: gathered Tier R records of rank {r} differ")
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    px.close()
    ds.close()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("multigpu_check", "ok" if int(t.item()) == 1 else "FAILED", f"({world} ranks, {frames} frames{', queued without host sync' if run_async else ''})")
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
