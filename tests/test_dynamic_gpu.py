"""GPU: BASELINE configs[3] — the dynamic scene: 100 M instances, 10 % of the MatrixLists rewritten per frame through
the upload path (cadr_b200_upload from pinned staging), then processed and culled.  Two variants of the rewrite:
in place (same device addresses) and realloc-on-write (new addresses + cadr_b200_patch_handles), as DataAllocation::alloc
does (DataAllocation.cpp:18-34)."""
import ctypes

import numpy as np
import pytest
import torch

from cadr_b200 import synth
from cadr_b200.synth import trs_matrices
from test_fullsize_gpu import build, gpu_summary, sample_check

pytestmark = pytest.mark.gpu


def new_matrices(rng, count):
    pos = (rng.random((count, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(3000)
    scale = np.float32(0.5) + np.float32(1.5) * rng.random(count, dtype=np.float32)
    return trs_matrices(pos, None, scale)


def test_c4_rewrite_ten_percent_then_cull(ctx):
    n, inst = 100_000, 1000
    scene = synth.config3(n, inst, state_sets=64, host_matrices=False)
    ds, arena, stream = build(ctx, scene)
    rewritten = 10_000
    staging_bytes = rewritten * inst * 64
    staging_ptr = ctx.host_alloc(staging_bytes)
    spare = ctx.arena_alloc(1000 * (64 + inst * 64))
    try:
        staging = np.ctypeslib.as_array((ctypes.c_uint8 * staging_bytes).from_address(staging_ptr))
        rng = np.random.default_rng(42)
        a = arena.tensor(ds.arena)
        with torch.cuda.stream(stream):
            ds.record_drawable_processing()
            for frame in (1, 2):
                lists = (frame * 10007 + np.arange(rewritten, dtype=np.int64) * 7919) % n          # Appendix D cfg 4
                assert len(np.unique(lists)) == rewritten
                staging.view(np.float32).reshape(-1, 16)[:] = new_matrices(rng, rewritten * inst)
                # in-place rewrite of the matrices (the 64-byte headers stay): one region per list
                regions = np.stack([np.uint64(ds.arena) + scene.ml_off[lists] + np.uint64(64),
                                    np.arange(rewritten, dtype=np.uint64) * np.uint64(inst * 64),
                                    np.full(rewritten, inst * 64, np.uint64)], axis=1)
                ctx.upload(regions, staging_ptr, stream=stream.cuda_stream)
                planes, eye = synth.orbit_camera(17 * frame, 1500.0, far=3000.0)
                ds.process_and_cull(planes, eye)
                stream.synchronize()
                # the device now holds exactly the staged bytes at the rewritten lists
                for k in (0, 1, rewritten // 2, rewritten - 1):
                    off = int(scene.ml_off[lists[k]]) + 64
                    assert np.array_equal(a[off:off + inst * 64].cpu().numpy(), staging[k * inst * 64:(k + 1) * inst * 64])
                dev_sum = int(torch.stack([a[int(scene.ml_off[l]) + 64:int(scene.ml_off[l]) + 64 + inst * 64].view(torch.int32).to(torch.int64).sum()
                                           for l in lists[::97]]).sum())
                host_sum = int(sum(staging[k * inst * 64:(k + 1) * inst * 64].view(np.int32).astype(np.int64).sum() for k in range(0, rewritten, 97)))
                assert dev_sum == host_sum
                # and the cull of the new frame equals the oracle on rewritten + untouched drawables
                s = gpu_summary(ds, arena, scene)
                assert s["status"] == 0
                d_of_list = np.argsort(scene.drawable_ml)       # drawable that uses list l
                sample = np.unique(np.concatenate([d_of_list[lists[:20]], rng.integers(0, n, 20)]))
                assert sample_check(ctx, ds, arena, scene, planes, eye, s, sample) > 0

            # realloc-on-write variant: 1000 lists get NEW device ranges and their handles are patched in place
            moved = (3 * 10007 + np.arange(1000, dtype=np.int64) * 7919) % n
            blk = 64 + inst * 64
            stage2 = staging[:1000 * blk]
            hdr = np.zeros(64, np.uint8); hdr[:8] = np.array([inst, inst], np.uint32).view(np.uint8)
            s2 = stage2.reshape(1000, blk)
            s2[:, :64] = hdr
            s2[:, 64:].view(np.float32).reshape(-1, 16)[:] = new_matrices(rng, 1000 * inst)
            new_addr = np.uint64(spare) + np.arange(1000, dtype=np.uint64) * np.uint64(blk)
            ctx.upload(np.stack([new_addr, np.arange(1000, dtype=np.uint64) * np.uint64(blk), np.full(1000, blk, np.uint64)], axis=1),
                       staging_ptr, stream=stream.cuda_stream)
            d_of_list = np.argsort(scene.drawable_ml)
            handles = scene.drawables[d_of_list[moved], 2]
            ctx.patch_handles(ds.root, scene.handle_level, list(zip(handles.tolist(), new_addr.tolist())), stream=stream.cuda_stream)
            planes, eye = synth.orbit_camera(99, 1500.0, far=3000.0)
            ds.process_and_cull(planes, eye)
            stream.synchronize()
            ptr = arena.tensor(ds.pointers).view(torch.int64)[:n * 4].view(-1, 4)
            got = ptr[torch.from_numpy(d_of_list[moved]).cuda(), 2].cpu().numpy().astype(np.uint64)
            assert np.array_equal(got, new_addr)                 # Tier R sees the relocated lists through the patched table
            untouched = np.setdiff1d(np.arange(n), d_of_list[moved])[:1000]
            exp = (np.uint64(ds.arena) + scene.ml_off[scene.drawable_ml[untouched]]).astype(np.uint64)
            assert np.array_equal(ptr[torch.from_numpy(untouched).cuda(), 2].cpu().numpy().astype(np.uint64), exp)
    finally:
        ctx.arena_free(spare)
        ctx.host_free(staging_ptr)
        ds.close()
