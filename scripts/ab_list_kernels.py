#!/usr/bin/env python
"""A/B of the long-list kernel variants over MatrixList lengths in ONE process: the scene of each length is built once,
then every variant (CADR_B200_CULL_VARIANT, read at each launch by the A/B library libcadr_b200_exp.so, which this script loads
instead of the product library) is timed on it, interleaved, with CUDA
events around the frames and the library's per-kernel events.  Every variant's result is compared with variant 2's
(counters and canonicalised commands) before it is timed.
usage: scripts/ab_list_kernels.py [--lengths 33,64,...] [--variants 2,4,5] [--total 100000000] [--steps 30] [--rounds 2]
-> one JSON line per (length, variant) on stdout"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lengths", default="33,64,100,200,1000")
    ap.add_argument("--variants", default="2,4,5")
    ap.add_argument("--total", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--state-sets", type=int, default=64)
    a = ap.parse_args()

    import torch
    from cadr_b200 import _capi, build
    _capi.LIB_PATH = build.build_cuda(experiments=True)     # the product library has one path and reads no environment
    import cadr_b200
    from cadr_b200 import synth
    from cadr_b200.frame import DeviceScene, canon_equal, canonicalise
    from cadr_b200.synth_torch import TorchArena, fill_matrix_lists

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = cadr_b200.Context(0)
    stream_t = torch.cuda.Stream(device=dev)
    stream = stream_t.cuda_stream
    cams = [synth.orbit_camera(k, 1500.0, far=3000.0) for k in range(360)]
    variants = a.variants.split(",")

    for length in (int(x) for x in a.lengths.split(",")):
        n = a.total // length
        scene = synth.config3(n, length, state_sets=a.state_sets, seed=0xC0FFEE03, host_matrices=False)
        arena = TorchArena(dev)
        with torch.cuda.stream(stream_t):
            ds = DeviceScene(ctx, scene, alloc=arena.alloc, free=arena.free, upload=False, stream=stream)
            ds.upload_static(with_matrices=False)
            fill_matrix_lists(scene, arena.tensor(ds.arena))
            ds.record_drawable_processing()
        torch.cuda.synchronize()
        inst = scene.total_instances

        def frame(k):
            planes, eye = cams[k % 360]
            ds.process_and_cull(planes, eye)

        # parity between variants (variant 2 is the one the test-suite pins to the oracle at these sizes)
        ref = None
        check = {}
        with torch.cuda.stream(stream_t):
            for v in ["2"] + [x for x in variants if x != "2"]:
                os.environ["CADR_B200_CULL_VARIANT"] = v
                frame(17)
                ctx.sync(stream)
                got = ds.read_tier_x()
                c = canonicalise(got) if n <= 120_000 else None      # a host loop over every command: long lists only
                if ref is None:
                    ref = (got, c)
                    check[v] = "reference"
                else:
                    same = (np.array_equal(got["inst_count"], ref[0]["inst_count"]) and np.array_equal(got["cmd_count"], ref[0]["cmd_count"])
                            and got["near_band"] == ref[0]["near_band"] and got["status"] == ref[0]["status"])
                    if same and c is not None:
                        same, _ = canon_equal(c, ref[1])
                    check[v] = "equal" if same else "DIFFERENT"

        results = {v: {"frame_ms": [], "small_ms": [], "list_ms": []} for v in variants}
        with torch.cuda.stream(stream_t):
            for r in range(a.rounds):
                for v in variants:
                    os.environ["CADR_B200_CULL_VARIANT"] = v
                    for k in range(3):
                        frame(k)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record(stream_t)
                    for k in range(a.steps):
                        frame(3 + k)
                    e1.record(stream_t)
                    torch.cuda.synchronize()
                    results[v]["frame_ms"].append(e0.elapsed_time(e1) / a.steps)
                    ctx.set_profiling(True)
                    ks = []
                    for k in range(8):
                        frame(3 + k)
                        stream_t.synchronize()
                        ks.append(ctx.kernel_times())
                    ctx.set_profiling(False)
                    ks = np.array(ks)
                    results[v]["small_ms"].append(float(ks[:, 1].mean()))
                    results[v]["list_ms"].append(float(ks[:, 2].mean()))
        for v in variants:
            fm = min(results[v]["frame_ms"])
            print(json.dumps({"matrices_per_list": length, "drawables": n, "variant": v, "parity_vs_variant_2": check.get(v),
                              "frame_ms": round(fm, 4), "G_instances_per_s": round(inst / fm / 1e6, 1),
                              "frame_ms_rounds": [round(x, 4) for x in results[v]["frame_ms"]],
                              "cullSmallKernel_ms": round(min(results[v]["small_ms"]), 4),
                              "list_kernel_ms": round(min(results[v]["list_ms"]), 4)}), flush=True)
        ds.close()
        del ds, arena
        torch.cuda.empty_cache()
    os.environ.pop("CADR_B200_CULL_VARIANT", None)
    ctx.close()


if __name__ == "__main__":
    main()
