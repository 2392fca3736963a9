#!/usr/bin/env python
"""Randomised parity soak on a GPU: random ragged scenes x random cameras x {fused, two calls} x {bounds on, off} against
the oracle, for a time budget.  usage: scripts/fuzz_parity.py [seconds] [first_seed]   (CADR_B200_CULL_VARIANT selects
the list kernel when FUZZ_EXPERIMENTS=1 loads the A/B library libcadr_b200_exp.so).  Prints one line per failure and a summary; exit code 1 if anything differed."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cadr_b200  # noqa: E402
from cadr_b200 import synth  # noqa: E402
if os.environ.get("FUZZ_EXPERIMENTS") == "1":      # the A/B library: CADR_B200_CULL_VARIANT / CADR_B200_SMALL_STAGED select its kernels
    from cadr_b200 import _capi, build
    _capi.LIB_PATH = build.build_cuda(experiments=True)
from cadr_b200.frame import DeviceScene, canon_equal, canonicalise  # noqa: E402
from helpers import oracle_tier_x  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = cadr_b200.Context(0)
t_end = time.time() + budget
runs = fails = 0
seed = seed0
only = [int(x) for x in os.environ.get("FUZZ_SEEDS", "").split(",") if x]      # replay exactly these scenes
while (only or time.time() < t_end):
    if only:
        seed = only.pop(0)
    rng = np.random.default_rng(seed)
    kind = int(rng.integers(0, 4))
    kw = dict(seed=seed, n=int(rng.integers(1, 900)), state_sets=int(rng.integers(1, 9)),
              first_handle=int(rng.choice([1, 1900, 2040, 4_194_000])), with_drawable_data=bool(rng.integers(0, 2)))
    if kind == 0:
        kw.update(num_lists=int(rng.integers(1, 120)), max_count=int(rng.choice([3, 40, 70, 300])), big_lists=int(rng.integers(0, 4)))
    else:
        kw.update(list_counts=[int(c) for c in rng.choice([0, 1, 2, 4, 5, 31, 32, 33, 34, 63, 64, 65, 100, 511, 512, 513, 1023, 1024, 1025, 2049],
                                                          int(rng.integers(1, 12)))])
    sc = synth.random_scene(**kw)
    if rng.integers(0, 2):       # clustered lists make the bounds pre-test bite
        start = 0
        for cnt in sc.ml_count:
            cnt = int(cnt)
            centre = (rng.random(3) - 0.5) * 400.0
            sc.matrices[start:start + cnt, 12:15] = (centre + rng.normal(0, float(rng.choice([0.5, 6.0, 40.0])), (cnt, 3))).astype(np.float32)
            start += cnt
    ds = DeviceScene(ctx, sc)
    try:
        ds.record_drawable_processing(); ctx.sync(ds.stream)
        for frame in rng.integers(0, 360, 3):
            radius = float(rng.choice([30.0, 250.0, 600.0]))
            planes, eye = synth.orbit_camera(int(frame), radius, far=float(rng.choice([100.0, 500.0, 2000.0])))
            _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
            for bounds in (False, True):
                if bounds and not ds.bounds:
                    ds.compute_bounds(); ctx.sync(ds.stream)
                saved, ds.bounds = ds.bounds, (ds.bounds if bounds else 0)
                for fused in (False, True):
                    if fused:
                        ds.upload_drawable_list(); ds.process_and_cull(planes, eye)
                    else:
                        ds.record_drawable_processing(); ds.cull(planes, eye)
                    ctx.sync(ds.stream)
                    got = ds.read_tier_x()
                    ok = got["status"] == 0 and ref["status"] == 0 and np.array_equal(got["inst_count"], ref["inst_count"]) and got["near_band"] == ref["near_band"]
                    why = (f"status {got['status']}/{ref['status']} near {got['near_band']}/{ref['near_band']} "
                           f"inst {got['inst_count'].tolist()}/{ref['inst_count'].tolist()} queued {got['chunk_count']}")
                    if ok:
                        ok, why = canon_equal(canonicalise(got), canonicalise(ref))
                    runs += 1
                    if not ok:
                        fails += 1
                        print(f"FAIL seed {seed} frame {int(frame)} radius {radius} bounds {bounds} fused {fused}: {why}", flush=True)
                ds.bounds = saved
    finally:
        ds.close()
    if os.environ.get("FUZZ_SEEDS") and not only:
        break
    seed += 1
print(f"fuzz: {runs} frames over {seed - seed0} scenes, {fails} failures (list variant {os.environ.get('CADR_B200_CULL_VARIANT', '2')}, "
      f"staged fused pass {os.environ.get('CADR_B200_SMALL_STAGED', '0')}, library {'exp' if os.environ.get('FUZZ_EXPERIMENTS') == '1' else 'product'})")
sys.exit(1 if fails else 0)
ctx.close()
sys.exit(1 if fails else 0)
