#!/usr/bin/env python
"""Randomised soak of the upload path on a GPU: random non-overlapping copy regions (any alignment, 0 .. 3 MiB, dense and
sparse sources) through cadr_b200_upload (host staging) and cadr_b200_scatter_copy (device staging) against the oracle's
byte-exact restatement.  usage: scripts/fuzz_upload.py [seconds] [first_seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cadr_b200  # noqa: E402
from oracle import binding as ob  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = cadr_b200.Context(0)
size = 24 << 20
arena = ctx.arena_alloc(size)
stage_dev = ctx.arena_alloc(size)
host = np.zeros(size, np.uint8)
got = np.empty(size, np.uint8)
ctx.memset(arena, 0, size); ctx.sync()
t_end, runs, fails = time.time() + budget, 0, 0
while time.time() < t_end:
    rng = np.random.default_rng(seed)
    staging = rng.integers(0, 256, size, dtype=np.uint8)
    n = int(rng.integers(1, 400))
    # destination: sorted cut points -> disjoint slots; each region uses a random part of its slot
    cuts = np.sort(rng.integers(0, size, 2 * n))
    regs = []
    for k in range(n):
        lo, hi = int(cuts[2 * k]), int(cuts[2 * k + 1])
        if rng.integers(0, 3) == 0:
            lo = (lo + 15) & ~15                              # the allocator's guarantee, most of the time
        b = min(hi - lo, int(rng.choice([0, 1, 7, 16, 100, 4096, 70001, 1 << 20, 3 << 20])))
        if b < 0:
            continue
        src = int(rng.integers(0, size - b + 1))
        if rng.integers(0, 2):
            src &= ~15
        regs.append((arena + lo, src, b))
    regions = np.array(regs, np.uint64).reshape(-1, 3)
    device = bool(rng.integers(0, 2))
    if device:
        ctx.memcpy_h2d(stage_dev, staging); ctx.scatter_copy(regions, stage_dev)
    else:
        ctx.upload(regions, staging)
    ctx.memcpy_d2h(got, arena); ctx.sync()
    ob.upload(ob.Memory([(arena, host)]), regions, staging)
    runs += 1
    if not np.array_equal(got, host):
        fails += 1
        bad = np.flatnonzero(got != host)
        print(f"FAIL seed {seed} ({'scatter_copy' if device else 'upload'}, {len(regs)} regions): {bad.size} bytes differ, first at {bad[0]}", flush=True)
        host[:] = got
    seed += 1
print(f"fuzz_upload: {runs} calls, {fails} failures")
ctx.arena_free(stage_dev); ctx.arena_free(arena); ctx.close()
sys.exit(1 if fails else 0)
