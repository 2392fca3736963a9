#!/usr/bin/env python
"""Upload path at the BASELINE configs[3] shape: 10 000 MatrixLists x 64 000 B rewritten per frame (640 MB).
Prints the device-side scatter rate (the HBM-roofline item) and the whole cadr_b200_upload call (PCIe-bound)."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cadr_b200  # noqa: E402

ctx = cadr_b200.Context(0)
lists, blk = 10_000, 64_000
total = lists * blk
arena = ctx.arena_alloc(100_000 * 64_064)
stage_dev = ctx.arena_alloc(total)
stage_host = ctx.host_alloc(total)
np.ctypeslib.as_array((ctypes.c_uint8 * total).from_address(stage_host))[:] = 7
dst = np.uint64(arena) + ((np.arange(lists, dtype=np.uint64) * np.uint64(7919)) % np.uint64(100_000)) * np.uint64(64_064) + np.uint64(64)
regions = np.stack([dst, np.arange(lists, dtype=np.uint64) * np.uint64(blk), np.full(lists, blk, np.uint64)], axis=1)
ctx.set_profiling(True)
scatter, whole = [], []
import time
for it in range(8):
    ctx.scatter_copy(regions, stage_dev)
    ctx.sync()
    scatter.append(ctx.kernel_times()[3])
for it in range(5):
    t0 = time.perf_counter()
    ctx.upload(regions, stage_host)
    ctx.sync()
    whole.append((time.perf_counter() - t0) * 1e3)
peak = 6541.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ms = float(np.median(scatter[2:]))
print(json.dumps({"workload": "configs[3] upload: 10000 regions x 64000 B", "scatter_kernel_ms": round(ms, 4),
                  "scatter_GBps_algorithmic(2x bytes)": round(2 * total / ms / 1e6, 1), "frac_of_measured_hbm": round(2 * total / ms / 1e6 / peak, 3),
                  "upload_call_ms(H2D 640 MB + scatter)": round(float(np.median(whole[1:])), 3),
                  "upload_GBps": round(total / float(np.median(whole[1:])) / 1e6, 1)}))
