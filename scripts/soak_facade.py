#!/usr/bin/env python
"""Soak of the C++ facade on a GPU: facade_scene_test for many frames (dynamic scene: uploads, handle-table growth,
realloc-on-write, swap-remove, shared StateSets), every frame checked against the oracle, with and without the bounds
pre-test.  usage: scripts/soak_facade.py [frames]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import binding as ob  # noqa: E402
from facade_dump import parse  # noqa: E402
from helpers import assert_tier_x_equal  # noqa: E402
from test_host_cpu import check_frame_against_oracle  # noqa: E402

frames = sys.argv[1] if len(sys.argv) > 1 else "40"
bad = 0
for mode in ([], ["bounds"]):
    out = tempfile.mktemp(suffix=".bin")
    r = subprocess.run([os.path.join(ROOT, "cadr_b200", "host", "bin", "facade_scene_test"), "0", out, frames] + mode, capture_output=True, text=True)
    if r.returncode != 0:
        print("facade_scene_test failed:", r.stderr[-1500:]); bad += 1; continue
    fs = parse(out)
    os.remove(out)
    for i, f in enumerate(fs):
        try:
            mem, lst, ind, ptr = check_frame_against_oracle(f)
            assert np.array_equal(f["gpu_indirect"], ind) and np.array_equal(f["gpu_pointers"], ptr)
            ref = ob.cull_compact(mem, f["root"], f["level"], lst, f["n"], ind, ptr, f["cull"], f["planes"], f["eye"], f["regions"])
            assert_tier_x_equal(f["gpu_cull"], ref)
        except AssertionError as e:
            bad += 1
            print(f"FAIL mode {mode} frame {i}: {e}")
    print(f"mode {mode or ['plain']}: {len(fs)} frames, levels {sorted(set(f['level'] for f in fs))}, drawables {fs[0]['n']}..{fs[-1]['n']}")
print("soak_facade:", "ok" if not bad else f"{bad} failures")
sys.exit(1 if bad else 0)
