"""GPU: the CUDA kernel against the committed outputs of the reference's own shader (tests/golden/).
The fixtures were produced for a fixed arena base; handle-table entries (the only absolute addresses in the
arena) and emitted pointers are relocated to wherever cudaMalloc placed the arena."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "process_drawables_*.npz")))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[len("process_drawables_"):-4])
def test_cuda_reproduces_reference_shader_output(ctx, path):
    g = np.load(path)
    base, level, n = int(g["base"]), int(g["level"]), g["drawables"].shape[0]
    img = g["image"].copy()
    arena = ctx.arena_alloc(img.nbytes)
    dl, ind_d, ptr_d = ctx.arena_alloc(n * 48), ctx.arena_alloc(n * 16), ctx.arena_alloc(n * 32)
    try:
        delta = np.uint64(arena) - np.uint64(base) if arena >= base else None
        words = img.view(np.uint64)
        reloc = g["reloc"].astype(np.int64)
        with np.errstate(over="ignore"):
            words[reloc] = words[reloc] + np.uint64((arena - base) % (1 << 64))
        ctx.memcpy_h2d(arena, img)
        ctx.memcpy_h2d(dl, np.ascontiguousarray(g["drawables"]))
        ctx.process_drawables(arena + int(g["root_off"]), level, dl, ind_d, ptr_d, n)
        ind = np.empty((n, 4), np.uint32)
        ptr = np.empty((n, 4), np.uint64)
        ctx.memcpy_d2h(ind, ind_d)
        ctx.memcpy_d2h(ptr, ptr_d)
        ctx.sync()
        exp_ptr = g["pointers"].copy()
        with np.errstate(over="ignore"):
            exp_ptr[exp_ptr != 0] += np.uint64((arena - base) % (1 << 64))
        assert np.array_equal(ind, g["indirect"])
        assert np.array_equal(ptr, exp_ptr)
    finally:
        for a in (arena, dl, ind_d, ptr_d):
            ctx.arena_free(a)
