"""GPU: export of device buffers as POSIX file descriptors (north_star's optional cudaExternalMemory /
Vulkan hand-over).  No Vulkan ICD exists on the test machines, so the importing side is CUDA again: a second mapping in
the same process and a separate consumer process that inherits the descriptor."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import cadr_b200
from cadr_b200 import synth
from cadr_b200.frame import DeviceScene
from helpers import assert_tier_x_equal, oracle_tier_x

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_export_import_round_trip_same_process(ctx):
    addr, size = ctx.external_alloc(1_000_000)
    assert size >= 1_000_000 and size % 65536 == 0
    fd = ctx.external_export_fd(addr)
    try:
        assert fd >= 0
        view = ctx.external_import_fd(fd, size)
        assert view != addr
        pattern = np.random.default_rng(5).integers(0, 256, 1_000_000, dtype=np.uint8)
        ctx.memcpy_h2d(addr, pattern); ctx.sync()
        back = np.empty_like(pattern)
        ctx.memcpy_d2h(back, view); ctx.sync()
        assert np.array_equal(back, pattern)               # the second mapping shows the same memory
        ctx.memcpy_h2d(view + 4096, pattern[:4096][::-1].copy()); ctx.sync()
        ctx.memcpy_d2h(back, addr); ctx.sync()
        assert np.array_equal(back[4096:8192], pattern[:4096][::-1])
        ctx.external_free(view)
    finally:
        os.close(fd)
        ctx.external_free(addr)


def test_misuse_is_a_logic_error(ctx):
    a = ctx.arena_alloc(4096)
    try:
        with pytest.raises(cadr_b200.LogicError):
            ctx.external_export_fd(a)                      # only external_alloc buffers can be exported
        with pytest.raises(cadr_b200.LogicError):
            ctx.external_free(a)
        with pytest.raises(cadr_b200.LogicError):
            ctx.external_import_fd(-1, 1 << 21)
    finally:
        ctx.arena_free(a)


def test_culled_frame_in_exported_buffers_is_read_by_another_process(ctx):
    """The frame's outputs live in exportable buffers; a consumer process maps the instance-index buffer and the
    counters from inherited descriptors and sees exactly what the producer sees; the result still equals the oracle."""
    sizes = {}

    def alloc(nbytes):
        addr, size = ctx.external_alloc(nbytes)
        sizes[addr] = size
        return addr

    sc = synth.random_scene(61, n=600, num_lists=70, max_count=300, state_sets=4, big_lists=2)
    ds = DeviceScene(ctx, sc, alloc=alloc, free=ctx.external_free)
    try:
        planes, eye = synth.orbit_camera(33, 250.0, far=500.0)
        ds.record_drawable_processing()
        ds.cull(planes, eye)
        ctx.sync(ds.stream)
        got = ds.read_tier_x()
        _, _, ref = oracle_tier_x(sc, planes, eye, arena_base=ds.arena, list_base=ds.drawable_list)
        assert_tier_x_equal(got, ref)
        for addr, nbytes, mine in ((ds.inst_out, ds.inst_cap * 4, got["inst"].view(np.uint8)),
                                   (ds.counters, ds.counters_bytes, None)):
            if mine is None:
                mine = np.empty(nbytes, np.uint8); ctx.memcpy_d2h(mine, addr); ctx.sync()
            fd = ctx.external_export_fd(addr)
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "external_consumer.py"), str(fd), str(sizes[addr]), str(nbytes)],
                                   pass_fds=[fd], capture_output=True, text=True, timeout=300)
            finally:
                os.close(fd)
            assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
            assert f"digest {hashlib.sha256(mine[:nbytes].tobytes()).hexdigest()}" in r.stdout
    finally:
        ds.close()
