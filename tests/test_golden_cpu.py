"""CPU suite: pin the oracle against outputs of the REFERENCE ITSELF.

tests/golden/process_drawables_*.npz hold inputs and the outputs of the reference's own compute shader
(processDrawables.comp compiled unmodified by the reference's vendored glslangValidator and executed by
oracle/spirv_run.py; generator: oracle/make_golden.py).  The C oracle must reproduce them bit for bit.
When oracle/_ref/*.spv is present (build container), the SPIR-V is additionally executed live on fresh scenes."""
import glob
import os
import struct

import numpy as np
import pytest

from cadr_b200 import synth
from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "process_drawables_*.npz")))
LIST = 0x7F2000000000


def test_golden_files_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[len("process_drawables_"):-4])
def test_oracle_reproduces_reference_shader_output(path):
    g = np.load(path)
    base, level, root = int(g["base"]), int(g["level"]), int(g["base"]) + int(g["root_off"])
    mem = ob.Memory([(base, g["image"]), (LIST, g["drawables"])])
    n = g["drawables"].shape[0]
    ind, ptr = ob.process_drawables(mem, root, level, LIST, n, threads=2)
    assert np.array_equal(ind, g["indirect"])
    assert np.array_equal(ptr, g["pointers"])


def test_dispatch_tail_fixture_really_crosses_32768():
    g = np.load([p for p in GOLDEN if "dispatch_base_tail" in p][0])
    assert g["drawables"].shape[0] > 32768   # exercises gl_WorkGroupID.y*32768 + x (processDrawables.comp:95)


SPV = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.skipif(not os.path.exists(os.path.join(SPV, "processDrawables_L1.spv")), reason="oracle/_ref not built")
@pytest.mark.parametrize("kw", [dict(seed=201, n=80), dict(seed=202, n=80, first_handle=2040),
                                dict(seed=203, n=80, first_handle=4_194_290)], ids=["L1", "L2", "L3"])
def test_live_spirv_execution_matches_oracle(kw):
    from oracle import spirv_run as sr
    base, ind_a, ptr_a = 0x7F1200000000, 0x7F3000000000, 0x7F4000000000
    sc = synth.random_scene(**kw)
    img = sc.image(base)
    dl = np.ascontiguousarray(sc.drawables)
    ind = np.zeros((sc.n, 4), np.uint32)
    ptr = np.zeros((sc.n, 4), np.uint64)
    mod = sr.load_module(os.path.join(SPV, f"processDrawables_L{sc.handle_level}.spv"))
    sr.dispatch(mod, sr.Memory([(base, img), (LIST, dl), (ind_a, ind), (ptr_a, ptr)]),
                struct.pack("<4Q", base + sc.root_off, LIST, ind_a, ptr_a), sc.n)
    e_ind, e_ptr = ob.process_drawables(ob.Memory([(base, img), (LIST, dl)]), base + sc.root_off, sc.handle_level, LIST, sc.n)
    assert np.array_equal(ind, e_ind) and np.array_equal(ptr, e_ptr)
