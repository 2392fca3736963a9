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


# ---- Tier X: the one step that exists in the reference (BoundingSphere.h:70-87) ---------------------------------
def _sphere_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "bounding_sphere_ref.npz"))
    return g["matrices"], g["spheres"], g["world"]


def test_sphere_transform_matches_reference_bounding_sphere():
    """The world-space sphere of an instance as the Tier X evaluation computes it (explicit FMA chain)
    against outputs of the reference's own `operator*(const glm::mat4&, BoundingSphere)` (its vendored GLM, expression as
    written).  The two differ only in where roundings fall: tolerance 6 units of fp32 round-off on the magnitude of the
    terms (measured: <= 2.6 on the centre, <= 3.2 on the radius; ~70 % of the values are bit-identical)."""
    M, b, world = _sphere_golden()
    assert len(M) == 4096
    got = ob.transform_spheres(M, b)
    eps = 2.0 ** -24
    M64, b64 = M.astype(np.float64), b.astype(np.float64)
    for a in range(3):
        mag = np.abs(M64[:, a] * b64[:, 0]) + np.abs(M64[:, 4 + a] * b64[:, 1]) + np.abs(M64[:, 8 + a] * b64[:, 2]) + np.abs(M64[:, 12 + a])
        assert (np.abs(got[:, a].astype(np.float64) - world[:, a]) <= 6 * eps * mag).all()
    empty = ~np.isfinite(world[:, 3])
    assert empty.sum() > 10 and np.array_equal(got[empty, 3], world[empty, 3])           # BoundingSphere::empty() stays -inf
    rel = np.abs(got[~empty, 3].astype(np.float64) - world[~empty, 3]) / np.abs(world[~empty, 3].astype(np.float64))
    assert (rel <= 6 * eps).all()
    assert (got[:, :3] == world[:, :3]).mean() > 0.5


def test_plane_decisions_on_reference_spheres_differ_only_within_rounding():
    """Culling the REFERENCE's world-space spheres against a frustum gives the decisions of the oracle's evaluation
    except where the deciding plane distance is within rounding of zero - the same kind of exception north_star
    tolerates (and counts) for bounding spheres on a frustum plane."""
    M, b, world = _sphere_golden()
    got = ob.transform_spheres(M, b).astype(np.float64)
    ref = world.astype(np.float64)
    keep = np.isfinite(ref[:, 3])
    got, ref = got[keep], ref[keep]
    differing = total = 0
    for frame in range(0, 360, 24):
        planes, _ = synth.orbit_camera(frame, 1500.0, far=3000.0)
        p = planes.astype(np.float64)
        slack = lambda s: (s[:, :3] @ p[:, :3].T + p[:, 3] + s[:, 3:4]).min(axis=1)       # >= 0: visible
        sg, sr = slack(got), slack(ref)
        flip = (sg >= 0) != (sr >= 0)
        mag = np.abs(ref[:, :3]).sum(axis=1) + np.abs(p[:, 3]).max() + ref[:, 3]
        assert (np.abs(sr[flip]) <= 16 * 2.0 ** -24 * mag[flip]).all()
        differing += int(flip.sum()); total += len(flip)
        assert 0.02 < (sr >= 0).mean() < 0.98                                               # the camera culls some, keeps some
    assert differing <= total // 1000


def test_reference_struct_layouts_match_the_abi():
    """oracle/_ref/layout_check exists only if its static_asserts held when it was compiled against the reference's own
    PrimitiveSet.h / BoundingSphere.h / GLM (oracle/ref_layout_check.cpp, `make -C oracle ref`)."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "layout_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "layout_check ok" in r.stdout


def test_spirv_interpreter_against_independent_known_answers():
    """The interpreter that executes the reference's shader (oracle/spirv_run.py) is builder-written, so it gets a
    witness that is not the oracle: oracle/interp_kat.comp uses the same SPIR-V feature set as processDrawables.comp
    (buffer references, 64-bit integers, push constants, gl_WorkGroupID incl. the y component, struct / array access
    chains, a function call, shifts, masks, adds, subtractions, multiplies, pointer <-> integer conversions) to compute
    something else; its binary (compiled by the reference's vendored glslangValidator, committed as
    tests/golden/interp_kat.spv) is interpreted over 33 000 workgroups and every output word is compared with plain
    numpy arithmetic on the inputs."""
    import struct
    from oracle import spirv_run as sr
    mod = sr.load_module(os.path.join(ROOT, "tests", "golden", "interp_kat.spv"))
    rng = np.random.default_rng(5)
    n = 33_000                              # > 32768: the DispatchBase tail and gl_WorkGroupID.y
    IN, OUT, T0 = 0x7F0000100000, 0x7F0000900000, 0x7F0000F00000
    inp = np.zeros(n, dtype=[("key", "<u8"), ("a", "<u4"), ("b", "<u4"), ("link", "<u8")])
    inp["key"] = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n).astype(np.uint64)
    inp["a"] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    inp["b"] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    inp["link"] = np.uint64(IN) + rng.integers(0, n, n).astype(np.uint64) * np.uint64(24)
    tables = rng.integers(1, 1 << 62, (65, 64), dtype=np.uint64)
    tables[0] = np.uint64(T0) + np.uint64(512) * np.arange(1, 65, dtype=np.uint64)      # first level: addresses of the 64 second-level tables
    out = np.zeros(n * 40, np.uint8)
    mem = sr.Memory([(IN, inp.view(np.uint8)), (OUT, out), (T0, tables.view(np.uint8))])
    salt = 0x9E3779B97F4A7C15
    sr.dispatch(mod, mem, struct.pack("<QQQQ", T0, IN, OUT, salt), n)
    got = mem.segs[1][1].view([("mixed", "<u8"), ("lo", "<u4"), ("hi", "<u4"), ("via", "<u8"), ("linkKey", "<u8"), ("self", "<u8")])
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = inp["key"]
        expect = {
            "mixed": (k << np.uint64(3)) + inp["a"].astype(np.uint64) * inp["b"].astype(np.uint64) - np.uint64(salt) + i,
            "lo": (k & np.uint64(0x7FF)).astype(np.uint32),
            "hi": (k >> np.uint64(43)).astype(np.uint32) | (inp["a"] << np.uint32(5)),
            "via": tables[1 + ((k >> np.uint64(6)) & np.uint64(0x3F)).astype(np.int64), (k & np.uint64(0x3F)).astype(np.int64)],
            "self": np.uint64(IN) + i * np.uint64(24) + np.uint64(16),
        }
        li = ((inp["link"] - np.uint64(IN)) // np.uint64(24)).astype(np.int64)
        expect["linkKey"] = inp["key"][li] + inp["b"][li].astype(np.uint64)
    for name, e in expect.items():
        assert np.array_equal(got[name], e), name
    # an access outside device memory is reported, not emulated
    with pytest.raises(MemoryError):
        sr.dispatch(mod, sr.Memory([(IN, inp.view(np.uint8))]), struct.pack("<QQQQ", T0, IN, OUT, salt), 1)
