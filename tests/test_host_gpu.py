"""GPU: the C++ facade end to end — CadR::Renderer frames on a real B200 against the oracle.

facade_scene_test drives a dynamic scene (uploads through DataStorage/StagingManager, handle-table growth over
both level transitions, realloc-on-write, swap-remove, a StateSet with two parents) through
beginFrame .. executeCopyOperations .. recordDrawableProcessing .. recordDrawableCulling .. submit, and dumps
what the GPU produced.  The oracle runs on the device image reconstructed from the copy regions the facade issued."""
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from facade_dump import parse
from helpers import assert_tier_x_equal
from test_host_cpu import BOX_SCENES, boxes_expectation, check_frame_against_oracle, run_boxes_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "cadr_b200", "host", "bin")
pytestmark = pytest.mark.gpu


def test_reference_data_allocation_scenarios_on_device():
    r = subprocess.run([os.path.join(BIN, "data_allocation_test"), "0", "420"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.fixture(scope="module", params=["plain", "bounds"])
def gpu_frames(tmp_path_factory, request):
    """`bounds`: the same scene with Renderer::setDrawableBounds(true) — the pre-test must not change any result."""
    out = str(tmp_path_factory.mktemp("facade") / "scene.bin")
    r = subprocess.run([os.path.join(BIN, "facade_scene_test"), "0", out, "5"] + (["bounds"] if request.param == "bounds" else []),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    frames = parse(out)
    os.remove(out)
    for f, line in zip(frames, [l for l in r.stderr.splitlines() if l.startswith("frame ")]):
        f["list_upload_bytes"] = int(line.split("list upload ")[1].split()[0])
    return frames


def test_facade_uploads_only_changed_drawable_ranges(gpu_frames):
    """Device-resident drawable list: frame 0 copies everything, a frame that only rewrites transforms
    copies nothing — and the results above are identical either way."""
    assert gpu_frames[0]["list_upload_bytes"] == gpu_frames[0]["n"] * 96     # 48 B record + 48 B culling record
    assert gpu_frames[4]["list_upload_bytes"] == 0                           # frame 4 only rewrites a MatrixList
    assert 0 < gpu_frames[2]["list_upload_bytes"] <= gpu_frames[2]["n"] * 96


def test_facade_frames_on_gpu_match_oracle(gpu_frames):
    assert [f["level"] for f in gpu_frames] == [1, 2, 2, 3, 3]
    for f in gpu_frames:
        assert f["has_device"]
        mem, lst, ind, ptr = check_frame_against_oracle(f)
        # Tier R: what the GPU wrote into the indirect / pointers buffers
        assert np.array_equal(f["gpu_indirect"], ind)
        assert np.array_equal(f["gpu_pointers"], ptr)
        # Tier X: compacted command lists of the same frame
        ref = ob.cull_compact(mem, f["root"], f["level"], lst, f["n"], ind, ptr, f["cull"], f["planes"], f["eye"], f["regions"])
        assert_tier_x_equal(f["gpu_cull"], ref)
        assert ref["num_instances"] > 0


def test_facade_long_run_with_several_staging_blocks():
    """40 frames of the dynamic facade scene, with and without the bounds pre-test, every frame against the oracle
    (scripts/soak_facade.py).  Regression: from frame 5 on the uploads of one executeCopyOperations come from more than one
    staging block."""
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "soak_facade.py"), "40"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "soak_facade: ok" in r.stdout


@pytest.mark.parametrize("scene", BOX_SCENES)
def test_reference_example_scenes_on_gpu_match_oracle(scene, tmp_path):
    """The scenes of the reference's examples/RenderingPerformance (Tests.cpp) driven through CadR::Renderer on the
    device, the show/hide scenes destroying and re-creating their drawables every frame: Tier R buffers and the
    compacted Tier X buffers of every frame against the oracle."""
    frames = run_boxes_scene(scene, 0, tmp_path)
    for k, f in enumerate(frames):
        assert f["has_device"]
        mem, lst, ind, ptr = check_frame_against_oracle(f)
        assert f["n"] == boxes_expectation(scene, k)[0]
        assert np.array_equal(f["gpu_indirect"], ind)
        assert np.array_equal(f["gpu_pointers"], ptr)
        ref = ob.cull_compact(mem, f["root"], f["level"], lst, f["n"], ind, ptr, f["cull"], f["planes"], f["eye"], f["regions"])
        assert_tier_x_equal(f["gpu_cull"], ref)
        assert ref["num_instances"] > 0
