"""GPU: the JSON contract of bench.py's product arm on a small instance of the default workload (the driver parses this
line: metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / vs_baseline / dtype /
data / config / e2e with its byte counts / gpu_launches / roofline / cpu_baseline / clocks), and its internal consistency."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_product_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "6", "--warmup", "3", "--drawables", "3000", "--cpu-sample", "200"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "culled+emitted instances/sec" and d["unit"] == "M instances/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 6 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "configs[2]" in d["config"]["workload"] and "model" not in d["config"]
    inst = d["config"]["per_gpu_instances"]
    assert inst == 3_000_000
    # value and ms_per_step describe the same measurement
    assert d["value"] > 0 and abs(d["value"] * 1e6 * d["ms_per_step"] * 1e-3 / inst - 1) < 0.02
    # end to end: host list in, counters out, every step; not a copy of the device-resident number
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 3000 * 48 + 232 and e["d2h_bytes_per_step"] == 64 + 8 * 64 and e["value"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] == 3 * d["steps"]                      # cullSmallKernel, cullListWarpKernel, cullMediumKernel per frame
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["kernel"] in ("cullListWarpKernel", "cullSmallKernel")
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3 and rf["algorithmic_bytes_per_launch"] > 64 * inst
    assert rf["launch_ms"] * d["steps"] <= d["ms_per_step"] * d["steps"] * 1.05            # the kernel fits inside the step
    assert 0 < rf["read_stream_ceiling"]["frac"] < rf["frac"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "200 of the workload's drawables" in cb["sample"]
    assert 0.05 < d["survivor_fraction"] < 0.8
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "workloads" not in d and "e2e_facade" not in d           # only the full-size default run appends those
