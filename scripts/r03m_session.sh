#!/bin/bash
# last code state: GPU tests + smoke + the driver-form bench once more (the product library was rebuilt after r02z: emitItem became a template)
tag=r03m
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
( timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_driver.json 2> gpurun_out/${tag}_bench_c3_driver.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r03m_bench_c3_driver.json").read().strip().splitlines()[-1])
    print("c3", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], d["kernels_ms"])
    for w, c in d.get("workloads", {}).items():
        print(w, {k: c.get(k) for k in ("value", "ms_per_step", "error")}, "e2e", c.get("e2e", {}).get("value"))
    print("facade", {k: d["e2e_facade"].get(k) for k in ("value", "ms_per_step", "error")})
except Exception as e:
    print("parse failed", e)
PY
