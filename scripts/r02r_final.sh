#!/bin/bash
# Round 2, last code state, one GPU: what the driver runs at round end (GPU tests, smoke, both bench arms), kept as evidence.
tag=r02r
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
( timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_reference.json 2>/dev/null; tail -c 250 gpurun_out/${tag}_bench_reference.json; echo
t0=$(date +%s)
( timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_driver.json 2> gpurun_out/${tag}_bench_c3_driver.err
echo "default bench.py took $(( $(date +%s) - t0 )) s"; tail -c 300 gpurun_out/${tag}_bench_c3_driver.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02r_bench_c3_driver.json").read().strip().splitlines()[-1])
    print("c3", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["read_stream_ceiling"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"), "launches", d["gpu_launches"], d["clocks"])
    for w, c in d.get("workloads", {}).items():
        print(w, {k: c.get(k) for k in ("value", "ms_per_step", "error")}, "e2e", c.get("e2e", {}).get("value"), "frac", c.get("roofline", {}).get("frac"), c.get("roofline", {}).get("frac_of_line_granular_floor"))
    print("facade", {k: d["e2e_facade"].get(k) for k in ("value", "ms_per_step", "gpuDrawableProcessing_ms", "error")})
except Exception as e:
    print("parse failed", e)
PY
( timeout 900 cadr_b200/host/bin/facade_bench 0 c4 30 ) > gpurun_out/${tag}_facade_c4.json 2> gpurun_out/${tag}_facade_c4.err; tail -c 600 gpurun_out/${tag}_facade_c4.json
