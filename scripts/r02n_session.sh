#!/bin/bash
# Round 2: the dynamic scene (configs[3]) through the C++ facade on the device, small then full size; final driver-form bench.
tag=r02n
mkdir -p gpurun_out
( timeout 300 cadr_b200/host/bin/facade_bench 0 c4 10 3000 1000 ) > gpurun_out/${tag}_facade_c4_small.json 2> gpurun_out/${tag}_facade_c4_small.err; tail -c 900 gpurun_out/${tag}_facade_c4_small.json; tail -2 gpurun_out/${tag}_facade_c4_small.err
( timeout 900 cadr_b200/host/bin/facade_bench 0 c4 30 ) > gpurun_out/${tag}_facade_c4.json 2> gpurun_out/${tag}_facade_c4.err; tail -c 1200 gpurun_out/${tag}_facade_c4.json; tail -2 gpurun_out/${tag}_facade_c4.err
t0=$(date +%s)
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_driver.json 2> gpurun_out/${tag}_bench_c3_driver.err
echo "default bench.py took $(( $(date +%s) - t0 )) s"; tail -c 300 gpurun_out/${tag}_bench_c3_driver.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02n_bench_c3_driver.json").read().strip().splitlines()[-1])
    print("c3", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["read_stream_ceiling"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"))
    for w, c in d.get("workloads", {}).items():
        print(w, {k: c.get(k) for k in ("value", "ms_per_step", "error")}, "e2e", c.get("e2e", {}).get("value"), "frac", c.get("roofline", {}).get("frac"), c.get("roofline", {}).get("frac_of_line_granular_floor"))
    print("facade", {k: d["e2e_facade"].get(k) for k in ("value", "ms_per_step", "gpuDrawableProcessing_ms", "error")})
except Exception as e:
    print("parse failed", e)
PY
