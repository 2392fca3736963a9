#!/bin/bash
# Round 2, session w: e2e with the counters read back on their own stream (two counter blocks in turn).
tag=r02w
mkdir -p gpurun_out
for w in c3 c2 c1; do
  ( timeout 300 python bench.py --workload $w --steps 200 --warmup 5 --no-workloads --no-cpu-baseline ) > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_$w.json").read().strip().splitlines()[-1])
    print("$w value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "resident", d["e2e_resident_list"]["value"], d["e2e_resident_list"]["ms_per_step"], "facade", (d.get("e2e_facade") or {}).get("value"))
except Exception as e:
    print("$w parse failed", e)
PY
  tail -2 gpurun_out/${tag}_bench_$w.err | cut -c1-300
done
( timeout 600 python -m pytest tests/test_bench_gpu.py -x -q ) > gpurun_out/${tag}_pytest_bench.log 2>&1; tail -3 gpurun_out/${tag}_pytest_bench.log
