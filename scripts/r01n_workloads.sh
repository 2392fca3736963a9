#!/bin/bash
# Final code state, one B200: randomised parity soak (all kernel variants, with and without the bounds table) and the
# bench lines of the other BASELINE configs (c1, c2, c4, c5) + the optional bounds pre-test.
tag=${1:-r01n}
mkdir -p gpurun_out
( timeout 120 python scripts/fuzz_parity.py 40 5000 ) > gpurun_out/${tag}_fuzz.log 2>&1; tail -2 gpurun_out/${tag}_fuzz.log
for w in c2 c4 c1 c5; do
  ( timeout 120 python bench.py --workload $w --no-cpu-baseline --steps 200 ) > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["kernels_ms"])
except Exception as e:
    print("$w failed", e)
PY
done
( timeout 120 python bench.py --list-bounds --no-cpu-baseline --steps 200 ) > gpurun_out/${tag}_bench_c3_list_bounds.json 2>/dev/null; tail -c 300 gpurun_out/${tag}_bench_c3_list_bounds.json
