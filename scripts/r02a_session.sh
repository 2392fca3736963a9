#!/bin/bash
# Round 2, first GPU session (one B200): GPU tests after the PDL / hygiene / allocator changes incl. the whole-scene
# parity tests, and bench lines of c3 / c2 / c1 as the starting point for the kernel work of this round.
tag=r02a
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/${tag}_pytest.log 2>&1; tail -15 gpurun_out/${tag}_pytest.log
for w in c3 c2 c1; do
  ( timeout 200 python bench.py --workload $w --no-cpu-baseline --steps 200 ) > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["kernels_ms"])
except Exception as e:
    print("$w failed", e)
PY
done
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
