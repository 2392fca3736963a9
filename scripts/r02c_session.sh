#!/bin/bash
# Round 2, GPU session (one B200): parity of the staged fused kernel (cullSmallStagedKernel) and its A/B against the
# direct-load version on the same box (c2, c1, a 16-matrix shape), plus c3.
tag=r02c
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -8 gpurun_out/${tag}_pytest.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("frac_of_line_granular_floor"), d["kernels_ms"])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for rep in 1 2; do
for w in c2 c1; do
  ( timeout 200 python scripts/exp_bench.py --workload $w --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_ab_${w}_staged_$rep.json 2> gpurun_out/${tag}_ab_${w}_staged_$rep.err; show gpurun_out/${tag}_ab_${w}_staged_$rep.json "$w staged"
  ( CADR_B200_SMALL_DIRECT=1 timeout 200 python scripts/exp_bench.py --workload $w --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_ab_${w}_direct_$rep.json 2> gpurun_out/${tag}_ab_${w}_direct_$rep.err; show gpurun_out/${tag}_ab_${w}_direct_$rep.json "$w direct"
done
done
( timeout 200 python scripts/exp_bench.py --instances 16 --drawables 2000000 --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_ab_16_staged.json 2>/dev/null; show gpurun_out/${tag}_ab_16_staged.json "16x2M staged"
( CADR_B200_SMALL_DIRECT=1 timeout 200 python scripts/exp_bench.py --instances 16 --drawables 2000000 --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_ab_16_direct.json 2>/dev/null; show gpurun_out/${tag}_ab_16_direct.json "16x2M direct"
( timeout 200 python bench.py --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; show gpurun_out/${tag}_bench_c3.json "c3"
tail -3 gpurun_out/${tag}_ab_c2_staged_1.err
