#!/bin/bash
# Round 2, GPU session (one B200): GPU test suite incl. the one-device exchange test, smoke, and A/B of the look-ahead L2
# prefetch experiment in the fused first kernel; c3 with ONE StateSet (drawable order == list order: DRAM locality check).
tag=r02g
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/${tag}_pytest.log 2>&1; tail -10 gpurun_out/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("frac_of_line_granular_floor"), d["kernels_ms"])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
ab() {
  local name=$1 envs=$2; shift 2
  ( if [ "$envs" != "-" ]; then export $envs; fi; timeout 200 python scripts/exp_bench.py "$@" --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  show gpurun_out/${tag}_${name}.json "$name"
}
( FUZZ_EXPERIMENTS=1 CADR_B200_SMALL_PREFETCH=1 timeout 100 python scripts/fuzz_parity.py 15 11000 ) > gpurun_out/${tag}_fuzz_pf.log 2>&1; tail -1 gpurun_out/${tag}_fuzz_pf.log
for rep in 1 2; do
  ab c2_direct_$rep - --workload c2
  for pf in 640 1024 2048 4096; do ab c2_pf${pf}_$rep CADR_B200_SMALL_PREFETCH=$pf --workload c2; done
done
ab c1_direct - --workload c1
for pf in 640 1024; do ab c1_pf${pf} CADR_B200_SMALL_PREFETCH=$pf --workload c1; done
ab l16_direct - --instances 16 --drawables 2000000
ab l16_pf1024 CADR_B200_SMALL_PREFETCH=1024 --instances 16 --drawables 2000000
ab c3_64sets -
ab c3_1set - --state-sets 1
