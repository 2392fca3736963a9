#!/bin/bash
# Round 2, GPU session (one B200): A/B of the experiment kernels against the product path on the same box, facade_bench.
tag=r02d
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("frac_of_line_granular_floor"), d["kernels_ms"])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
ab() {  # ab <name> <env assignment or -> <bench args...>
  local name=$1 envs=$2; shift 2
  ( if [ "$envs" != "-" ]; then export $envs; fi; timeout 200 python scripts/exp_bench.py "$@" --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  show gpurun_out/${tag}_${name}.json "$name"
}
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -6 gpurun_out/${tag}_pytest.log
# parity of the experiment kernels first (fuzz through the exp library)
( FUZZ_EXPERIMENTS=1 CADR_B200_CULL_VARIANT=6 CADR_B200_SMALL_STAGED=2 timeout 150 python scripts/fuzz_parity.py 25 7000 ) > gpurun_out/${tag}_fuzz_exp.log 2>&1; tail -3 gpurun_out/${tag}_fuzz_exp.log
( FUZZ_EXPERIMENTS=1 CADR_B200_SMALL_STAGED=1 timeout 100 python scripts/fuzz_parity.py 15 8000 ) > gpurun_out/${tag}_fuzz_exp2.log 2>&1; tail -2 gpurun_out/${tag}_fuzz_exp2.log
( timeout 100 python scripts/fuzz_parity.py 15 9000 ) > gpurun_out/${tag}_fuzz.log 2>&1; tail -2 gpurun_out/${tag}_fuzz.log
for rep in 1 2; do
  ab c2_direct_$rep - --workload c2
  ab c2_staged256_$rep CADR_B200_SMALL_STAGED=1 --workload c2
  ab c2_staged128_$rep CADR_B200_SMALL_STAGED=2 --workload c2
  ab c3_warp_$rep - 
  ab c3_ringpair_$rep CADR_B200_CULL_VARIANT=6
done
ab c1_direct - --workload c1
ab c1_staged256 CADR_B200_SMALL_STAGED=1 --workload c1
ab c1_staged128 CADR_B200_SMALL_STAGED=2 --workload c1
for m in 100 200 500; do
  ab l${m}_warp - --instances $m --drawables $((100000000 / m))
  ab l${m}_ringpair CADR_B200_CULL_VARIANT=6 --instances $m --drawables $((100000000 / m))
done
ab c3_ringpair_noeval "CADR_B200_CULL_VARIANT=6 CADR_B200_DIAG_NOEVAL=1"
ab c3_warp_noeval "CADR_B200_DIAG_NOEVAL=1"
( timeout 300 python bench.py --workload c4 --no-cpu-baseline --no-workloads --steps 100 ) > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02d_bench_c4.json").read().strip().splitlines()[-1])
    print("c4", d["value"], d["ms_per_step"], "e2e", d["e2e"], "upload", d["upload"])
except Exception as e:
    print("c4 failed", e)
PY
tail -3 gpurun_out/${tag}_bench_c4.err
# the facade's frame loop on the device
for sc in c1 c2 c3; do
  ( timeout 300 cadr_b200/host/bin/facade_bench 0 $sc 200 ) > gpurun_out/${tag}_facade_$sc.json 2> gpurun_out/${tag}_facade_$sc.err; tail -c 900 gpurun_out/${tag}_facade_$sc.json; tail -2 gpurun_out/${tag}_facade_$sc.err
done
