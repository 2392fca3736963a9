#!/bin/bash
# Round 2, GPU session (one B200): A/B of two more variants of the fused first kernel against the product (records-only staging; 128-thread CTAs).
tag=r02h
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
ab() {
  local name=$1 envs=$2; shift 2
  ( if [ "$envs" != "-" ]; then export $envs; fi; timeout 200 python scripts/exp_bench.py "$@" --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  show gpurun_out/${tag}_${name}.json "$name"
}
( FUZZ_EXPERIMENTS=1 CADR_B200_SMALL_STAGED=3 timeout 100 python scripts/fuzz_parity.py 15 12000 ) > gpurun_out/${tag}_fuzz_s3.log 2>&1; tail -1 gpurun_out/${tag}_fuzz_s3.log
( FUZZ_EXPERIMENTS=1 CADR_B200_SMALL_THREADS=128 timeout 100 python scripts/fuzz_parity.py 15 13000 ) > gpurun_out/${tag}_fuzz_t128.log 2>&1; tail -1 gpurun_out/${tag}_fuzz_t128.log
for rep in 1 2; do
  ab c2_direct_$rep - --workload c2
  ab c2_recstaged_$rep CADR_B200_SMALL_STAGED=3 --workload c2
  ab c2_t128_$rep CADR_B200_SMALL_THREADS=128 --workload c2
done
ab c1_direct - --workload c1
ab c1_recstaged CADR_B200_SMALL_STAGED=3 --workload c1
ab c1_t128 CADR_B200_SMALL_THREADS=128 --workload c1
ab l16_direct - --instances 16 --drawables 2000000
ab l16_recstaged CADR_B200_SMALL_STAGED=3 --instances 16 --drawables 2000000
ab l16_t128 CADR_B200_SMALL_THREADS=128 --instances 16 --drawables 2000000
ab l100_direct - --instances 100 --drawables 1000000
ab l100_recstaged CADR_B200_SMALL_STAGED=3 --instances 100 --drawables 1000000
ab l100_t128 CADR_B200_SMALL_THREADS=128 --instances 100 --drawables 1000000
CADR_B200_SMALL_STAGED=3 timeout 200 ncu --set full --import-source on --clock-control none -k regex:cullSmallRecordsStaged -s 3 -c 1 -f -o gpurun_out/${tag}_cullSmallRecordsStagedKernel_c2 \
    python scripts/exp_bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > /dev/null 2>&1
