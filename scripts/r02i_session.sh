#!/bin/bash
# Round 2, GPU session (one B200): CTA size of the thread-per-drawable kernel (256 = round 1, 128, 64, 32), interleaved A/B.
tag=r02i
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
( FUZZ_EXPERIMENTS=1 CADR_B200_SMALL_THREADS=64 timeout 100 python scripts/fuzz_parity.py 12 14000 ) > gpurun_out/${tag}_fuzz_t64.log 2>&1; tail -1 gpurun_out/${tag}_fuzz_t64.log
( FUZZ_EXPERIMENTS=1 CADR_B200_SMALL_THREADS=32 timeout 100 python scripts/fuzz_parity.py 12 15000 ) > gpurun_out/${tag}_fuzz_t32.log 2>&1; tail -1 gpurun_out/${tag}_fuzz_t32.log
for rep in 1 2; do
  ab c2_t256_$rep - --workload c2
  for t in 128 64 32; do ab c2_t${t}_$rep CADR_B200_SMALL_THREADS=$t --workload c2; done
done
for w in c1; do
  ab ${w}_t256 - --workload $w
  for t in 128 64 32; do ab ${w}_t${t} CADR_B200_SMALL_THREADS=$t --workload $w; done
done
ab l16_t256 - --instances 16 --drawables 2000000
for t in 128 64 32; do ab l16_t$t CADR_B200_SMALL_THREADS=$t --instances 16 --drawables 2000000; done
ab l100_t256 - --instances 100 --drawables 1000000
for t in 128 64 32; do ab l100_t$t CADR_B200_SMALL_THREADS=$t --instances 100 --drawables 1000000; done
ab c3_t256 -
for t in 128 64 32; do ab c3_t$t CADR_B200_SMALL_THREADS=$t; done
