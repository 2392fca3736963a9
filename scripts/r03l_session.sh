#!/bin/bash
# Round 2, session 3l: the fused first kernel (64-thread CTAs) with a register cap: 18 / 20 CTAs per SM (56 / 48 registers, 36 / 40 warps) against the product (64 registers, 32 warps).
tag=r03l
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(json.dumps({"run": sys.argv[2], "value": d["value"], "ms_per_step": d["ms_per_step"], "kernel": d["roofline"]["kernel"], "frac": d["roofline"]["frac"], "frac_of_line_granular_floor": d["roofline"].get("frac_of_line_granular_floor"), "kernels_ms": d["kernels_ms"]}))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
ab() {
  local name=$1 envs=$2; shift 2
  ( if [ "$envs" != "-" ]; then export $envs; fi; timeout 200 python scripts/exp_bench.py "$@" --no-cpu-baseline --no-workloads --steps 200 ) > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  show gpurun_out/${tag}_${name}.json "$name" | tee -a gpurun_out/${tag}_ab_small_minctas.jsonl
}
: > gpurun_out/${tag}_ab_small_minctas.jsonl
( FUZZ_EXPERIMENTS=1 CADR_B200_SMALL_MINCTAS=18 timeout 100 python scripts/fuzz_parity.py 12 18000 ) > gpurun_out/${tag}_fuzz_m18.log 2>&1; tail -1 gpurun_out/${tag}_fuzz_m18.log
for rep in 1 2; do
  ab c2_product_$rep - --workload c2
  for m in 18 20; do ab c2_minctas${m}_$rep CADR_B200_SMALL_MINCTAS=$m --workload c2; done
done
ab c1_product - --workload c1
for m in 18 20; do ab c1_minctas$m CADR_B200_SMALL_MINCTAS=$m --workload c1; done
ab l16_product - --instances 16 --drawables 2000000
ab l16_minctas18 CADR_B200_SMALL_MINCTAS=18 --instances 16 --drawables 2000000
