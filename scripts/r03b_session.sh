#!/bin/bash
# Round 2, session 3b: ring + packed-pair list kernel with two stages per warp and three CTAs per SM (variant 11) against variants 6 and 2.
tag=r03b
mkdir -p gpurun_out
( timeout 700 python scripts/ab_list_kernels.py --lengths 1000,100,200,500,5000 --variants 2,6,11 --steps 30 --rounds 3 ) > gpurun_out/${tag}_ab_ringpair.jsonl 2> gpurun_out/${tag}_ab_ringpair.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/${tag}_ab_ringpair.jsonl; tail -3 gpurun_out/${tag}_ab_ringpair.err
( FUZZ_EXPERIMENTS=1 CADR_B200_CULL_VARIANT=11 timeout 100 python scripts/fuzz_parity.py 20 12011 ) > gpurun_out/${tag}_fuzz_v11.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/${tag}_fuzz_v11.log
for v in 2 11; do
  ( CADR_B200_DIAG_NOEVAL=1 CADR_B200_CULL_VARIANT=$v timeout 200 python scripts/exp_bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads ) > gpurun_out/${tag}_noeval_v$v.json 2> gpurun_out/${tag}_noeval_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_noeval_v$v.json").read().strip().splitlines()[-1]); print("noeval variant $v:", d["ms_per_step"], d["kernels_ms"])
except Exception as e: print("noeval $v failed", e)
PY
done
