#!/bin/bash
# Round 2, session 3g: the product list kernel with 4 / 3 / 2 CTAs per SM launched (32 / 24 / 16 warps = 64 / 48 / 32 KiB of matrix loads in flight per SM).
tag=r03g
mkdir -p gpurun_out
( timeout 700 python scripts/ab_list_kernels.py --lengths 1000 --variants 2,13,14 --steps 30 --rounds 2 ) > gpurun_out/${tag}_ab_warps_in_flight.jsonl 2> gpurun_out/${tag}_ab_warps_in_flight.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/${tag}_ab_warps_in_flight.jsonl; tail -3 gpurun_out/${tag}_ab_warps_in_flight.err
for v in 2 13 14; do
  ( CADR_B200_DIAG_NOEVAL=1 CADR_B200_CULL_VARIANT=$v timeout 200 python scripts/exp_bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads ) > gpurun_out/${tag}_noeval_v$v.json 2> gpurun_out/${tag}_noeval_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_noeval_v$v.json").read().strip().splitlines()[-1]); print("noeval variant $v:", d["ms_per_step"], d["kernels_ms"])
except Exception as e: print("noeval $v failed", e)
PY
done
