#!/bin/bash
# Round 2, session 3j: list kernel with contiguous load instructions + pair exchange (variant 16) against the product (variant 2).
tag=r03j
mkdir -p gpurun_out
( FUZZ_EXPERIMENTS=1 CADR_B200_CULL_VARIANT=16 timeout 100 python scripts/fuzz_parity.py 25 16016 ) > gpurun_out/${tag}_fuzz_v16.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/${tag}_fuzz_v16.log
( timeout 700 python scripts/ab_list_kernels.py --lengths 1000,100,200,500,5000 --variants 2,16 --steps 30 --rounds 3 ) > gpurun_out/${tag}_ab_contig.jsonl 2> gpurun_out/${tag}_ab_contig.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/${tag}_ab_contig.jsonl; tail -3 gpurun_out/${tag}_ab_contig.err
for v in 2 16; do
  ( CADR_B200_DIAG_NOEVAL=1 CADR_B200_CULL_VARIANT=$v timeout 200 python scripts/exp_bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads ) > gpurun_out/${tag}_noeval_v$v.json 2> gpurun_out/${tag}_noeval_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_noeval_v$v.json").read().strip().splitlines()[-1]); print("noeval variant $v:", d["ms_per_step"], d["kernels_ms"])
except Exception as e: print("noeval $v failed", e)
PY
done
