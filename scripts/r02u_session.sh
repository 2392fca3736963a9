#!/bin/bash
# Round 2, session u: cullListTmaKernel with a step cut into 4 x 512 B (variant 7) / 8 x 256 B (variant 8) bulk copies: parity soak + A/B.
tag=r02u
mkdir -p gpurun_out
for v in 7 8; do
  ( FUZZ_EXPERIMENTS=1 CADR_B200_CULL_VARIANT=$v timeout 120 python scripts/fuzz_parity.py 30 $((9000 + v)) ) > gpurun_out/${tag}_fuzz_v$v.log 2>&1; echo "fuzz v$v rc=$?"; tail -2 gpurun_out/${tag}_fuzz_v$v.log
done
( timeout 700 python scripts/ab_list_kernels.py --lengths 1000,100,200,5000 --variants 2,7,8 --steps 30 --rounds 2 ) > gpurun_out/${tag}_ab_tma.jsonl 2> gpurun_out/${tag}_ab_tma.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/${tag}_ab_tma.jsonl; tail -3 gpurun_out/${tag}_ab_tma.err
for v in 2 7 8; do
  ( CADR_B200_DIAG_NOEVAL=1 CADR_B200_CULL_VARIANT=$v timeout 200 python scripts/exp_bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads ) > gpurun_out/${tag}_noeval_v$v.json 2> gpurun_out/${tag}_noeval_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_noeval_v$v.json").read().strip().splitlines()[-1]); print("noeval variant $v:", d["ms_per_step"], d["kernels_ms"])
except Exception as e: print("noeval $v failed", e)
PY
done
