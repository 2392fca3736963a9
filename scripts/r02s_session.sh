#!/bin/bash
# Round 2, session s: the TMA list kernel (experiment variant 7) - tensor-map probe, parity soak, interleaved A/B against the product kernel.
tag=r02s
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/tma_probe.cu -lcuda > gpurun_out/${tag}_probe.log 2>&1
( timeout 60 /tmp/tma_probe ) >> gpurun_out/${tag}_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/${tag}_probe.log
( FUZZ_EXPERIMENTS=1 CADR_B200_CULL_VARIANT=7 timeout 200 python scripts/fuzz_parity.py 70 7000 ) > gpurun_out/${tag}_fuzz_v7.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/${tag}_fuzz_v7.log
( timeout 700 python scripts/ab_list_kernels.py --lengths 1000,65,100,200,500,5000 --variants 2,7 --steps 30 --rounds 2 ) > gpurun_out/${tag}_ab_tma.jsonl 2> gpurun_out/${tag}_ab_tma.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/${tag}_ab_tma.jsonl; tail -3 gpurun_out/${tag}_ab_tma.err
for v in 2 7; do
  ( CADR_B200_DIAG_NOEVAL=1 CADR_B200_CULL_VARIANT=$v timeout 200 python scripts/exp_bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-workloads ) > gpurun_out/${tag}_noeval_v$v.json 2> gpurun_out/${tag}_noeval_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_noeval_v$v.json").read().strip().splitlines()[-1]); print("noeval variant $v:", d["ms_per_step"], d["kernels_ms"])
except Exception as e: print("noeval $v failed", e)
PY
done
