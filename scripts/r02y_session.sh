#!/bin/bash
# Round 2, session y: the product list kernel at 5 / 6 CTAs per SM (variants 9 / 10: 40 / 48 warps, 48 / 40 registers, a few spills) against 4.
tag=r02y
mkdir -p gpurun_out
( timeout 700 python scripts/ab_list_kernels.py --lengths 1000,100,200,500,5000 --variants 2,9,10 --steps 30 --rounds 3 ) > gpurun_out/${tag}_ab_occupancy.jsonl 2> gpurun_out/${tag}_ab_occupancy.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/${tag}_ab_occupancy.jsonl; tail -3 gpurun_out/${tag}_ab_occupancy.err
for v in 9 10; do
( FUZZ_EXPERIMENTS=1 CADR_B200_CULL_VARIANT=$v timeout 100 python scripts/fuzz_parity.py 20 $((11000 + v)) ) > gpurun_out/${tag}_fuzz_v$v.log 2>&1; echo "fuzz v$v rc=$?"; tail -1 gpurun_out/${tag}_fuzz_v$v.log
done
