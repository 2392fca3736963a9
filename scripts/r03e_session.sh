#!/bin/bash
# Round 2, session 3e: the product list kernel with the lane-run index write-out (variant 12) against the product (variant 2).
tag=r03e
mkdir -p gpurun_out
( FUZZ_EXPERIMENTS=1 CADR_B200_CULL_VARIANT=12 timeout 100 python scripts/fuzz_parity.py 25 13012 ) > gpurun_out/${tag}_fuzz_v12.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/${tag}_fuzz_v12.log
( timeout 700 python scripts/ab_list_kernels.py --lengths 1000,65,100,200,500,5000 --variants 2,12 --steps 30 --rounds 3 ) > gpurun_out/${tag}_ab_laneruns.jsonl 2> gpurun_out/${tag}_ab_laneruns.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/${tag}_ab_laneruns.jsonl; tail -3 gpurun_out/${tag}_ab_laneruns.err
