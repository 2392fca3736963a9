#!/bin/bash
# Round 2, session 3i: ring + pair kernel with the lane-run write-out (variant 15) against variants 6 and 2.
tag=r03i
mkdir -p gpurun_out
( FUZZ_EXPERIMENTS=1 CADR_B200_CULL_VARIANT=15 timeout 100 python scripts/fuzz_parity.py 20 15015 ) > gpurun_out/${tag}_fuzz_v15.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/${tag}_fuzz_v15.log
( timeout 700 python scripts/ab_list_kernels.py --lengths 1000,200,500,5000 --variants 2,6,15 --steps 30 --rounds 3 ) > gpurun_out/${tag}_ab_ringpair_laneruns.jsonl 2> gpurun_out/${tag}_ab_ringpair_laneruns.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/${tag}_ab_ringpair_laneruns.jsonl; tail -3 gpurun_out/${tag}_ab_ringpair_laneruns.err
