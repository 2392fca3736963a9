#!/bin/bash
tag=${1:-r01l}
mkdir -p gpurun_out
: > gpurun_out/${tag}_pad.jsonl
for pad in 0 4096 8192 12288 20480 28672 32768; do
  echo "pad $pad" >> gpurun_out/${tag}_pad.jsonl
  CADR_B200_DIAG_SMEM_PAD=$pad timeout 200 python scripts/ab_list_kernels.py --variants 4,6 --lengths 1000 --rounds 3 --steps 60 >> gpurun_out/${tag}_pad.jsonl 2>> gpurun_out/${tag}_pad.err
done
cut -c1-330 gpurun_out/${tag}_pad.jsonl | grep -v '"variant": "2"'
