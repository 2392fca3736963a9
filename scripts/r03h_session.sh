#!/bin/bash
# Round 2, session 3h: does the access pattern of the two LDG.256 per lane (each instruction half of 16 lines) cost bandwidth?  L2 prefetch-size hints, contiguous instructions.
tag=r03h
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_stream scripts/tma_stream.cu -lcuda > gpurun_out/${tag}_tma_stream.log 2>&1
: > gpurun_out/${tag}_ldg_patterns.jsonl
for m in 5 6 7 8 5 6 7 8; do
  ( timeout 60 /tmp/tma_stream $m ) >> gpurun_out/${tag}_ldg_patterns.jsonl 2>> gpurun_out/${tag}_tma_stream.log
done
cat gpurun_out/${tag}_ldg_patterns.jsonl; tail -3 gpurun_out/${tag}_tma_stream.log
