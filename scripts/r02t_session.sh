#!/bin/bash
# Round 2, session t: what the TMA can stream into warp-private 2-KiB stages (scripts/tma_stream.cu; one process per mode).
tag=r02t
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_stream scripts/tma_stream.cu -lcuda > gpurun_out/${tag}_tma_stream.log 2>&1
: > gpurun_out/${tag}_tma_stream.jsonl
for m in 5 0 10 1 11 2 3 4 5; do
  ( timeout 60 /tmp/tma_stream $m ) >> gpurun_out/${tag}_tma_stream.jsonl 2>> gpurun_out/${tag}_tma_stream.log
done
cat gpurun_out/${tag}_tma_stream.jsonl; tail -3 gpurun_out/${tag}_tma_stream.log
