#!/bin/bash
tag=${1:-r01k}
mkdir -p gpurun_out
( time timeout 500 python scripts/ab_list_kernels.py --variants 2,4,5,6 --lengths 33,64,65,100,1000 --rounds 3 ) > gpurun_out/${tag}_ab.jsonl 2> gpurun_out/${tag}_ab.err; cut -c1-330 gpurun_out/${tag}_ab.jsonl; tail -5 gpurun_out/${tag}_ab.err
