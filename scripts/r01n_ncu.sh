#!/bin/bash
# ncu --set full of the dominant kernels with the final code state (c3: cullListWarpKernel, c2: fused cullSmallKernel).
tag=${1:-r01n}
mkdir -p gpurun_out
timeout 110 ncu --set full --import-source on --clock-control none -k regex:cullListWarp -c 1 -f -o gpurun_out/${tag}_cullListWarpKernel_c3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 110 ncu --set full --import-source on --clock-control none -k regex:cullSmall -s 3 -c 1 -f -o gpurun_out/${tag}_cullSmallKernel_fused_c2 \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
