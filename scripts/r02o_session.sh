#!/bin/bash
# Round 2: facade c4 with the DataMemory reuse (default) and with the reference's policy; facade GPU tests.
tag=r02o
mkdir -p gpurun_out
( timeout 900 cadr_b200/host/bin/facade_bench 0 c4 40 ) > gpurun_out/${tag}_facade_c4.json 2> gpurun_out/${tag}_facade_c4.err; tail -c 700 gpurun_out/${tag}_facade_c4.json; tail -2 gpurun_out/${tag}_facade_c4.err
( timeout 600 python -m pytest tests/test_host_gpu.py tests/test_dynamic_gpu.py -x -q -m gpu ) > gpurun_out/${tag}_pytest_host.log 2>&1; tail -2 gpurun_out/${tag}_pytest_host.log
