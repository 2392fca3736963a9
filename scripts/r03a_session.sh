#!/bin/bash
# exchange tests on one GPU (bounded wait, one-launch frame close) after the timeoutMs change
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_exchange_gpu.py tests/test_capi_cpu.py -x -q ) > gpurun_out/r03a_pytest_exchange.log 2>&1; tail -5 gpurun_out/r03a_pytest_exchange.log
