#!/bin/bash
# compute-sanitizer on the medium-list path (cullMediumKernel, medium queue in cullSmallKernel) with the final code.
tag=${1:-r01n}
mkdir -p gpurun_out
SEL='medium_lists and medium-kernel'
if [ "$2" != wide ]; then
( timeout 70 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL" ) > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck.log | tail -3
( timeout 60 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL and mixed and fused" ) > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_racecheck.log | tail -3
( timeout 40 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL and mixed and fused" ) > gpurun_out/${tag}_synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_synccheck.log | tail -3
fi
if [ "$2" = wide ]; then
  SEL2='(boundaries and medium-kernel) or random_scenes or fused_process or work_item_queue'
  ( timeout 75 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL2" ) > gpurun_out/${tag}_memcheck_wide.log 2>&1; echo "memcheck wide rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck_wide.log | tail -3
  ( timeout 45 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "boundaries and medium-kernel" ) > gpurun_out/${tag}_racecheck_wide.log 2>&1; echo "racecheck wide rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_racecheck_wide.log | tail -3
fi
