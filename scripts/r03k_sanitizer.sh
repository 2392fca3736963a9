#!/bin/bash
# compute-sanitizer initcheck (reads of uninitialised device memory) and synccheck over the cull / exchange / upload tests at small sizes
tag=r03k
mkdir -p gpurun_out
T="tests/test_exchange_gpu.py tests/test_parity_gpu.py::test_tier_x_list_length_boundaries tests/test_parity_gpu.py::test_tier_x_medium_lists_in_flat_batches tests/test_parity_gpu.py::test_upload_scatter_and_patch tests/test_parity_gpu.py::test_two_phase_upload_stages_on_a_copy_stream_and_commits_in_order tests/test_golden_gpu.py"
( timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 20 python -m pytest $T -x -q -m gpu ) > gpurun_out/${tag}_initcheck.log 2>&1; echo "initcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Uninitialized|error" gpurun_out/${tag}_initcheck.log | tail -8
( timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_exchange_gpu.py tests/test_parity_gpu.py::test_tier_x_list_length_boundaries tests/test_parity_gpu.py::test_tier_x_medium_lists_in_flat_batches -x -q -m gpu ) > gpurun_out/${tag}_synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|error" gpurun_out/${tag}_synccheck.log | tail -5
