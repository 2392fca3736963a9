#!/bin/bash
# compute-sanitizer over the kernels added or changed in round 2 (memcheck + racecheck on a subset of the GPU tests that
# exercises them at small sizes), and a longer randomised parity soak of the product library.
tag=r03d
mkdir -p gpurun_out
T="tests/test_exchange_gpu.py tests/test_parity_gpu.py::test_two_phase_upload_stages_on_a_copy_stream_and_commits_in_order tests/test_parity_gpu.py::test_tier_x_list_length_boundaries tests/test_parity_gpu.py::test_tier_x_medium_lists_in_flat_batches tests/test_parity_gpu.py::test_primitive_set_offsets_need_only_four_byte_alignment tests/test_parity_gpu.py::test_consumer_side_walk_on_the_baseline_shapes tests/test_parity_gpu.py::test_upload_scatter_and_patch"
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest $T -x -q -m gpu ) > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|error" gpurun_out/${tag}_memcheck.log | tail -6
( timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_exchange_gpu.py tests/test_parity_gpu.py::test_tier_x_list_length_boundaries tests/test_parity_gpu.py::test_tier_x_medium_lists_in_flat_batches -x -q -m gpu ) > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/${tag}_racecheck.log | tail -6
( timeout 400 python scripts/fuzz_parity.py 120 30000 ) > gpurun_out/${tag}_fuzz.log 2>&1; tail -2 gpurun_out/${tag}_fuzz.log
( timeout 200 python -m pytest tests/test_parity_gpu.py::test_back_to_back_frames_read_complete_counters -q -m gpu ) > gpurun_out/${tag}_stream_order.log 2>&1; tail -2 gpurun_out/${tag}_stream_order.log
( timeout 300 python scripts/soak_facade.py ) > gpurun_out/${tag}_soak_facade.log 2>&1; tail -2 gpurun_out/${tag}_soak_facade.log
