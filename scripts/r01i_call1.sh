#!/bin/bash
# One B200 call: GPU test-suite, A/B of the list-kernel variants over list lengths, default bench line, CPU arm.
tag=${1:-r01i}
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
( time timeout 500 python scripts/ab_list_kernels.py ) > gpurun_out/${tag}_ab.jsonl 2> gpurun_out/${tag}_ab.err; cat gpurun_out/${tag}_ab.jsonl | cut -c1-260; tail -5 gpurun_out/${tag}_ab.err
( time timeout 400 python bench.py ) > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; tail -c 1500 gpurun_out/${tag}_bench_c3.json; tail -4 gpurun_out/${tag}_bench_c3.err
( time timeout 300 python bench.py --impl reference ) > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; tail -c 400 gpurun_out/${tag}_bench_reference.json
