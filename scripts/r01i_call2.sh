#!/bin/bash
# One B200 call: GPU test-suite, randomised parity soak, A/B of the list-kernel variants over list lengths, c3 and c2 bench lines.
tag=${1:-r01j}
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -15 gpurun_out/${tag}_pytest.log
( timeout 200 python scripts/fuzz_parity.py 45 ) > gpurun_out/${tag}_fuzz.log 2>&1; tail -3 gpurun_out/${tag}_fuzz.log
( time timeout 500 python scripts/ab_list_kernels.py --variants 2,4 --lengths 33,40,48,64,65,100,1000 ) > gpurun_out/${tag}_ab.jsonl 2> gpurun_out/${tag}_ab.err; cut -c1-330 gpurun_out/${tag}_ab.jsonl; tail -5 gpurun_out/${tag}_ab.err
( time timeout 400 python bench.py ) > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; tail -c 1300 gpurun_out/${tag}_bench_c3.json; tail -4 gpurun_out/${tag}_bench_c3.err
( timeout 400 python bench.py --workload c2 --no-cpu-baseline --steps 200 ) > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err; tail -c 900 gpurun_out/${tag}_bench_c2.json; tail -4 gpurun_out/${tag}_bench_c2.err
