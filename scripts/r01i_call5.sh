#!/bin/bash
tag=${1:-r01m}
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -6 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( timeout 200 python scripts/fuzz_parity.py 30 5000 ) > gpurun_out/${tag}_fuzz.log 2>&1; tail -2 gpurun_out/${tag}_fuzz.log
( time timeout 500 python scripts/ab_list_kernels.py --variants 2,4,5 --lengths 33,64,65,100,1000 --rounds 3 ) > gpurun_out/${tag}_ab.jsonl 2> gpurun_out/${tag}_ab.err; cut -c1-330 gpurun_out/${tag}_ab.jsonl; tail -5 gpurun_out/${tag}_ab.err
( time timeout 400 python bench.py --no-cpu-baseline ) > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; tail -c 1000 gpurun_out/${tag}_bench_c3.json; tail -4 gpurun_out/${tag}_bench_c3.err
