#!/bin/bash
# Evidence run for the final code state of round 1 on one B200 (outputs under gpurun_out/): GPU tests, smoke, the
# default bench line, A/B of the medium-list path, ncu launch list of the bench step, the CPU reference arm and one
# full capture of cullMediumKernel.  Every leg has its own timeout; the most important ones come first.
tag=${1:-r01n}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -5 gpurun_out/${tag}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${tag}_smoke.log
( time timeout 200 python bench.py ) > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; tail -c 1500 gpurun_out/${tag}_bench_c3.json; tail -4 gpurun_out/${tag}_bench_c3.err
( time timeout 150 python scripts/ab_list_kernels.py --variants 2,4 --lengths 33,48,64,1000 --rounds 2 ) > gpurun_out/${tag}_ab.jsonl 2> gpurun_out/${tag}_ab.err; cut -c1-300 gpurun_out/${tag}_ab.jsonl; tail -4 gpurun_out/${tag}_ab.err
timeout 100 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches_c3.log 2>&1
tail -12 gpurun_out/${tag}_launches_c3.csv | cut -c1-200
( time timeout 150 python bench.py --impl reference ) > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; tail -c 600 gpurun_out/${tag}_bench_reference.json
timeout 100 ncu --set full --import-source on --clock-control none -k regex:cullMedium -c 1 -f -o gpurun_out/${tag}_cullMediumKernel_64 \
    python bench.py --instances 64 --drawables 1562500 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
( timeout 100 python bench.py --instances 64 --drawables 1562500 --no-cpu-baseline ) > gpurun_out/${tag}_bench_64.json 2>/dev/null; tail -c 700 gpurun_out/${tag}_bench_64.json
ls -la gpurun_out/ | tail -15
