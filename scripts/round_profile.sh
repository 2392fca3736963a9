#!/bin/bash
# Evidence run on one B200: GPU tests, bench lines of every workload, ncu launch lists of the timed region and full
# captures of the dominant kernels.  usage: scripts/round_profile.sh <tag>   (outputs under gpurun_out/)
tag=${1:-r01h}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in c3 c1 c2 c4 c5; do
  python bench.py --workload $w $([ $w = c3 ] || echo --no-cpu-baseline) > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  tail -c 400 gpurun_out/${tag}_bench_$w.json
done
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2>/dev/null
python bench.py --list-bounds --no-cpu-baseline > gpurun_out/${tag}_bench_c3_list_bounds.json 2>/dev/null
python bench.py --workload c4 --list-bounds --no-cpu-baseline > gpurun_out/${tag}_bench_c4_list_bounds.json 2>/dev/null
for w in c3 c2; do
  ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_$w.csv \
      python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches_$w.log 2>&1
done
ncu --set full --import-source on --clock-control none -k regex:cullListWarp -c 1 -f -o gpurun_out/${tag}_cullListWarpKernel_c3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
CADR_B200_CULL_VARIANT=3 ncu --set full --import-source on --clock-control none -k regex:cullListRing -c 1 -f -o gpurun_out/${tag}_cullListRingKernel_c3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:cullListWarp -c 1 -f -o gpurun_out/${tag}_cullListWarpKernel_64 \
    python bench.py --instances 64 --drawables 1562500 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out/${tag}_*
