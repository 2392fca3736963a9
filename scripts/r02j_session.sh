#!/bin/bash
# Round 2, final 1-GPU evidence run with the product code (64-thread CTAs in the thread-per-drawable kernels): GPU tests, smoke,
# the two driver commands, facade_bench, launch lists and full ncu captures of the dominant kernels.
tag=r02j
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
t0=$(date +%s)
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_driver.json 2> gpurun_out/${tag}_bench_c3_driver.err
echo "default bench.py took $(( $(date +%s) - t0 )) s"; tail -c 400 gpurun_out/${tag}_bench_c3_driver.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02j_bench_c3_driver.json").read().strip().splitlines()[-1])
    print("c3", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"), d["kernels_ms"])
    for w, c in d.get("workloads", {}).items():
        print(w, {k: c.get(k) for k in ("value", "ms_per_step", "error")}, "e2e", c.get("e2e", {}).get("value"), "frac", c.get("roofline", {}).get("frac"), c.get("roofline", {}).get("frac_of_line_granular_floor"), "tier_r", c.get("tier_r"))
    print("facade", d.get("e2e_facade"))
except Exception as e:
    print("parse failed", e)
PY
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_reference.json 2>/dev/null; tail -c 200 gpurun_out/${tag}_bench_reference.json
for sc in c1 c2; do
  ( timeout 400 cadr_b200/host/bin/facade_bench 0 $sc 200 ) > gpurun_out/${tag}_facade_$sc.json 2> gpurun_out/${tag}_facade_$sc.err; tail -c 700 gpurun_out/${tag}_facade_$sc.json | cut -c1-600
done
for w in c3 c2; do
  timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_$w.csv \
      python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/${tag}_ncu_launches_$w.log 2>&1
done
cap() { local name=$1 rx=$2 skip=$3; shift 3
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${tag}_$name python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > /dev/null 2>&1; }
cap cullSmallKernel_fused_c2 cullSmall 3 --workload c2
cap cullSmallKernel_fused_c1 cullSmall 3 --workload c1
cap cullListWarpKernel_c3 cullListWarp 3
for l in 33 64 100 200; do
  ( timeout 200 python bench.py --instances $l --drawables $((100000000 / l)) --no-cpu-baseline --no-workloads --steps 100 ) > gpurun_out/${tag}_len_$l.json 2>/dev/null
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_len_$l.json").read().strip().splitlines()[-1]); print("list length $l:", d["value"], d["ms_per_step"], d["kernels_ms"])
except Exception as e: print("len $l failed", e)
PY
done
ls gpurun_out/${tag}_*.ncu-rep
