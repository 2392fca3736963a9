#!/bin/bash
# Round 2, GPU session (one B200): the facade's frame loop on the device, the default bench line exactly as the driver runs it
# (with the workloads block and e2e_facade), the reference arm, and the ncu evidence of the round (launch lists of the
# timed region + full captures of the dominant kernels of every workload).
tag=r02e
mkdir -p gpurun_out
for sc in c1 c2 c3; do
  ( timeout 400 cadr_b200/host/bin/facade_bench 0 $sc 200 ) > gpurun_out/${tag}_facade_$sc.json 2> gpurun_out/${tag}_facade_$sc.err; tail -c 1100 gpurun_out/${tag}_facade_$sc.json; tail -2 gpurun_out/${tag}_facade_$sc.err
done
t0=$(date +%s)
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_driver.json 2> gpurun_out/${tag}_bench_c3_driver.err
echo "default bench.py took $(( $(date +%s) - t0 )) s"; tail -c 600 gpurun_out/${tag}_bench_c3_driver.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02e_bench_c3_driver.json").read().strip().splitlines()[-1])
    print("c3", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"))
    for w, c in d.get("workloads", {}).items():
        print(w, {k: c.get(k) for k in ("value", "ms_per_step", "error")}, "e2e", c.get("e2e", {}).get("value"), "frac", c.get("roofline", {}).get("frac"), c.get("roofline", {}).get("frac_of_line_granular_floor"))
    print("facade", d.get("e2e_facade"))
except Exception as e:
    print("parse failed", e)
PY
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/${tag}_bench_reference.json
# ---- ncu: launch lists of the timed region
for w in c3 c2; do
  timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_$w.csv \
      python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/${tag}_ncu_launches_$w.log 2>&1
done
# ---- ncu: full captures
cap() {  # cap <name> <kernel regex> <skip> <bench args...>
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${tag}_$name python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > /dev/null 2>&1
}
cap cullListWarpKernel_c3 cullListWarp 3
cap cullSmallKernel_fused_c2 cullSmall 3 --workload c2
cap cullSmallKernel_fused_c1 cullSmall 3 --workload c1
cap cullListWarpKernel_c5 cullListWarp 3 --workload c5
cap scatterCopyKernel_c4 scatterCopy 3 --workload c4
cap cullMediumKernel_64 cullMedium 3 --instances 64 --drawables 1562500
cap processDrawablesKernel_c2 processDrawables 1 --workload c2
# the staged experiment kernel, to see what it stalls on
CADR_B200_SMALL_STAGED=2 timeout 200 ncu --set full --import-source on --clock-control none -k regex:cullSmallStaged -s 3 -c 1 -f -o gpurun_out/${tag}_cullSmallStagedKernel128_c2 \
    python scripts/exp_bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > /dev/null 2>&1
CADR_B200_CULL_VARIANT=6 timeout 200 ncu --set full --import-source on --clock-control none -k regex:cullListRingPair -s 3 -c 1 -f -o gpurun_out/${tag}_cullListRingPairKernel_c3 \
    python scripts/exp_bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
