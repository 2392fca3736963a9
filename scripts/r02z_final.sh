#!/bin/bash
# Round 2, last code state, one GPU: what the driver runs at round end (GPU tests, smoke, both bench arms), kept as evidence,
# + the facade benches, the ncu launch list of the bench command and a full capture of the dominant kernel.
tag=r02z
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
( timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_reference.json 2>/dev/null; tail -c 250 gpurun_out/${tag}_bench_reference.json; echo
t0=$(date +%s)
( timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_driver.json 2> gpurun_out/${tag}_bench_c3_driver.err
echo "default bench.py took $(( $(date +%s) - t0 )) s"; tail -c 300 gpurun_out/${tag}_bench_c3_driver.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02z_bench_c3_driver.json").read().strip().splitlines()[-1])
    print("c3", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["read_stream_ceiling"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"), "launches", d["gpu_launches"], d["clocks"])
    for w, c in d.get("workloads", {}).items():
        print(w, {k: c.get(k) for k in ("value", "ms_per_step", "error")}, "e2e", c.get("e2e", {}).get("value"), "frac", c.get("roofline", {}).get("frac"), c.get("roofline", {}).get("frac_of_line_granular_floor"))
    print("facade", {k: d["e2e_facade"].get(k) for k in ("value", "ms_per_step", "gpuDrawableProcessing_ms", "error")})
except Exception as e:
    print("parse failed", e)
PY
( timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-workloads --no-cpu-baseline ) > gpurun_out/${tag}_bench_c3_200.json 2>/dev/null
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02z_bench_c3_200.json").read().strip().splitlines()[-1])
    print("c3 200 steps", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "resident", d["e2e_resident_list"]["value"], "frac", d["roofline"]["frac"], d["kernels_ms"])
except Exception as e:
    print("parse failed", e)
PY
for sc in c1 c2 c3; do
  ( timeout 400 cadr_b200/host/bin/facade_bench 0 $sc 200 ) > gpurun_out/${tag}_facade_$sc.json 2> gpurun_out/${tag}_facade_$sc.err; tail -c 500 gpurun_out/${tag}_facade_$sc.json | cut -c1-500; echo
done
( timeout 900 cadr_b200/host/bin/facade_bench 0 c4 30 ) > gpurun_out/${tag}_facade_c4.json 2> gpurun_out/${tag}_facade_c4.err; tail -c 600 gpurun_out/${tag}_facade_c4.json; echo
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/${tag}_ncu_launches_c3.log 2>&1
tail -8 gpurun_out/${tag}_launches_c3.csv | cut -c1-200
timeout 200 ncu --set full --import-source on --clock-control none -k regex:cullListWarp -s 3 -c 1 -f -o gpurun_out/${tag}_cullListWarpKernel_c3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > /dev/null 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
