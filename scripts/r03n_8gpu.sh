#!/bin/bash
# final tree, N GPUs: multigpu_check + the driver-form bench command (c3) under torchrun
tag=r03n
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
( timeout 600 $TR 29571 tests/multigpu_check.py ) > gpurun_out/${tag}_multigpu_check_$N.log 2>&1; echo "check rc=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_multigpu_check_$N.log | tail -2 | cut -c1-400
( timeout 600 $TR 29573 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_${N}gpu.json 2> gpurun_out/${tag}_bench_c3_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_c3_${N}gpu.json").read().strip().splitlines()[-1])
    print("c3 value", d["value"], d["ms_per_step"], "cull_only", d["cull_only"]["value"], d["cull_only"]["ms_per_step"], "verified", d.get("exchange_verified"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "pull", d["with_instance_pull"]["ms_per_step"], d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_bench_c3_${N}gpu.err | tail -3 | cut -c1-300
