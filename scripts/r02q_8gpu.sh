#!/bin/bash
# 8 GPUs: multigpu_check incl. part D (deferred wait) and the bench line with the deferred-wait series (driver form and 200 steps).
tag=r02q
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
( timeout 600 $TR 29581 tests/multigpu_check.py ) > gpurun_out/${tag}_multigpu_check_8.log 2>&1; echo "check rc=$?"; grep "multigpu_check\|rank [0-9]*:" gpurun_out/${tag}_multigpu_check_8.log | tail -4 | cut -c1-400
for steps in 20 200; do
( timeout 900 $TR 29582 bench.py --gpus 8 --steps $steps --warmup 5 ) > gpurun_out/${tag}_bench_c3_8gpu_$steps.json 2> gpurun_out/${tag}_bench_c3_8gpu_$steps.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_c3_8gpu_$steps.json").read().strip().splitlines()[-1])
    print("steps $steps: value", d["value"], d["ms_per_step"], "cull_only", d["cull_only"]["value"], d["cull_only"]["ms_per_step"], "deferred", d["with_deferred_wait"]["value"], d["with_deferred_wait"]["ms_per_step"], d["with_deferred_wait"]["last_frame_cross_checked_over_nccl"], "e2e", d["e2e"]["value"], "pull", d["with_instance_pull"]["value"], "verified", d.get("exchange_verified"))
except Exception as e:
    print("parse failed", e)
PY
done
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_bench_c3_8gpu_20.err | tail -4 | cut -c1-300
