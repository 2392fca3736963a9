#!/bin/bash
# 4 GPUs, final code of round 2: multigpu_check, bench c3 (one-launch frame close, two-launch form beside it, e2e with the read-back stream), bench c5.
tag=r02x
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
( timeout 600 $TR 29571 tests/multigpu_check.py ) > gpurun_out/${tag}_multigpu_check_$N.log 2>&1; echo "check rc=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_multigpu_check_$N.log | tail -3 | cut -c1-400
for w in c3 c5; do
  ( timeout 600 $TR 29573 bench.py --gpus $N --workload $w --steps 200 --warmup 5 --split-sync ) > gpurun_out/${tag}_bench_${w}_${N}gpu.json 2> gpurun_out/${tag}_bench_${w}_${N}gpu.err; echo "bench $w rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${w}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$w value", d["value"], d["ms_per_step"], "cull_only", d["cull_only"]["value"], d["cull_only"]["ms_per_step"], "split", d["with_split_sync"]["ms_per_step"], d["with_split_sync"]["default_measured_again_ms_per_step"],
          "verified", d.get("exchange_verified"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "pull", d["with_instance_pull"]["ms_per_step"])
except Exception as e:
    print("$w parse failed", e)
PY
  grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_bench_${w}_${N}gpu.err | tail -3 | cut -c1-300
done
