#!/bin/bash
# Round 2, 2-GPU session with the final code: the bench line at N=2 as the driver runs it, the other workloads at N=2
# (c2 / c1 run one scene per rank: their shapes are not partitioned), multigpu_check queued without host sync.
tag=${1:-r02m}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value", d["value"], d["ms_per_step"], "cull_only", d["cull_only"]["value"], "e2e", d["e2e"]["value"],
          "pull", d.get("with_instance_pull", {}).get("value"), d.get("with_instance_pull", {}).get("ms_per_step"), "verified", d.get("exchange_verified"))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
( timeout 600 $TR 29561 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_2gpu.json 2> gpurun_out/${tag}_bench_c3_2gpu.err; echo "c3 rc=$?"; show gpurun_out/${tag}_bench_c3_2gpu.json "c3 x2"
for w in c5 c2 c1 c4; do
  ( timeout 600 $TR 29562 bench.py --gpus 2 --workload $w --steps 50 --warmup 5 ) > gpurun_out/${tag}_bench_${w}_2gpu.json 2> gpurun_out/${tag}_bench_${w}_2gpu.err; echo "$w rc=$?"; show gpurun_out/${tag}_bench_${w}_2gpu.json "$w x2"
  grep -h "Error\|error" gpurun_out/${tag}_bench_${w}_2gpu.err | grep -v "^W0" | head -3
done
( MG_ASYNC=1 MG_FRAMES=60 timeout 600 $TR 29563 tests/multigpu_check.py ) > gpurun_out/${tag}_multigpu_async.log 2>&1; echo "async rc=$?"; grep multigpu_check gpurun_out/${tag}_multigpu_async.log | cut -c1-200
( timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu ) > gpurun_out/${tag}_pytest_2gpu.log 2>&1; tail -2 gpurun_out/${tag}_pytest_2gpu.log
