#!/bin/bash
# 2 GPUs: the exchange closed by ONE launch (publish + wait) - multigpu_check (all parts), bench c3 with the two-launch form beside it.
tag=r03c
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
( timeout 600 $TR 29571 tests/multigpu_check.py ) > gpurun_out/${tag}_multigpu_check.log 2>&1; echo "check rc=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_multigpu_check.log | tail -6 | cut -c1-400
( MG_ASYNC=1 MG_FRAMES=30 timeout 600 $TR 29572 tests/multigpu_check.py ) > gpurun_out/${tag}_multigpu_async.log 2>&1; echo "async rc=$?"
( timeout 600 $TR 29573 bench.py --gpus 2 --steps 200 --warmup 5 --split-sync ) > gpurun_out/${tag}_bench_c3_2gpu.json 2> gpurun_out/${tag}_bench_c3_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r03c_bench_c3_2gpu.json").read().strip().splitlines()[-1])
    print("value", d["value"], d["ms_per_step"], "cull_only", d["cull_only"], "split", d.get("with_split_sync"), "verified", d.get("exchange_verified"), "e2e", d["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
PY
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_bench_c3_2gpu.err | tail -5 | cut -c1-300
( timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q ) > gpurun_out/${tag}_pytest_2gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_2gpu.log
