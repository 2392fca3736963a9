#!/bin/bash
# Round 2, 2-GPU session: multigpu_check (independent scenes, ONE scene partitioned + consumer walk through peer mappings +
# NCCL cross-check + pulled instance runs, Tier R gather over NCCL) and the bench line at N=2 with its self-verification.
tag=${1:-r02b}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
( timeout 600 $TR 29541 tests/multigpu_check.py ) > gpurun_out/${tag}_multigpu_check.log 2>&1; echo "multigpu_check rc=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_multigpu_check.log | tail -12
# (async variant: see r02b first run)
( timeout 900 $TR 29543 bench.py --gpus $N --steps 100 --warmup 5 ) > gpurun_out/${tag}_bench_c3_${N}gpu.json 2> gpurun_out/${tag}_bench_c3_${N}gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${tag}_bench_c3_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_c3_${N}gpu.json").read().strip().splitlines()[-1])
    print("value", d["value"], d["ms_per_step"], "cull_only", d["cull_only"], "e2e", d["e2e"]["value"])
    print("pull", d.get("with_instance_pull"))
    print("verified", d.get("exchange_verified")); print(json.dumps(d.get("verification"))[:3000])
except Exception as e:
    print("bench parse failed", e)
PY
