#!/bin/bash
# Round 2, 8-GPU session: multigpu_check at 8 ranks, the bench line at N=8 exactly as the driver runs it (ONE 800 M-instance
# scene partitioned, self-verifying), BASELINE configs[4] (ONE 1 B-instance scene, 125 M per GPU) and N=4.
tag=${1:-r02f}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 600 $TR --nproc-per-node 8 --master-port 29551 tests/multigpu_check.py ) > gpurun_out/${tag}_multigpu_check_8.log 2>&1; echo "multigpu_check rc=$?"; grep "multigpu_check\|rank [0-9]*:" gpurun_out/${tag}_multigpu_check_8.log | tail -8
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    v = d.get("verification", {})
    print(sys.argv[2], "value", d["value"], d["ms_per_step"], "cull_only", d["cull_only"]["value"], d["cull_only"]["ms_per_step"], "e2e", d["e2e"]["value"],
          "pull", d.get("with_instance_pull", {}).get("value"), d.get("with_instance_pull", {}).get("ms_per_step"), "verified", d.get("exchange_verified"))
    print("   ", json.dumps({k: v.get(k) for k in ("counts_equal_cull_only", "draws_for_whole_scene", "tier_r_gather")}), json.dumps(v.get("consumer_walk", {}))[:600], json.dumps(v.get("nccl_cross_check"))[:200])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
( timeout 900 $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench_c3_8gpu.json 2> gpurun_out/${tag}_bench_c3_8gpu.err; echo "bench c3 x8 rc=$?"; show gpurun_out/${tag}_bench_c3_8gpu.json "c3 x8"
( timeout 900 $TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --workload c5 --steps 100 --warmup 5 ) > gpurun_out/${tag}_bench_c5_8gpu.json 2> gpurun_out/${tag}_bench_c5_8gpu.err; echo "bench c5 x8 rc=$?"; show gpurun_out/${tag}_bench_c5_8gpu.json "c5 x8"
( timeout 900 $TR --nproc-per-node 4 --master-port 29554 bench.py --gpus 4 --steps 100 --warmup 5 ) > gpurun_out/${tag}_bench_c3_4gpu.json 2> gpurun_out/${tag}_bench_c3_4gpu.err; echo "bench c3 x4 rc=$?"; show gpurun_out/${tag}_bench_c3_4gpu.json "c3 x4"
( timeout 900 $TR --nproc-per-node 8 --master-port 29555 bench.py --gpus 8 --steps 100 --warmup 5 --exchange nccl --no-verify ) > gpurun_out/${tag}_bench_c3_8gpu_nccl.json 2> gpurun_out/${tag}_bench_c3_8gpu_nccl.err; echo "bench c3 x8 nccl rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench_c3_8gpu_nccl.json").read().strip().splitlines()[-1])
    print("c3 x8 nccl-after-cull value", d["value"], d["ms_per_step"], "cull_only", d["cull_only"])
except Exception as e:
    print("nccl line failed", e)
PY
grep -h "Error\|error" gpurun_out/${tag}_*.err | grep -v "^W0" | head -5
