#!/bin/bash
# What the next GPU session should run first (left over from round 1, whose GPU budget ended before these could run).
#   one GPU :  scripts/next_gpu_session.sh single <tag>
#   two GPUs:  scripts/next_gpu_session.sh multi <tag>     (gpurun --gpus 2; also fine with 4)
mode=${1:-single}; tag=${2:-r02a}
mkdir -p gpurun_out
if [ "$mode" = single ]; then
  ( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
  # the facade itself on the device at the reference's full grid and at configs[1]'s size: host phases + the library's GPU time
  for scene in independent shared-geometry showhide instanced; do
    ( timeout 120 cadr_b200/host/bin/boxes_scene_test 0 - $scene 100 6 ) > gpurun_out/${tag}_facade_$scene.log 2>&1; tail -4 gpurun_out/${tag}_facade_$scene.log
  done
  ( timeout 200 cadr_b200/host/bin/boxes_scene_test 0 - shared-geometry 216 6 ) > gpurun_out/${tag}_facade_10M.log 2>&1; tail -4 gpurun_out/${tag}_facade_10M.log
  ( timeout 200 python bench.py ) > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; tail -c 600 gpurun_out/${tag}_bench_c3.json
else
  n=$(python -c "import torch; print(torch.cuda.device_count())")
  MG_TIER_R=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 tests/multigpu_check.py 2>&1 | tail -3 | tee gpurun_out/${tag}_mg_tier_r.log
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/${tag}_bench_c3_${n}gpu_peer.json 2> gpurun_out/${tag}_bench_c3_${n}gpu_peer.err; tail -c 500 gpurun_out/${tag}_bench_c3_${n}gpu_peer.json
fi
