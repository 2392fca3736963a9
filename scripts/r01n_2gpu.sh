#!/bin/bash
# Two-GPU check of the final code state: fused peer exchange against the oracle (incl. medium lists), MG_ASYNC soak, bench line.
tag=${1:-r01n}
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q ) > gpurun_out/${tag}_pytest_2gpu.log 2>&1; tail -4 gpurun_out/${tag}_pytest_2gpu.log
MG_FRAMES=101 MG_ASYNC=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multigpu_check.py 2>&1 | tail -2 | tee gpurun_out/${tag}_mg_async.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/${tag}_bench_c3_2gpu_peer.json 2> gpurun_out/${tag}_bench_c3_2gpu_peer.err; tail -c 900 gpurun_out/${tag}_bench_c3_2gpu_peer.json
