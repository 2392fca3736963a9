python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_c3_new.log 2>&1; tail -c 1200 gpurun_out/bench_c3_new.log
bash scripts/sweep_list_length.sh new
