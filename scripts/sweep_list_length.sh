#!/bin/bash
# Frame rate as a function of MatrixList length at a fixed 100 M instances (one bench.py line per length).
# usage: [TOTAL=instances] scripts/sweep_list_length.sh <tag> [lengths...]   -> gpurun_out/sweep_<tag>.jsonl
tag=$1; shift
lens=${@:-"2 8 16 32 33 48 64 100 200 500 1000 5000"}
out=gpurun_out/sweep_$tag.jsonl; : > $out
for n in $lens; do
  d=$(( ${TOTAL:-100000000} / n ))
  python bench.py --workload c3 --instances $n --drawables $d --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
done
python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    k = d["kernels_ms"]
    print(d["config"]["workload"].split(",")[1].strip()[:60], "| ms/step", d["ms_per_step"], "| G inst/s", round(d["value"]/1e3,1), "|", k)
PY
