#!/bin/bash
# A/B of one environment switch inside ONE gpurun call (same box): tools/ab_bench.sh VAR v1 v2 ...
# prints ms_per_step of `bench.py --no-fp16 --no-cpu-baseline` twice per value, interleaved.
var=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    ms=$(env $var=$v python bench.py --steps 30 --warmup 5 --no-fp16 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_per_step'])")
    echo "$var=$v rep$rep ms_per_step=$ms"
  done
done
