#!/bin/bash
# A/B of one environment switch inside ONE gpurun call (same box): tools/ab_bench.sh VAR v1 v2 ...
# prints ms_per_step (dynamic and static scales) of `bench.py --no-fp16 --no-cpu-baseline` twice per
# value, interleaved. Extra bench arguments: BENCH_ARGS="--batch 8".
var=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    ms=$(env $var=$v python bench.py --steps 30 --warmup 5 --no-fp16 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); s=d.get('static_scales'); print(round(d['ms_per_step'],4), (round(s['ms_per_step'],4) if s else None))")
    echo "$var=$v rep$rep ms_per_step(dynamic, static)=$ms"
  done
done
