#!/bin/bash
# Static-scale fused producers: parity tests, then the static and dynamic batch-1 lines on one box.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_static.sh [tag]'
tag=${1:-st}
o=gpurun_out
mkdir -p $o
timeout 500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_modules.py -m gpu -x -q -k "static" > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_pytest.log
tail -5 $o/${tag}_pytest.log
timeout 400 python bench.py --mode static --no-cpu-baseline > $o/${tag}_static_b1.json 2> $o/${tag}_static_b1.err; tail -c 300 $o/${tag}_static_b1.err
python - <<PY
import json
d=json.loads(open("$o/${tag}_static_b1.json").readline())
print("static", d["ms_per_step"], d["fp16_baseline"]["ms_per_step"], d["fp16_baseline"]["speedup_w8a8_over_fp16"], d["gpu_launches_per_step"], d["launch_families"])
PY
timeout 400 python bench.py --no-cpu-baseline > $o/${tag}_dyn_b1.json 2> $o/${tag}_dyn_b1.err; tail -c 300 $o/${tag}_dyn_b1.err
python - <<PY
import json
d=json.loads(open("$o/${tag}_dyn_b1.json").readline())
print("dynamic", d["ms_per_step"], d["fp16_baseline"]["ms_per_step"], d["fp16_baseline"]["speedup_w8a8_over_fp16"], d["gpu_launches_per_step"])
PY
timeout 300 python tools/step_breakdown.py --batch 1 --mode static --out $o/${tag}_b1s.json > $o/${tag}_b1s.txt 2>&1
python tools/crit_path.py $o/${tag}_b1s.json 24
