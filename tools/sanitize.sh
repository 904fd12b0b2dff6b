#!/bin/bash
# compute-sanitizer over the mbarrier / TMEM / cluster-workspace kernels.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/sanitize.sh [tag]'
tag=${1:-r02}
mkdir -p gpurun_out
SEL='every_tile_width or split_k or geglu or w4_packed or split_shortcut'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_fused.py tests/test_gpu_persist.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/${tag}_sanitizer_${tool}.log
  tail -5 gpurun_out/${tag}_sanitizer_${tool}.log
done
