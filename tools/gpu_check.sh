#!/bin/bash
# One GPU session: parity tests, the bench line, the in-graph critical-path table.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_check.sh [tag]'
tag=${1:-check}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.err
head -c 300 gpurun_out/${tag}_bench.json; echo
timeout 300 python tools/step_breakdown.py --out gpurun_out/${tag}_breakdown_w8a8.json > gpurun_out/${tag}_breakdown_w8a8.txt 2>&1
python tools/crit_path.py gpurun_out/${tag}_breakdown_w8a8.json 30
python -c "from __graft_entry__ import smoke; smoke()"
