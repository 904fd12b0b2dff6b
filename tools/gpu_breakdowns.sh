#!/bin/bash
tag=${1:-el3}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_modules.py tests/test_gpu_realsize.py -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log
python tools/step_breakdown.py --batch 1 --out $o/${tag}_b1.json > $o/${tag}_b1.txt 2>&1
python tools/crit_path.py $o/${tag}_b1.json 16
python tools/step_breakdown.py --model sd-turbo --batch 64 --mode static --out $o/${tag}_sd64s.json > $o/${tag}_sd64s.txt 2>&1
python tools/crit_path.py $o/${tag}_sd64s.json 12
python tools/step_breakdown.py --batch 8 --mode static --out $o/${tag}_b8s.json > $o/${tag}_b8s.txt 2>&1
python tools/crit_path.py $o/${tag}_b8s.json 12
