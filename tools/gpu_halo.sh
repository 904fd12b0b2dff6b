#!/bin/bash
tag=${1:-halo}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_persist.py -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_pytest.log
tail -15 $o/${tag}_pytest.log
timeout 300 python tools/conv_halo_ab.py 2>&1 | tee $o/${tag}_ab.txt
