#!/bin/bash
o=gpurun_out
echo "== new"; python tools/gn_bench.py 2>&1 | tee $o/gn_new.txt
echo "== prev"; MIXDQ_B200_LIB=$PWD/mixdq_b200/libmixdq_prev.so python tools/gn_bench.py 2>&1 | tee $o/gn_prev.txt
