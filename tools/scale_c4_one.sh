#!/bin/bash
# one point of the config-4 strong-scaling table: bash tools/scale_c4_one.sh N [tag]
n=$1; tag=${2:-r02f}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
  --master-port $((29520 + n)) bench.py --config 4 --gpus $n --no-cpu-baseline \
  > gpurun_out/${tag}_c4_n$n.json 2> gpurun_out/${tag}_c4_n$n.err
tail -n 1 gpurun_out/${tag}_c4_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print($n, round(d['ms_per_step'],2), round(d['value'],1), round(d['fp16_baseline']['ms_per_step'],2), round(d['fp16_baseline']['speedup_w8a8_over_fp16'],3))"
tail -c 300 gpurun_out/${tag}_c4_n$n.err
