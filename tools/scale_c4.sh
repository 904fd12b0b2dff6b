#!/bin/bash
# BASELINE config 4: SD-Turbo W8A8 (static scales), global batch 64 sharded over 1/2/4/8 GPUs.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash tools/scale_c4.sh r02'
tag=${1:-r02}
nmax=${2:-8}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -gt $nmax ] && break
  if [ $n -eq 1 ]; then
    timeout 400 python bench.py --config 4 --no-cpu-baseline > gpurun_out/${tag}_c4_n$n.json 2> gpurun_out/${tag}_c4_n$n.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29520 + n)) bench.py --config 4 --gpus $n --no-cpu-baseline \
      > gpurun_out/${tag}_c4_n$n.json 2> gpurun_out/${tag}_c4_n$n.err
  fi
  tail -n 1 gpurun_out/${tag}_c4_n$n.json | head -c 200; echo
done
python - <<PY
import json
base = None
print("| GPUs | batch/GPU | ms/step | img/s | efficiency | FP16 ms/step | x FP16 |")
print("|---|---|---|---|---|---|---|")
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/${tag}_c4_n{n}.json").read().strip().splitlines()[-1])
    except Exception:
        continue
    base = base or d["value"]
    f = d["fp16_baseline"]
    print(f"| {n} | {64 // n} | {d['ms_per_step']:.2f} | {d['value']:.1f} | {d['value'] / base / n:.3f} | "
          f"{f['ms_per_step']:.2f} | {f['speedup_w8a8_over_fp16']:.3f} |")
PY
