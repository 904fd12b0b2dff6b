#!/bin/bash
# The round's measured evidence in one GPU session (outputs under gpurun_out/, copied to profiles/):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/profile_round.sh r02'
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
timeout 500 python bench.py > $o/${tag}_bench_c2.json 2> $o/${tag}_bench_c2.err
timeout 500 python bench.py --config 3 --no-cpu-baseline > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err
timeout 600 python bench.py --config 5 --sweep-out $o/${tag}_sweep.json > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err
# launch list of the bench's timed region (cold-cache, serialised: compare SHARES)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $o/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp16 \
  > $o/${tag}_launches_bench.log 2>&1
# one warm launch of every batch-1 kernel family, full metric set
timeout 600 ncu --set full --clock-control none --import-source on -o $o/${tag}_repr_full -f \
  python tools/ncu_repr.py > $o/${tag}_repr.log 2>&1
ncu -i $o/${tag}_repr_full.ncu-rep --page raw --csv > $o/${tag}_repr_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $o/${tag}_repr_full_raw.csv > $o/${tag}_repr_full_summary.txt
# the persistent kernel on a tensor-bound and on a mid-size batch-8 shape
for shape in 8192,2560,2560 2048,1280,1280; do
  n=${shape//,/_}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_i8_persist -s 2 -c 1 \
    -o $o/${tag}_persist_$n -f python tools/ncu_persist.py $shape > $o/${tag}_persist_$n.log 2>&1
  ncu -i $o/${tag}_persist_$n.ncu-rep --page raw --csv > $o/${tag}_persist_${n}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $o/${tag}_persist_${n}_raw.csv > $o/${tag}_persist_${n}_summary.txt
done
timeout 300 python tools/tops_sweep.py > $o/${tag}_tops_sweep.txt 2>&1
python tools/step_breakdown.py --batch 1 --out $o/${tag}_b1.json > $o/${tag}_breakdown_b1.txt 2>&1
python tools/crit_path.py $o/${tag}_b1.json 40 > $o/${tag}_critpath_b1.txt
python tools/step_breakdown.py --batch 8 --out $o/${tag}_b8.json > $o/${tag}_breakdown_b8.txt 2>&1
python tools/crit_path.py $o/${tag}_b8.json 40 > $o/${tag}_critpath_b8.txt
python tools/step_breakdown.py --batch 8 --fp16 --out $o/${tag}_b8_fp16.json > $o/${tag}_breakdown_b8_fp16.txt 2>&1
rm -f $o/${tag}_repr_full.ncu-rep   # 20+ MB; the raw CSV and the summary are kept
ls -la $o | grep ${tag}_ | awk '{print $5, $9}'
