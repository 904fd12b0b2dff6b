#!/bin/bash
# End-of-round evidence on one box: full GPU suite, the bench lines, launch list, ncu of the new
# kernels, sanitizer over the kernels changed late in the round.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_final.sh r02f'
tag=${1:-r02f}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log
timeout 600 python bench.py > $o/${tag}_bench_c2.json 2> $o/${tag}_bench_c2.err; tail -c 300 $o/${tag}_bench_c2.err
timeout 400 python bench.py --config 4 --no-cpu-baseline > $o/${tag}_bench_c4_n1.json 2> $o/${tag}_bench_c4_n1.err; tail -c 300 $o/${tag}_bench_c4_n1.err
timeout 400 python bench.py --config 3 --no-cpu-baseline > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err; tail -c 300 $o/${tag}_bench_c3.err
timeout 400 python bench.py --batch 8 --mode static --no-cpu-baseline > $o/${tag}_bench_b8_static.json 2> $o/${tag}_bench_b8_static.err
timeout 400 python bench.py --batch 8 --no-cpu-baseline --no-static > $o/${tag}_bench_b8_dynamic.json 2> $o/${tag}_bench_b8_dynamic.err
timeout 600 python bench.py --config 5 --sweep-out $o/${tag}_sweep.json > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err; tail -c 300 $o/${tag}_bench_c5.err
python - <<PY
import json
for f in ("bench_c2", "bench_c4_n1", "bench_c3", "bench_b8_static", "bench_b8_dynamic", "bench_c5"):
    try:
        d = json.loads(open("$o/${tag}_%s.json" % f).readline())
        print(f, round(d["ms_per_step"], 3), round(d["value"], 1), d.get("fp16_baseline") and round(d["fp16_baseline"]["speedup_w8a8_over_fp16"], 3),
              d.get("static_scales") and (round(d["static_scales"]["ms_per_step"], 3), round(d["static_scales"]["speedup_over_fp16"], 3)),
              "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step", 0), 3), d["roofline"]["bound"], round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
# launch list of the bench's timed region (cold-cache, serialised: compare SHARES)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $o/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp16 --no-static \
  > $o/${tag}_launches_bench.log 2>&1
# the halo convolution and the new elementwise kernels, full metric set
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_i8_persist -s 3 -c 1 \
  -o $o/${tag}_halo_conv -f python tools/conv_halo_ab.py > $o/${tag}_halo_conv.log 2>&1
ncu -i $o/${tag}_halo_conv.ncu-rep --page raw --csv > $o/${tag}_halo_conv_raw.csv 2>/dev/null
python tools/ncu_summary.py $o/${tag}_halo_conv_raw.csv | tee $o/${tag}_halo_conv_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gn_|ln_minmax|quant_" -s 13 -c 13 \
  -o $o/${tag}_elem_sd64 -f python tools/ncu_elem.py sd64 > $o/${tag}_elem_sd64.log 2>&1
ncu -i $o/${tag}_elem_sd64.ncu-rep --page raw --csv > $o/${tag}_elem_sd64_raw.csv 2>/dev/null
python tools/ncu_summary.py $o/${tag}_elem_sd64_raw.csv | tee $o/${tag}_elem_sd64_summary.txt
rm -f $o/${tag}_halo_conv.ncu-rep $o/${tag}_elem_sd64.ncu-rep
python tools/step_breakdown.py --batch 1 --out $o/${tag}_b1.json > $o/${tag}_b1.txt 2>&1
python tools/crit_path.py $o/${tag}_b1.json 40 > $o/${tag}_critpath_b1.txt
python tools/step_breakdown.py --batch 1 --mode static --out $o/${tag}_b1s.json > $o/${tag}_b1s.txt 2>&1
python tools/crit_path.py $o/${tag}_b1s.json 40 > $o/${tag}_critpath_b1_static.txt
python tools/step_breakdown.py --model sd-turbo --batch 64 --mode static --out $o/${tag}_sd64s.json > $o/${tag}_sd64s.txt 2>&1
python tools/crit_path.py $o/${tag}_sd64s.json 40 > $o/${tag}_critpath_sd64s.txt
python tools/step_breakdown.py --batch 8 --out $o/${tag}_b8.json > $o/${tag}_b8.txt 2>&1
python tools/crit_path.py $o/${tag}_b8.json 40 > $o/${tag}_critpath_b8.txt
timeout 300 python tools/tops_sweep.py > $o/${tag}_tops_sweep.txt 2>&1
# sanitizer: halo convolutions, static producers, GroupNorm statistics, 4-channel convolution
SEL='halo or static or groupnorm or other_geometries or layernorm_quant'
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_fused.py tests/test_gpu_persist.py -m gpu -q -x -k "$SEL" \
    > $o/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a $o/${tag}_sanitizer_${tool}.log
  tail -3 $o/${tag}_sanitizer_${tool}.log
done
python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -2
rm -f $o/${tag}_*.json.tmp
ls -la $o | grep ${tag}_ | awk '{print $5, $9}' | head -60
