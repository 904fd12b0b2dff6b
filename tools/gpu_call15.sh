set -x
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:tc_i8|ln_minmax|minmax_rows|premm|gn_quant" -s 15 -c 15 -o gpurun_out/r01_repr_full -f python tools/ncu_repr.py > gpurun_out/c15_ncu.log 2>&1
tail -3 gpurun_out/c15_ncu.log; ls -la gpurun_out/*.ncu-rep
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c15_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp16 > gpurun_out/c15_ncu_bench.log 2>&1
wc -l gpurun_out/c15_launches_bench.csv
timeout 600 python bench.py --batch 8 --no-cpu-baseline > gpurun_out/c15_bench_b8.json 2> gpurun_out/c15_bench_b8.err; head -c 300 gpurun_out/c15_bench_b8.json; tail -c 300 gpurun_out/c15_bench_b8.err
