set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -3 gpurun_out/c1_pytest.log
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 600 gpurun_out/c1_bench.err
timeout 300 python tools/step_breakdown.py --out gpurun_out/c1_breakdown_w8a8.json > gpurun_out/c1_breakdown_w8a8.txt 2>&1
timeout 300 python tools/step_breakdown.py --fp16 --out gpurun_out/c1_breakdown_fp16.json > gpurun_out/c1_breakdown_fp16.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp16 > gpurun_out/c1_ncu_bench.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c1_traffic_eager.csv python bench.py --profile-step > gpurun_out/c1_ncu_traffic.log 2>&1
wc -l gpurun_out/*.csv
