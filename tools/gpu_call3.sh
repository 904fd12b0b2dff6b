set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
tail -15 gpurun_out/c3_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; tail -c 1500 gpurun_out/c3_bench.err
timeout 300 python tools/step_breakdown.py --out gpurun_out/c3_breakdown_w8a8.json > gpurun_out/c3_breakdown_w8a8.txt 2>&1
python tools/crit_path.py gpurun_out/c3_breakdown_w8a8.json 24
head -c 300 gpurun_out/c3_bench.json
