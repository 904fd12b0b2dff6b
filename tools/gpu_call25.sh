set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c25_pytest.log
tail -4 gpurun_out/c25_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err; tail -c 300 gpurun_out/c25_bench.err
head -c 260 gpurun_out/c25_bench.json; echo
timeout 300 python tools/step_breakdown.py --out gpurun_out/c25_breakdown_w8a8.json > gpurun_out/c25_breakdown_w8a8.txt 2>&1
python tools/crit_path.py gpurun_out/c25_breakdown_w8a8.json 12
