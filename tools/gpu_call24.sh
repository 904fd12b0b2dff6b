set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c24_pytest.log
tail -5 gpurun_out/c24_pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err; tail -c 600 gpurun_out/c24_bench.err
head -c 300 gpurun_out/c24_bench.json; echo
timeout 300 python tools/step_breakdown.py --out gpurun_out/c24_breakdown_w8a8.json > gpurun_out/c24_breakdown_w8a8.txt 2>&1
python tools/crit_path.py gpurun_out/c24_breakdown_w8a8.json 30
python -c "from __graft_entry__ import smoke; smoke()"
