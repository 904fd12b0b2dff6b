set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
tail -12 gpurun_out/c13_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; tail -c 800 gpurun_out/c13_bench.err
timeout 300 python tools/step_breakdown.py --out gpurun_out/c13_breakdown_w8a8.json > gpurun_out/c13_breakdown_w8a8.txt 2>&1
python tools/crit_path.py gpurun_out/c13_breakdown_w8a8.json 24
head -c 300 gpurun_out/c13_bench.json
timeout 200 python tools/quant_phase.py 256 1280 > gpurun_out/c13_quant_phase_256x1280.txt 2>&1
sed -n 2,7p gpurun_out/c13_quant_phase_256x1280.txt | cut -c1-220
