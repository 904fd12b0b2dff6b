set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
tail -12 gpurun_out/c17_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; tail -c 800 gpurun_out/c17_bench.err
head -c 260 gpurun_out/c17_bench.json; echo; head -c 260 gpurun_out/c17_bench_mode1.json; echo
timeout 300 python tools/step_breakdown.py --out gpurun_out/c17_breakdown_w8a8.json > gpurun_out/c17_breakdown_w8a8.txt 2>&1
python tools/crit_path.py gpurun_out/c17_breakdown_w8a8.json 24
timeout 200 python tools/quant_phase.py 256 1280 1 > gpurun_out/c17_quant_phase_256x1280.txt 2>&1
sed -n 2,9p gpurun_out/c17_quant_phase_256x1280.txt | cut -c1-220
