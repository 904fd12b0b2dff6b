set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
tail -12 gpurun_out/c11_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; tail -c 800 gpurun_out/c11_bench.err
timeout 300 python tools/step_breakdown.py --out gpurun_out/c11_breakdown_w8a8.json > gpurun_out/c11_breakdown_w8a8.txt 2>&1
python tools/crit_path.py gpurun_out/c11_breakdown_w8a8.json 24
head -c 300 gpurun_out/c11_bench.json
timeout 300 python tools/phase_sweep.py > gpurun_out/c11_phase_sweep.txt 2>&1
grep "mode=0" gpurun_out/c11_phase_sweep.txt | grep "N=1280 \|N=640 " | cut -c1-230
