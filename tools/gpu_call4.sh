set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
tail -5 gpurun_out/c4_pytest.log
timeout 200 python tools/quant_phase.py 256 1280 > gpurun_out/c4_quant_phase.txt 2>&1
MIXDQ_CARVEOUT=100 timeout 200 python tools/quant_phase.py 256 1280 > gpurun_out/c4_quant_phase_carve100.txt 2>&1
timeout 400 python tools/phase_sweep.py > gpurun_out/c4_phase_sweep.txt 2>&1
MIXDQ_CARVEOUT=100 timeout 300 python bench.py --no-cpu-baseline --no-fp16 > gpurun_out/c4_bench_carve100.json 2> gpurun_out/c4_bench_carve100.err
cat gpurun_out/c4_quant_phase.txt | tail -12
head -c 400 gpurun_out/c4_bench_carve100.json
