set -x
timeout 120 tools/mma_bench > gpurun_out/c2_mma_bench.txt 2>&1
timeout 120 tools/mma_bench2 > gpurun_out/c2_mma_bench2.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -15 gpurun_out/c2_pytest.log
timeout 600 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; tail -c 1500 gpurun_out/c2_bench.err
timeout 300 python tools/step_breakdown.py --out gpurun_out/c2_breakdown_w8a8.json > gpurun_out/c2_breakdown_w8a8.txt 2>&1
MIXDQ_NO_CLUSTER=1 timeout 300 python bench.py --no-cpu-baseline --no-fp16 > gpurun_out/c2_bench_nocluster.json 2> gpurun_out/c2_bench_nocluster.err
head -c 300 gpurun_out/c2_bench.json; echo; head -c 300 gpurun_out/c2_bench_nocluster.json
