set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
tail -5 gpurun_out/c14_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err; tail -c 500 gpurun_out/c14_bench.err
MIXDQ_A_PREFETCH=0 timeout 300 python bench.py --no-cpu-baseline --no-fp16 > gpurun_out/c14_bench_noapf.json 2> gpurun_out/c14_bench_noapf.err
head -c 260 gpurun_out/c14_bench.json; echo; head -c 260 gpurun_out/c14_bench_noapf.json; echo
timeout 120 python tools/ncu_repr.py
timeout 900 ncu --set full --import-source on --clock-control none -k regex:mixdq -s 26 -c 26 -o gpurun_out/r01_repr_full -f python tools/ncu_repr.py > gpurun_out/c14_ncu.log 2>&1
tail -3 gpurun_out/c14_ncu.log; ls -la gpurun_out/*.ncu-rep
