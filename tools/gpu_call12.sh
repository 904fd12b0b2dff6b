set -x
for b in flash efficient cudnn; do
MIXDQ_SDPA_BACKEND=$b timeout 300 python bench.py --no-cpu-baseline --no-fp16 > gpurun_out/c12_bench_sdpa_$b.json 2> gpurun_out/c12_bench_sdpa_$b.err
head -c 260 gpurun_out/c12_bench_sdpa_$b.json; echo; tail -2 gpurun_out/c12_bench_sdpa_$b.err
done
