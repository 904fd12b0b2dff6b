#!/bin/bash
# full GPU suite + the headline lines on one box
tag=${1:-rb}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log
timeout 600 python bench.py > $o/${tag}_bench_c2.json 2> $o/${tag}_bench_c2.err; tail -c 300 $o/${tag}_bench_c2.err
timeout 400 python bench.py --config 4 --no-cpu-baseline > $o/${tag}_bench_c4_n1.json 2> $o/${tag}_bench_c4_n1.err; tail -c 300 $o/${tag}_bench_c4_n1.err
timeout 400 python bench.py --config 3 --no-cpu-baseline > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err; tail -c 300 $o/${tag}_bench_c3.err
timeout 400 python bench.py --batch 8 --mode static --no-cpu-baseline > $o/${tag}_bench_b8_static.json 2> $o/${tag}_bench_b8_static.err; tail -c 300 $o/${tag}_bench_b8_static.err
python - <<PY
import json
for f in ("bench_c2", "bench_c4_n1", "bench_c3", "bench_b8_static"):
    try:
        d = json.loads(open("$o/${tag}_%s.json" % f).readline())
        print(f, round(d["ms_per_step"], 3), round(d["value"], 1), d.get("fp16_baseline") and round(d["fp16_baseline"]["speedup_w8a8_over_fp16"], 3),
              d.get("static_scales") and (round(d["static_scales"]["ms_per_step"], 3), round(d["static_scales"]["speedup_over_fp16"], 3)),
              "e2e", round(d["e2e"]["ms_per_step"], 3), d["roofline"]["bound"], round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -2
