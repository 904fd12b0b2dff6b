"""Critical-path attribution from a tools/step_breakdown.py JSON: kernels sorted by END time, each
charged end[i] - end[i-1] (what it adds to the serial chain), next to its own duration."""
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
tl = sorted(d["timeline"], key=lambda e: e[1] + e[2])
prev = 0.0
fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
for n, s, dur in tl:
    end = s + dur
    f = fam[n]
    f[0] += 1; f[1] += end - prev; f[2] += dur
    prev = end
print(f"step span {prev:.1f} us, {len(tl)} kernels")
for n, v in sorted(fam.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"  {v[1]:8.1f} us  {v[0]:4d} x  inc {v[1] / v[0]:6.2f}  dur {v[2] / v[0]:6.2f}  {n[:72]}")
