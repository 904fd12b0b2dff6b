"""Per-kernel breakdown of one UNet step INSIDE the whole-UNet CUDA graph (CUPTI activity records via
torch.profiler: real back-to-back durations, unlike ncu's serialised cold-cache replays).

  python tools/step_breakdown.py [--fp16] [--batch B] [--model sdxl-turbo] [--out gpurun_out/x.json]

Prints, per kernel family: launches per step, total us, mean us; plus the GPU idle time between
kernels (step wall - sum of kernel durations). Development aid, not part of the product path.
"""
import argparse
import collections
import json
import re
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"mixdq::tc_i8_kernel<(\d+),\s*(\d+),\s*(\d+)(?:,\s*\(?\w*\)?(\w+))?>", name)
    if m:
        w4 = ",W4" if m.group(4) in ("true", "1") else ""
        return f"tc_i8<BN={m.group(1)},ST={m.group(2)},KIND={m.group(3)}{w4}>"
    if name.startswith("at::"):
        parts = re.findall(r"at::native::(?:\(anonymous namespace\)::|<unnamed>::)?(\w+)", name)
        if parts:
            return "at::" + "/".join(parts[:3])
    return re.sub(r"\(.*", "", name)[:80]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fp16", action="store_true")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--model", default="sdxl-turbo")
    ap.add_argument("--mode", default="dynamic")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    from mixdq_b200 import _lib
    _lib.load()
    unet16 = bench.build_fp16_unet(args.model, dev, seed=0)
    inputs = unet16.example_inputs(args.batch, dev, torch.float16, seed=1)
    unet = unet16 if args.fp16 else bench.quantize_copy(unet16, args.mode)
    graph, _ = bench.capture(unet, inputs)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 10
    from torch.profiler import ProfilerActivity, profile
    reps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            graph.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
           and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    fam = collections.defaultdict(lambda: [0, 0.0])
    for e in evs:
        f = fam[short(e.name)]
        f[0] += 1
        f[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    tot = sum(v[1] for v in fam.values()) / reps
    n = sum(v[0] for v in fam.values()) / reps
    # GPU busy span of the last replay (first kernel start .. last kernel end)
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // reps
    last = evs[-per:]
    t0 = min(e.time_range.start for e in last)
    span = (max(e.time_range.end for e in last) - t0)
    busy = sum(e.time_range.end - e.time_range.start for e in last)
    print(f"last replay: span {span:.1f} us, sum of durations {busy:.1f} us (overlap via PDL makes sum > span possible)")
    print(f"{'fp16' if args.fp16 else 'w8a8'} {args.model} B={args.batch}: step {step_ms:.3f} ms (events), "
          f"{n:.0f} kernels/step, sum of kernel durations {tot / 1e3:.3f} ms, traced span {span / 1e3:.3f} ms")
    rows = []
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        rows.append({"kernel": k, "per_step": v[0] / reps, "us_per_step": v[1] / reps,
                     "mean_us": v[1] / v[0]})
        print(f"  {v[1] / reps:9.1f} us  {v[0] / reps:6.0f} x {v[1] / v[0]:7.2f} us  {k}")
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        json.dump({"step_ms": step_ms, "kernels_per_step": n, "sum_kernel_ms": tot / 1e3,
                   "span_ms": span / 1e3, "rows": rows,
                   "timeline": [[short(e.name), (e.time_range.start - t0), e.time_range.end - e.time_range.start]
                                for e in last]}, open(args.out, "w"))


if __name__ == "__main__":
    main()
