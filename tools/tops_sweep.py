"""INT8 TOP/s of the tcgen05 GEMM / conv kernels on tensor-bound shapes (SDXL layer shapes at
batch 8 / 32), back-to-back launches inside a CUDA graph, weights rotated through several buffers.
Prints TOP/s and the fraction of 2x the measured dense bf16 peak (MEASURED_PEAKS.json)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
peaks = json.loads(Path("MEASURED_PEAKS.json").read_text()) if Path("MEASURED_PEAKS.json").exists() else {}
PEAK = 2 * peaks.get("bf16_tflops", 1671.4)


def graph_time(fns, iters=5):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / len(fns) * 1e-3   # seconds per launch


def main():
    rows = []
    for (M, N, K) in [(2048, 1280, 1280), (2048, 10240, 1280), (2048, 1280, 5120), (8192, 640, 640),
                      (8192, 5120, 640), (8192, 640, 2560), (8192, 1280, 1280), (8192, 10240, 1280),
                      (32768, 320, 320), (16384, 2560, 2560)]:
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
        ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(4)]
        z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
        outs = []
        t = graph_time([(lambda w=w: outs.append(ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)))
                        for w in ws] * 2)
        tops = 2.0 * M * N * K / t / 1e12
        rows.append(("gemm", M, N, K, t * 1e6, tops))
        print(f"gemm  M={M:6d} N={N:6d} K={K:6d}  {t*1e6:8.1f} us  {tops:7.1f} TOP/s  {100*tops/PEAK:5.1f} % of 2x bf16 "
              f"({lib.mixdq_last_path().decode()})", flush=True)
        del outs, ws, a
    for (n, h, w_, c, k) in [(8, 64, 64, 320, 320), (8, 32, 32, 640, 640), (8, 16, 16, 1280, 1280),
                             (8, 64, 64, 640, 320), (32, 32, 32, 640, 640)]:
        x = torch.randint(-128, 128, (n, c, h, w_), dtype=torch.int8, device=dev).contiguous(
            memory_format=torch.channels_last)
        ws = [torch.randint(-127, 128, (k, c, 3, 3), dtype=torch.int8, device=dev).contiguous(
            memory_format=torch.channels_last) for _ in range(3)]
        sc = torch.ones(k, device=dev); s1 = torch.tensor(1.0, device=dev); zp = torch.tensor(3.0, device=dev)
        wsum = [w.float().sum(1, keepdim=True).contiguous() for w in ws]
        outs = []
        t = graph_time([(lambda w=w, s=s: outs.append(ops.qconv2d_w8_a8_ohalf(x, w, sc, s1, zp, sc, s, None, None, 1, 1, 1)))
                        for w, s in zip(ws, wsum)] * 2)
        tops = 2.0 * n * h * w_ * k * c * 9 / t / 1e12
        print(f"conv3x3 n={n} {h}x{w_} c={c} k={k}  {t*1e6:8.1f} us  {tops:7.1f} TOP/s  {100*tops/PEAK:5.1f} % of 2x bf16",
              flush=True)
        del outs, ws, x


if __name__ == "__main__":
    main()
